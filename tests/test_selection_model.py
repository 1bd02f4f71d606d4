"""CPU: the selection algebra of the fused kernel (csrc/match_kernel.cu), restated in numpy and checked by
randomised properties.  These are the three facts the kernel's exactness argument rests on:

  1. keys6_insert2: inserting a sorted pair x <= y into an ascending 6-list with
         c_j = min(a_j, max(a_{j-1}, x), max(a_{j-2}, y))            keeps the six smallest, sorted;
  2. pair_scan_round: min(k[j], partner[5-j]) over two ascending 6-lists are the six smallest of the union;
  3. block_scan_private / finish_private: ranking on packed keys (bits(d~) & ~127) | ordinal with a ranking
     distance d~ within 4 ulp of the exact one, keeping the six smallest keys and re-ranking them exactly, gives
     the exact five nearest (ties by ordinal) WHENEVER the acceptance test passes
         bucket(key6) >= bucket(exact d5) + 2 * 128."""
import numpy as np

INF = np.float32(np.inf)


def insert2(a, x, y):
    x, y = min(x, y), max(x, y)
    c = a.copy()
    c[0] = min(a[0], x)
    c[1] = min(a[1], max(a[0], x), y)
    for j in range(2, 6):
        c[j] = min(a[j], max(a[j - 1], x), max(a[j - 2], y))
    return c


def test_pair_insertion_identity():
    rng = np.random.default_rng(0)
    for _ in range(3000):
        n_fin = rng.integers(0, 7)
        a = np.sort(np.concatenate([rng.random(n_fin), np.full(6 - n_fin, np.inf)])).astype(np.float32)
        x, y = (np.float32(v) for v in rng.random(2))
        if rng.random() < 0.2:
            y = INF
        if rng.random() < 0.1:
            x = a[rng.integers(6)]                     # duplicates of list entries (only +inf occurs on the device)
        want = np.sort(np.concatenate([a, [x, y]]))[:6]
        assert np.array_equal(insert2(a, x, y), want)


def test_pair_merge_identity():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        k = np.sort(np.where(rng.random(6) < 0.2, np.inf, rng.random(6))).astype(np.float32)
        p = np.sort(np.where(rng.random(6) < 0.2, np.inf, rng.random(6))).astype(np.float32)
        m = np.sort(np.minimum(k, p[::-1]))
        assert np.array_equal(m, np.sort(np.concatenate([k, p]))[:6])


def _bits(f):
    return np.asarray(f, np.float32).view(np.uint32)


def test_packed_key_acceptance_is_sound():
    rng = np.random.default_rng(2)
    accepted = rejected = 0
    for trial in range(4000):
        n = int(rng.integers(1, 100))
        base = rng.random(n).astype(np.float32) * np.float32(2.0) + np.float32(1e-4)
        if trial % 3 == 0:                              # adversarial: clusters of almost equal distances
            centre = np.float32(rng.random() + 0.01)
            m = int(rng.integers(1, min(n, 12) + 1))
            base[:m] = (_bits(np.full(m, centre)) + rng.integers(0, 300, m).astype(np.uint32)).view(np.float32)
        if trial % 7 == 0:
            base[rng.integers(n)] = base[rng.integers(n)]                       # exact ties
        d = base
        rank_d = (_bits(d).astype(np.int64) + rng.integers(-4, 5, n)).astype(np.uint32).view(np.float32)   # within 4 ulp
        ordn = np.arange(n, dtype=np.uint32)
        keys = (_bits(rank_d) & ~np.uint32(127)) | ordn
        order = np.argsort(keys, kind="stable")
        kept = order[:6]
        k6 = keys[order[5]] if n >= 6 else None
        # exact re-rank of the kept candidates by (exact distance bits, ordinal)
        kept_sorted = sorted(kept.tolist(), key=lambda i: (int(_bits(d[i])), int(ordn[i])))
        top5_kept = kept_sorted[:5]
        truth = sorted(range(n), key=lambda i: (int(_bits(d[i])), int(ordn[i])))[:5]
        if n < 6:
            ok = True
        else:
            d5 = _bits(d[top5_kept[4]]) if len(top5_kept) == 5 else np.uint32(0x7F800000)
            ok = int(k6 & ~np.uint32(127)) >= int(d5 & ~np.uint32(127)) + 256
        if ok:
            accepted += 1
            assert top5_kept == truth, (trial, n)
        else:
            rejected += 1
    assert accepted > 2000 and rejected > 50            # both branches are exercised


# ---- 4. block_is_final: the geometric guarantee behind "the block's answer is the global answer" -----------------
F = np.float32


def _cell_units(x, o, inv):
    return F(F(x - o) * inv)


def _cell_coord(x, o, inv, n):
    u = _cell_units(x, o, inv)
    c = int(min(np.floor(u), 2.0e9)) if u >= 0 else 0
    return min(c, n - 1)


def _guaranteed_r2(q, h, g):
    """block_is_final's gr2 (match_kernel.cu): squared radius around q inside which every map point is in the block."""
    ox, inv, cell, n = g
    u = [_cell_units(q[a], ox[a], inv) for a in range(3)]
    f = [F(u[a] - F(h[a])) for a in range(3)]
    slack = F(F(F(6.0e-7) * max(abs(u[0]), abs(u[1]), abs(u[2]))) + F(1.0e-6))
    m = None
    for a in range(3):
        if h[a] - 1 > 0:
            v = F(f[a] + F(1.0))
            m = v if m is None else min(m, v)
        if h[a] + 1 < n[a] - 1:
            v = F(F(F(1.0) - f[a]) + F(1.0))
            m = v if m is None else min(m, v)
    if m is None:
        return np.inf
    me = F(F(F(m - slack) * cell) * F(0.999999))
    return float(F(me * me)) if me > 0 else 0.0


def _sqdist(q, p):
    dx, dy, dz = F(q[0] - p[0]), F(q[1] - p[1]), F(q[2] - p[2])
    return float(F(F(dx * dx) + F(F(dy * dy) + F(dz * dz))))


def test_block_guarantee_radius_is_sound():
    """No point binned OUTSIDE the 3x3x3 block of a query can be closer than the radius block_is_final grants —
    including points a few ulp from a cell face, far from the origin (large cell indices) and at the grid rim."""
    rng = np.random.default_rng(3)
    worst = np.inf
    for trial in range(400):
        cell = F(rng.choice([0.25, 0.375, 0.5625, 0.84375, 1.265625, 1.8984375, 0.15, 0.6]))
        inv = F(F(1.0) / cell)
        n = [int(rng.integers(3, 2000)) for _ in range(3)]
        ox = [F(rng.uniform(-500, 500)) for _ in range(3)]
        g = (ox, inv, cell, n)
        # a query somewhere in the grid (sometimes in the rim cells)
        hq = [int(rng.integers(0, n[a])) if rng.random() < 0.8 else int(rng.choice([0, 1, n[a] - 2, n[a] - 1])) for a in range(3)]
        q = [F(F(ox[a]) + F((hq[a] + rng.random()) * float(cell))) for a in range(3)]
        h = [_cell_coord(q[a], ox[a], inv, n[a]) for a in range(3)]
        r2 = _guaranteed_r2(q, h, g)
        for _ in range(60):
            # candidate point: near a face of the block, just outside it along one axis (adversarial), anywhere else inside
            p = [F(F(ox[a]) + F((h[a] - 1 + 3 * rng.random()) * float(cell))) for a in range(3)]
            a = int(rng.integers(3))
            side = rng.random() < 0.5
            face = (h[a] + 2) if side else (h[a] - 1)
            x = F(F(ox[a]) + F(face * float(cell)))
            steps = int(rng.integers(0, 40))
            for _s in range(steps):                       # walk a few ulp away from the face, outwards
                x = np.nextafter(x, F(np.inf) if side else F(-np.inf))
            if rng.random() < 0.5:
                x = F(x + F((1 if side else -1) * rng.random() * 0.02 * float(cell)))
            p[a] = x
            c = [_cell_coord(p[b], ox[b], inv, n[b]) for b in range(3)]
            outside = any(c[b] < max(h[b] - 1, 0) or c[b] > min(h[b] + 1, n[b] - 1) for b in range(3))
            if not outside:
                continue
            d2 = _sqdist(q, p)
            assert d2 >= r2, (trial, q, p, h, c, d2, r2)
            if r2 > 0:
                worst = min(worst, d2 / r2)
    assert worst < 1.5                                      # the adversarial points really probe the margin


# ---- the scan's storage permutation (flimo_api.cu: coprime_stride / mod_inverse; match_kernel.cu: Barrett) ----
def _coprime_stride(n):
    from math import gcd
    if n < 3:
        return 1
    s = int(0.6180339887 * n) | 1
    while gcd(s, n) != 1:
        s += 2
    return s % n if s % n else 1


def _mod_inverse(a, n):
    t, nt, r, nr = 0, 1, n, a % n
    while nr:
        q = r // nr
        t, nt = nt, t - q * nt
        r, nr = nr, r - q * nr
    return t + n if t < 0 else t


def _barrett(prod, n, magic):
    r = prod - ((prod * magic) >> 64) * n          # __umul64hi(prod, magic) * n
    return r - n if r >= n else r


def test_scan_permutation_model():
    """Position j of the permuted order holds original point (j * s^-1) mod n, where the upload kernel stores point i
    at (i * s) mod n; the kernel evaluates the modulus with ONE multiply-high and ONE conditional subtraction."""
    rng = np.random.default_rng(5)
    sizes = [3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 17, 255, 256, 257, 1000, 4095, 4096, 16384, 131072, 300000, (1 << 20) - 1, 1 << 20,
             (1 << 22) + 1] + [int(v) for v in rng.integers(3, 1 << 22, 40)]
    for n in sizes:
        s = _coprime_stride(n)
        inv = _mod_inverse(s, n)
        assert 0 < s < n and (s * inv) % n == 1
        magic = ((1 << 64) - 1) // n
        qs = np.unique(np.concatenate([np.arange(min(n, 64)), np.arange(max(n - 64, 0), n), rng.integers(0, n, 200)]))
        for q in qs:
            q = int(q)
            orig = _barrett(q * inv, n, magic)
            assert orig == (q * inv) % n                   # the reduction is exact
            assert (orig * s) % n == q                     # and inverts the upload kernel's placement
    # a full bijection check on a few sizes
    for n in (3, 10, 257, 4096, 131072):
        s, magic = _coprime_stride(n), ((1 << 64) - 1) // n
        inv = _mod_inverse(s, n)
        j = np.arange(n, dtype=object)
        orig = np.array([_barrett(int(v) * inv, n, magic) for v in j[: min(n, 20000)]])
        assert len(set(orig.tolist())) == len(orig)
    # worst case for the single conditional subtraction: products up to (n - 1)^2 < 2^63
    for n in (3, (1 << 20) + 7, (1 << 31) - 1):
        magic = ((1 << 64) - 1) // n
        for prod in ((n - 1) * (n - 1), (n - 1) * (n - 2), n * (n - 1) - 1, n - 1, n, 0):
            assert _barrett(prod, n, magic) == prod % n


# ---- round 2: order-free pass sums, team-rescan pruning, in-place row merge (numpy / pure-Python models) ----------------
def _fx_split(s):
    """match_kernel.cu (fx_reduce): a float64 tile sum as two fixed-point integers, units 2^-18 and 2^-66."""
    hi = int(np.rint(s * 2.0 ** 18))
    lo = int(np.rint((s - hi * 2.0 ** -18) * 2.0 ** 66))
    return hi, lo


def test_fixed_point_pass_sums_are_exact_and_order_free():
    """The tiles add (hi, lo) with integer REDs; the finisher returns hi * 2^-18 + lo * 2^-66.  Integer addition is
    associative, so any arrival order gives the same bits, and the result is the correctly rounded exact sum up to the
    2^-66 granularity of a tile — closer to it than a float64 tree."""
    import math
    rng = np.random.default_rng(3)
    for scale in (1e-6, 1.0, 4.0e4, 3.0e9):                          # sum z^2 ... sum of 150 m lever arms squared
        tiles = (rng.standard_normal(2344) * scale).astype(np.float64)
        tiles[::7] *= 1e-9
        parts = [_fx_split(float(t)) for t in tiles]
        for hi, lo in parts:
            assert abs(hi) < 2 ** 62 and abs(lo) <= 2 ** 47          # rem <= 2^-19 -> lo <= 2^47: 2^15 tiles stay below 2^63
        ref = None
        for _ in range(5):
            order = rng.permutation(len(parts))
            H = sum(parts[i][0] for i in order)
            L = sum(parts[i][1] for i in order)
            assert -2 ** 63 <= H < 2 ** 63 and -2 ** 63 <= L < 2 ** 63
            got = float(H) * 2.0 ** -18 + float(L) * 2.0 ** -66
            ref = got if ref is None else ref
            assert got == ref                                         # bit-identical whatever the order
        exact = math.fsum(tiles.tolist())
        assert abs(ref - exact) <= 2344 * 2.0 ** -66 + abs(exact) * 2.0 ** -52


def test_team_rescan_bound_keeps_the_exact_answer():
    """block_scan_team: candidates with d2 > bound skip the insertion, where bound = the exact 5th squared distance inside a
    smaller block scanned before (an upper bound of the true one).  The five nearest of the larger block are unchanged."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        big = rng.random((int(rng.integers(6, 200)), 3)).astype(np.float32)
        q = rng.random(3).astype(np.float32)
        d2 = ((big - q) ** 2).sum(1)
        small = rng.choice(len(big), size=int(rng.integers(5, len(big) + 1)), replace=False)     # the block scanned first
        bound = np.sort(d2[small])[4]
        keep = d2 <= bound
        full = np.lexsort((np.arange(len(big)), d2))[:5]
        pruned_idx = np.flatnonzero(keep)
        pruned = pruned_idx[np.lexsort((pruned_idx, d2[keep]))[:5]]
        assert np.array_equal(full, pruned)


def _merge_row_in_place(row, cnt, new, cap, B=8):
    """upd_merge_kernel (map_index.cu), in place: row[:cnt] old entries (key, id) sorted by key, `new` sorted by key.
    Old entries move up by the number of new entries with a SMALLER key, back to front in chunks of B 'threads'
    (reads of a chunk complete before its writes); the new entries drop into the holes behind the old entries of their own
    and all earlier cells."""
    keys_new = [k for k, _ in new]
    import bisect
    first_new = keys_new[0]
    keep = bisect.bisect_left([k for k, _ in row[:cnt]], first_new)     # = slot[first new cell] - base
    old_keys = [k for k, _ in row[:cnt]]                               # (the kernel reads these from the table slots)
    span = cnt - keep
    for c in reversed(range((span + B - 1) // B)):
        idx = [keep + c * B + t for t in range(B) if keep + c * B + t < cnt]
        vals = [row[i] for i in idx]                                   # all reads of the chunk ...
        for i, v in zip(idx, vals):                                    # ... then all writes
            lo = bisect.bisect_left(keys_new, v[0])
            if lo:
                assert i + lo < cap
                row[i + lo] = v
    for i, v in enumerate(new):
        row[i + bisect.bisect_right(old_keys, v[0])] = v               # old entries with key <= own key stay in front
    return cnt + len(new)


def test_in_place_row_merge_model():
    rng = np.random.default_rng(5)
    for _ in range(400):
        nx = int(rng.integers(3, 40))
        cnt, k = int(rng.integers(0, 120)), int(rng.integers(1, 60))
        old = sorted((int(rng.integers(0, nx)), 1000 + i) for i in range(cnt))
        old = sorted(old, key=lambda e: e[0])
        new = sorted(((int(rng.integers(0, nx)), 5000 + i) for i in range(k)), key=lambda e: e[0])
        cap = cnt + k + int(rng.integers(0, 9))
        row = old + [(-1, -1)] * (cap - cnt)
        n = _merge_row_in_place(row, cnt, new, cap, B=int(rng.choice([1, 4, 8, 32])))
        want = sorted(old + new, key=lambda e: (e[0], e[1] >= 5000))   # by cell; old entries first inside a cell, orders kept
        assert n == cnt + k and row[:n] == want
