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
