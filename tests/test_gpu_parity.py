"""GPU (-m gpu): the CUDA path through the C ABI against the CPU oracle and the golden fixtures.

Tolerances:
  per-point float32 results (world point, neighbour distances, plane, signed distance, accept flag):
      BIT-EXACT — the kernel rounds like the oracle (--fmad=false, same operation order);
  float64 reductions (HTH, HTh): relative 1e-12 (summation order differs);
  pose after each pass: |dp| <= 1e-9 m, |dq| <= 1e-10 (the reference-level tolerance of 1e-4 m /
      1e-5 rad is met with five orders of margin).
"""
import os

import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BIG = 1 << 18


def mapper(**kw):
    kw.setdefault("MAX_NUM_MATCHES", BIG)
    kw.setdefault("MAX_NUM_PC2MATCH", BIG)
    return api.Mapper(api.MappingConfig(**kw), device=0)


def check_per_point(dbg, ref, max_dist=2.0):
    assert np.array_equal(dbg["world"], ref["world"])
    assert np.array_equal(dbg["good"], ref["good"])
    g = ref["good"]
    assert np.array_equal(dbg["plane"][g], ref["plane"][g])
    assert np.array_equal(dbg["dist"][g], ref["dist"][g])
    close = ref["nn_d2"][:, -1] < max_dist           # beyond MAX_DIST_PLANE the search may stop early (outcome-equivalent)
    assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])


@pytest.mark.parametrize("name", ["tiny", "c1"])
@pytest.mark.parametrize("cell,sort", [(0.0, True), (0.15, False), (0.6, True)])
def test_per_point_and_reduction_parity(oracle, flimo_lib, name, cell, sort):
    case = synth.make_case(name)
    m = mapper(knn_cell=cell, sort_scan=sort)
    m.add(case.map_pts, 1.5)
    assert m.exists() and m.size() == case.map_pts.shape[0] and m.last_time() == 1.5
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4)
    for st in (case.init, case.truth):
        ref = om.match(ocfg, st[:14], case.scan)
        m.set_scan(case.scan)
        check_per_point(m.match_debug(st), ref)
        r = m.match(st)
        assert (r.n_valid, r.n_rows) == (ref["n_valid"], ref["rows"])
        assert np.allclose(r.HTH, ref["HTH"], rtol=1e-12, atol=1e-12 * np.abs(ref["HTH"]).max())
        assert np.allclose(r.HTh, ref["HTh"], rtol=1e-12, atol=1e-12 * np.abs(ref["HTh"]).max())
        assert np.isclose(r.sum_sq_res, float((ref["dist"][ref["good"]].astype(np.float64) ** 2).sum()), rtol=1e-12)


@pytest.mark.parametrize("name,pc2,mm", [("tiny", BIG, BIG), ("tiny", 1500, 400), ("c1", BIG, BIG)])
def test_against_golden_fixture(flimo_lib, name, pc2, mm):
    g = np.load(os.path.join(G, f"{name}_m{mm}.npz"))
    case = synth.make_case(name)
    m = mapper(MAX_NUM_PC2MATCH=pc2, MAX_NUM_MATCHES=mm)
    m.add(case.map_pts)
    m.set_scan(case.scan)
    good = np.unpackbits(g["good"])[: int(g["n_q"])].astype(bool)
    dbg = m.match_debug(case.init)
    assert dbg["good"].shape[0] == int(g["n_q"])            # first-N rule of MAX_NUM_PC2MATCH
    assert np.array_equal(dbg["good"], good)
    assert np.array_equal(dbg["plane"][good], g["plane"]) and np.array_equal(dbg["dist"][good], g["dist"])
    r = m.match(case.init)
    assert (r.n_valid, r.n_rows) == (int(g["n_valid"]), int(g["rows"]))   # first-N rule of MAX_NUM_MATCHES
    assert np.allclose(r.HTH, g["HTH"], rtol=1e-12, atol=1e-12 * np.abs(g["HTH"]).max())
    x, P, passes = m.update(case.init, synth.default_P0(), int(g["max_iter"]), 0.0)
    assert passes == len(g["trace_rows"])
    assert np.abs(x[:3] - g["x_final"][:3]).max() <= 1e-9 and np.abs(x[3:7] - g["x_final"][3:7]).max() <= 1e-10
    assert np.allclose(x, g["x_final"], atol=1e-9)
    assert np.allclose(P, g["P_final"], rtol=1e-4, atol=1e-11)      # P = L - K P cancels ~5 digits


def test_iterated_update_per_pass_parity(oracle, flimo_lib):
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=2)
    for max_iter, lim in ((0, 0.001), (3, 0.0), (4, 1e9)):
        xo, Po, tr = om.update(ocfg, case.init, synth.default_P0(), max_iter, lim, case.scan)
        m.ekf_begin(case.init, synth.default_P0(), max_iter, lim)
        k, done = 0, False
        while not done:
            r = m.match(m.ekf_state())
            assert r.n_rows == tr[k]["rows"]
            done = m.ekf_step(r.HTH, r.HTh, r.n_rows)
            assert np.abs(m.ekf_state() - tr[k]["state"]).max() <= 1e-9
            k += 1
        x, P = m.ekf_end()
        assert k == len(tr)
        assert np.abs(x - xo).max() <= 1e-9 and np.allclose(P, Po, rtol=1e-4, atol=1e-11)
        assert np.abs(x[:3] - case.truth[:3]).max() < 0.01          # and it actually registers


def test_extrinsics_and_flags(oracle, flimo_lib):
    case = synth.make_case("tiny")
    offR = synth.quat_from_rpy(0.02, -0.01, 0.03)
    st = case.init.copy()
    st[7:11] = offR
    st[11:14] = [0.1, -0.05, 0.2]
    om = oracle.OracleMap()
    om.add(case.map_pts)
    for ee in (True, False):
        m = mapper(estimate_extrinsics=ee)
        m.add(case.map_pts)
        ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, estimate_extrinsics=ee), st[:14], case.scan)
        r = m.match(st, case.scan)
        assert r.n_valid == ref["n_valid"]
        assert np.allclose(r.HTH, ref["HTH"], rtol=1e-12, atol=1e-12 * np.abs(ref["HTH"]).max())
        if not ee:
            assert np.all(r.HTH[6:, :] == 0) and np.all(r.HTh[6:] == 0)


def test_thresholds_follow_config(oracle, flimo_lib):
    case = synth.make_case("tiny")
    om = oracle.OracleMap()
    om.add(case.map_pts)
    for mdp, thr in ((0.05, 0.05), (2.0, 0.01), (0.3, 0.2)):
        m = mapper(MAX_DIST_PLANE=mdp, PLANE_THRESHOLD=thr)
        m.add(case.map_pts)
        m.set_scan(case.scan)
        ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, max_dist_plane=mdp, plane_threshold=thr),
                       case.init[:14], case.scan)
        check_per_point(m.match_debug(case.init), ref, max_dist=mdp)


def test_edge_cases(oracle, flimo_lib):
    st = synth.make_state([0, 0, 0], [0, 0, 0, 1])
    m = mapper()
    # empty map: Mapper::match returns no matches (Mapper.cpp:61); the update leaves the state alone
    r = m.match(st, np.zeros((10, 3), np.float32))
    assert (r.n_valid, r.n_rows) == (0, 0) and not m.exists() and m.size() == 0 and np.all(r.HTH == 0)
    x, P, passes = m.update(st, synth.default_P0(), 2, 0.001)
    assert passes == 2 and np.array_equal(x, st)
    # a map with fewer than 5 points can never produce a match
    m.add(np.array([[0, 0, 1], [1, 0, 1], [0, 1, 1]], np.float32))
    assert m.size() == 3
    assert m.match(st, np.array([[0.2, 0.2, 1.0]], np.float32)).n_valid == 0
    # NaN map points are skipped (Octree.hpp:243); NaN / far-away / out-of-bbox scan points are rejected
    rng = np.random.default_rng(0)
    plane = np.c_[rng.uniform(-4, 4, (5000, 2)), 2.0 + rng.normal(0, 0.004, 5000)].astype(np.float32)
    plane[::97] = np.nan
    m2 = mapper()
    m2.add(plane)
    keep = ~np.isnan(plane[:, 0])
    assert m2.size() == int(keep.sum())
    om = oracle.OracleMap()
    om.add(plane)
    assert om.size() == m2.size()
    scan = np.array([[0.1, 0.2, 2.2], [np.nan, 0, 0], [500, 500, 500], [-500, 0, 2], [3.99, 3.99, 2.1], [0, 0, 1e9]], np.float32)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG), st[:14], scan)
    m2.set_scan(scan)
    dbg = m2.match_debug(st)
    assert np.array_equal(dbg["good"], ref["good"]) and ref["good"].tolist()[:4] == [True, False, False, False]
    assert np.array_equal(dbg["dist"][ref["good"]], ref["dist"][ref["good"]])
    # empty scan and stride-32 (fast_limo::Point) input
    assert m2.match(st, np.zeros((0, 3), np.float32)).n_valid == 0
    pts32 = np.zeros((scan.shape[0], 8), np.float32)
    pts32[:, :3] = scan
    assert m2.match(st, pts32).n_valid == int(ref["good"].sum())
    m3 = mapper()
    map32 = np.zeros((plane.shape[0], 8), np.float32)
    map32[:, :3] = plane
    m3.add(map32)
    assert m3.size() == m2.size() and np.array_equal(m3.match(st, scan).HTH, m2.match(st, scan).HTH)


def test_dense_cells_and_escalation(oracle, flimo_lib):
    """Very dense patches (long runs -> team scans) and sparse ones (several index levels)."""
    rng = np.random.default_rng(3)
    dense = np.c_[rng.uniform(-0.4, 0.4, (20000, 2)), 3.0 + rng.normal(0, 0.003, 20000)]
    sparse = np.c_[rng.uniform(-30, 30, (3000, 2)), 3.0 + rng.normal(0, 0.003, 3000)]
    pts = np.r_[dense, sparse].astype(np.float32)
    scan = np.c_[rng.uniform(-25, 25, (4096, 2)), rng.uniform(2.8, 3.3, 4096)].astype(np.float32)
    scan[:512, :2] = rng.uniform(-0.5, 0.5, (512, 2))
    st = synth.make_state([0, 0, 0], [0, 0, 0, 1])
    om = oracle.OracleMap()
    om.add(pts)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG), st[:14], scan)
    assert 100 < ref["good"].sum() < 4096
    for cell in (0.0, 0.1, 1.0):
        m = mapper(knn_cell=cell)
        m.add(pts)
        m.set_scan(scan)
        check_per_point(m.match_debug(st), ref)


def test_sharded_sum_equals_whole(flimo_lib):
    import torch
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    whole = m.match(case.init)
    n = case.scan.shape[0]
    acc = np.zeros(96)
    buf = torch.zeros(96, dtype=torch.float64, device="cuda")
    for a, b in ((0, n // 3), (n // 3, n // 2), (n // 2, n)):
        m.shard(a, b)
        m.match_async(case.init, buf.data_ptr())
        torch.cuda.synchronize()
        acc += buf.cpu().numpy()
    part = api.unpack96(acc)
    assert part.n_valid == whole.n_valid
    assert np.allclose(part.HTH, whole.HTH, rtol=1e-12, atol=1e-12 * np.abs(whole.HTH).max())
    m.shard(0, n)
    assert np.array_equal(m.match(case.init).HTH, whole.HTH)          # deterministic reduction


def test_headline_config_properties(oracle, flimo_lib):
    """BASELINE configs[1] at full size (131 072-pt scan, 5 M-pt map): parity of the pass with the
    oracle, run-to-run determinism, order independence, and registration onto the true pose."""
    case = synth.make_case("c2")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=oracle.max_threads()), case.init[:14], case.scan)
    check_per_point(m.match_debug(case.init), ref)
    r1, r2 = m.match(case.init), m.match(case.init)
    assert np.array_equal(r1.HTH, r2.HTH) and r1.n_valid == ref["n_valid"]
    assert np.allclose(r1.HTH, ref["HTH"], rtol=1e-12, atol=1e-12 * np.abs(ref["HTH"]).max())
    perm = np.random.default_rng(0).permutation(case.scan.shape[0])
    m.set_scan(case.scan[perm])
    r3 = m.match(case.init)
    assert r3.n_valid == r1.n_valid and np.allclose(r3.HTH, r1.HTH, rtol=1e-11, atol=1e-11 * np.abs(r1.HTH).max())
    m.set_scan(case.scan)
    x, P, passes = m.update(case.init, synth.default_P0(), 2, 0.0)
    assert passes == 3 and np.abs(x[:3] - case.truth[:3]).max() < 5e-3
    w = m.scan_to_world(x)
    assert w.shape == case.scan.shape and np.isfinite(w).all()


def _batches(seed, n_batches=10, n=6000):
    """A vehicle driving along +x over a noisy ground plane with a wall: dense re-observation (drops),
    new terrain (root doubling in several directions) and a few NaN points."""
    rng = np.random.default_rng(seed)
    out = []
    for b in range(n_batches):
        cx = 6.0 * b - (25.0 if b == 7 else 0.0)          # one jump backwards: expansion towards -x
        ground = np.c_[rng.uniform(cx - 12, cx + 12, n), rng.uniform(-10, 10 + b, n), 0.3 + rng.normal(0, 0.01, n)]
        wall = np.c_[rng.uniform(cx - 12, cx + 12, n // 3), np.full(n // 3, 10.0 + b) + rng.normal(0, 0.01, n // 3),
                     rng.uniform(0.3, 4.0 + 3 * b, n // 3)]
        pts = np.r_[ground, wall].astype(np.float32)
        pts[rng.integers(0, len(pts), 5)] = np.nan
        out.append(pts)
    return out


@pytest.mark.parametrize("downsample", [True, False])
@pytest.mark.parametrize("min_extent", [0.2, 0.35])
def test_incremental_insert_matches_octree(oracle, flimo_lib, downsample, min_extent):
    """Mapper::add / Octree::update (Octree.hpp:341-432): identical map CONTENTS after every batch."""
    om = oracle.OracleMap(min_extent=min_extent, downsample=downsample)
    m = mapper(octree_downsampling=downsample, octree_min_extent=min_extent)
    prev = 0
    for b, pts in enumerate(_batches(5)):
        om.add(pts)
        m.add(pts, float(b))
        assert m.size() == om.size(), (b, m.size(), om.size())
        assert m.last_time() == float(b)
        prev = m.size()
    a = m.points()
    r = om.points()
    key = lambda p: p[np.lexsort((p[:, 2], p[:, 1], p[:, 0]))]
    assert np.array_equal(key(a), key(r))
    if downsample:
        assert m.size() < sum(int((~np.isnan(p[:, 0])).sum()) for p in _batches(5))
    # and the grown map still answers queries like the octree
    st = synth.make_state([20.0, 0.0, 1.8], [0, 0, 0, 1])
    scan = np.c_[np.random.default_rng(1).uniform(-10, 10, (2000, 2)), np.full(2000, -1.5)].astype(np.float32)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG), st[:14], scan)
    m.set_scan(scan)
    check_per_point(m.match_debug(st), ref)
    assert ref["good"].sum() > 500


def test_insert_tiny_first_scan(oracle, flimo_lib):
    """First scan smaller than a min-level cell (root itself is the min level) and an all-NaN batch."""
    rng = np.random.default_rng(2)
    om = oracle.OracleMap()
    m = mapper()
    first = rng.uniform(-0.15, 0.15, (20, 3)).astype(np.float32)
    for pts in (first, first + 0.01, rng.uniform(-3, 3, (500, 3)).astype(np.float32), first - 0.02):
        om.add(pts)
        m.add(pts)
        assert m.size() == om.size()
    m.add(np.full((4, 3), np.nan, np.float32))
    assert m.size() == om.size()


def _xch_worker(rank, world, name, out):
    from multiprocessing import shared_memory
    from fast_limo_b200.dist import shard_bounds
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    shm = shared_memory.SharedMemory(name=name)
    m.exchange_attach(shm.buf, rank, world)
    m.set_scan(case.scan)
    m.shard(*shard_bounds(case.scan.shape[0], rank, world))
    res = []
    for rep in range(3):                                   # several updates: sequence numbers keep advancing
        x, P, passes = m.update_exchange(case.init, synth.default_P0(), 2, 0.0)
        res.append((x, P, passes))
    out[rank] = res
    m.close()
    shm.close()


@pytest.mark.timeout(600)
def test_fused_exchange_two_processes(oracle, flimo_lib):
    """Scan sharded over two PROCESSES (sharing cuda:0 here), partial sums exchanged through the mapped
    host segment written by the kernels' last CTAs: every rank must end in the identical state, equal to
    the single-process update."""
    import torch.multiprocessing as mp
    from multiprocessing import shared_memory
    world = 2
    shm = shared_memory.SharedMemory(create=True, size=2 * 4096)
    shm.buf[:2 * 4096] = bytes(2 * 4096)
    try:
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_xch_worker, args=(world, shm.name, out), nprocs=world, join=True)
    finally:
        shm.close()
        shm.unlink()
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    x1, P1, p1 = m.update(case.init, synth.default_P0(), 2, 0.0)
    for rep in range(3):
        (xa, Pa, pa), (xb, Pb, pb) = out[0][rep], out[1][rep]
        assert pa == pb == p1 == 3
        assert np.array_equal(xa, xb) and np.array_equal(Pa, Pb)
        assert np.abs(xa - x1).max() <= 1e-10 and np.allclose(Pa, P1, rtol=1e-4, atol=1e-11)


def _stall_worker(stall, out):
    os.environ["FLIMO_TEST_STALL_PASS"] = str(stall)
    os.environ["FLIMO_DEVICE_EKF"] = "0"                   # the host-driven form: persistent kernel + handshake
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    res = []
    for rep in range(3):
        x, P, passes = m.update(case.init, synth.default_P0(), 2, 0.0)
        res.append((x, P, passes))
    out[stall] = res
    m.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("stall", [0, 1, 2])
def test_persistent_kernel_watchdog_fallback(flimo_lib, stall):
    """The host goes silent for 60 ms before a pass (longer than the persistent kernel's 20 ms watchdog):
    the kernel must end by itself, the update must continue with per-pass launches and give the identical
    result — also on the following updates (command / ticket state stays consistent)."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    p = mp.get_context("spawn").Process(target=_stall_worker, args=(stall, out))
    p.start()
    p.join(300)
    assert p.exitcode == 0
    case = synth.make_case("tiny")
    os.environ["FLIMO_DEVICE_EKF"] = "0"
    try:
        m = mapper()
    finally:
        del os.environ["FLIMO_DEVICE_EKF"]
    m.add(case.map_pts)
    m.set_scan(case.scan)
    x1, P1, p1 = m.update(case.init, synth.default_P0(), 2, 0.0)
    for (x, P, passes) in out[stall]:
        assert passes == p1 == 3
        assert np.array_equal(x, x1) and np.array_equal(P, P1)


def test_incremental_index_equals_rebuild(oracle, flimo_lib):
    """Mapper::add merges the accepted points into the search index (sort of the new entries + one merge pass
    + table shift).  The result must be indistinguishable from rebuilding the index from scratch: same per-point
    answers bit for bit, after batches inside the grid box, a batch that enlarges the box (full rebuild) and
    more batches after it — and equal to the oracle."""
    case = synth.make_case("c1")
    rng = np.random.default_rng(5)
    base = case.map_pts[:60000]
    batches = [case.map_pts[60000 + 4000 * i: 60000 + 4000 * (i + 1)] for i in range(4)]
    far = (case.map_pts[:3000] + np.array([400.0, -250.0, 3.0], np.float32)).astype(np.float32)     # outside the padded box
    batches = batches[:2] + [far] + batches[2:] + [case.map_pts[80000:80050]]
    os.environ["FLIMO_INDEX_INCREMENTAL"] = "0"
    try:
        m_full = mapper()
    finally:
        del os.environ["FLIMO_INDEX_INCREMENTAL"]
    m_inc = mapper()
    om = oracle.OracleMap()
    ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4)
    for m in (m_full, m_inc):
        m.add(base, 0.0)
        m.set_scan(case.scan)
    om.add(base)
    for k, b in enumerate(batches):
        for m in (m_full, m_inc):
            m.add(b, float(k + 1))
        om.add(b)
        assert m_full.size() == m_inc.size() == om.size()
        a, c = m_full.match_debug(case.init), m_inc.match_debug(case.init)
        for key in ("world", "good", "plane", "dist", "nn_d2"):
            assert np.array_equal(a[key], c[key]), (k, key)
        ra, rc = m_full.match(case.init), m_inc.match(case.init)
        assert np.array_equal(ra.HTH, rc.HTH) and ra.n_valid == rc.n_valid
    ref = om.match(ocfg, case.init[:14], case.scan)
    check_per_point(m_inc.match_debug(case.init), ref)


def test_incremental_index_growth_stress(oracle, flimo_lib):
    """A map that grows by a factor of 20 in small batches wandering through the grid box: rows outgrow their segments and
    move to the free tail, the tail runs out (rebuild) — after every few batches the answers must equal those of an index
    rebuilt from scratch at every add, and at the end the oracle's."""
    case = synth.make_case("c1")
    order = np.argsort(case.map_pts[:, 0] + 0.3 * case.map_pts[:, 1], kind="stable")       # sweep through the world
    pts = case.map_pts[order]
    corners = np.float32([[-45, -45, -3], [45, 45, 28]])                                    # fixes the grid box up front
    base = np.concatenate([corners, pts[::20]])                                            # a thin sample everywhere
    rest = np.delete(pts, np.arange(0, len(pts), 20), axis=0)
    os.environ["FLIMO_INDEX_INCREMENTAL"] = "0"
    try:
        m_full = mapper(octree_downsampling=False)
    finally:
        del os.environ["FLIMO_INDEX_INCREMENTAL"]
    m_inc = mapper(octree_downsampling=False)
    om = oracle.OracleMap(downsample=False)
    for m in (m_full, m_inc):
        m.add(base, 0.0)
        m.set_scan(case.scan)
    om.add(base)
    n_b = 48
    step = len(rest) // n_b
    for k in range(n_b):
        b = rest[k * step: (k + 1) * step] if k + 1 < n_b else rest[k * step:]
        for m in (m_full, m_inc):
            m.add(b, float(k + 1))
        om.add(b)
        if k % 6 == 5 or k + 1 == n_b:
            assert m_full.size() == m_inc.size() == om.size()
            a, c = m_full.match_debug(case.init), m_inc.match_debug(case.init)
            for key in ("good", "plane", "dist", "nn_d2"):
                assert np.array_equal(a[key], c[key]), (k, key)
    st = m_inc.stats()
    assert st["index_updates"] >= n_b // 2 and st["index_rows_moved"] > 0, st       # the incremental path ran and rows did move
    ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4)
    check_per_point(m_inc.match_debug(case.init), om.match(ocfg, case.init[:14], case.scan))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_against_reference_octree_golden(flimo_lib, tag):
    """The CUDA path against vectors produced by the REFERENCE's own octree (tests/golden/ref_octree_*.npz,
    generated from oracle/_ref = Octree.hpp compiled unmodified): map size after every Mapper::add, final
    contents, and the exact five neighbour distances of 3 500 queries (identity pose: world point = scan point)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(G, f"ref_octree_{tag}.npz"))
    batches, queries = mg.ref_octree_inputs()
    m = mapper(octree_min_extent=float(g["min_extent"]), octree_downsampling=bool(g["downsample"]))
    sizes = []
    for k, b in enumerate(batches):
        m.add(b, float(k))
        sizes.append(m.size())
    assert sizes == g["sizes"].tolist()
    assert np.array_equal(mg.contents_checksum(m.points()), g["contents_checksum"])
    m.set_scan(queries)
    ident = synth.make_state([0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0])
    dbg = m.match_debug(ident)
    assert np.array_equal(dbg["world"], queries)
    close = g["knn_d2"][:, 4] < 2.0                 # beyond MAX_DIST_PLANE the device search may stop early (outcome-equivalent)
    assert close.sum() > 500
    assert np.array_equal(dbg["nn_d2"][close], g["knn_d2"][close])
