"""GPU (-m gpu): the update that stays on the device (registration kernel, csrc/match_kernel.cu + csrc/ekf_step.hpp).

One launch per flimo_update: the CTA that completes a measurement pass runs the filter step (esekfom.hpp:1652-1819)
and hands the next pose to the other CTAs; the first-N rule of MAX_NUM_MATCHES (Localizer.cpp:539,547-548) is resolved
by a prefix count over the accepted-match bits; with several ranks the pass sums travel over peer memory.
Tolerances: per-pass state <= 1e-9 vs the oracle (device libm differs from glibc in the last bits of sin/cos/atan);
covariance rtol 1e-4 (P = L - K P cancels ~5 digits), the same bars as tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BIG = 1 << 18


def mapper(env=None, **kw):
    kw.setdefault("MAX_NUM_MATCHES", BIG)
    kw.setdefault("MAX_NUM_PC2MATCH", BIG)
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        return api.Mapper(api.MappingConfig(**kw), device=0)
    finally:
        for k, v in old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


@pytest.mark.parametrize("max_iter,lim", [(0, 0.001), (3, 0.0), (4, 1e9), (2, 0.0)])
def test_per_pass_states_match_oracle(oracle, flimo_lib, max_iter, lim):
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=2)
    xo, Po, tr = om.update(ocfg, case.init, synth.default_P0(), max_iter, lim, case.scan)
    x, P, passes = m.update(case.init, synth.default_P0(), max_iter, lim)
    rec = m.update_trace()
    assert passes == len(tr) == len(rec)
    for k in range(passes):
        assert np.abs(rec[k, :26] - tr[k]["state"]).max() <= 1e-9
        assert int(rec[k, 27]) == tr[k]["rows"]
    assert np.abs(x - xo).max() <= 1e-9 and np.allclose(P, Po, rtol=1e-4, atol=1e-11)
    assert m.stats()["kernel_launches"] >= 1


def test_device_step_equals_host_step(flimo_lib):
    """Same scan, same map: the device-resident update and the host-driven one (FLIMO_DEVICE_EKF=0: persistent kernel +
    host filter step) agree to round-off, over repeated updates with the carried covariance."""
    case = synth.make_case("tiny")
    md, mh = mapper(), mapper(env={"FLIMO_DEVICE_EKF": "0"})
    for m in (md, mh):
        m.add(case.map_pts)
        m.set_scan(case.scan)
    xd, Pd = case.init.copy(), synth.default_P0()
    xh, Ph = case.init.copy(), synth.default_P0()
    for rep in range(4):
        xd, Pd, pd = md.update(xd, Pd, 3, 0.001)
        xh, Ph, ph = mh.update(xh, Ph, 3, 0.001)
        assert pd == ph
        assert np.abs(xd - xh).max() <= 1e-11 and np.allclose(Pd, Ph, rtol=1e-6, atol=1e-13)
        Pd, Ph = Pd + 1e-6 * np.eye(23), Ph + 1e-6 * np.eye(23)      # stand-in for the process noise between scans


@pytest.mark.parametrize("pc2,mm", [(1500, 400), (BIG, 1000), (BIG, 1)])
def test_first_n_rule_on_device(oracle, flimo_lib, pc2, mm):
    """More accepted matches than MAX_NUM_MATCHES: every pass is repeated once with the row limit found on the device;
    rows, per-pass states and the result equal the oracle's (and the golden fixture for the 1500 / 400 case)."""
    case = synth.make_case("tiny")
    m = mapper(MAX_NUM_PC2MATCH=pc2, MAX_NUM_MATCHES=mm)
    m.add(case.map_pts)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ocfg = oracle.make_cfg(max_pc2match=pc2, max_matches=mm, num_threads=2)
    xo, Po, tr = om.update(ocfg, case.init, synth.default_P0(), 3, 0.0, case.scan)
    launches0 = m.stats()["kernel_launches"]
    x, P, passes = m.update(case.init, synth.default_P0(), 3, 0.0)
    assert m.stats()["kernel_launches"] == launches0 + 2              # tiles + filter CTA: two launches, repeated passes included
    rec = m.update_trace()
    assert passes == len(tr) == 4
    for k in range(passes):
        assert int(rec[k, 27]) == tr[k]["rows"] == mm and int(rec[k, 26]) > mm
        assert rec[k, 29] < min(len(case.scan), pc2)                  # the pass ran with a row limit
        assert np.abs(rec[k, :26] - tr[k]["state"]).max() <= 1e-9
    assert np.abs(x - xo).max() <= 1e-9 and np.allclose(P, Po, rtol=1e-4, atol=1e-11)
    if (pc2, mm) == (1500, 400):
        g = np.load(os.path.join(G, "tiny_m400.npz"))
        assert np.allclose(x, g["x_final"], atol=1e-9) and np.allclose(P, g["P_final"], rtol=1e-4, atol=1e-11)


def test_non_finite_covariance_is_reported(flimo_lib):
    case = synth.make_case("tiny")
    m = mapper()
    m.add(case.map_pts)
    m.set_scan(case.scan)
    P = synth.default_P0()
    P[2, 2] = np.nan
    with pytest.raises(api.FlimoError, match="singular"):
        m.update(case.init, P, 2, 0.0)
    x, P2, passes = m.update(case.init, synth.default_P0(), 2, 0.0)    # the handle stays usable
    assert passes == 3 and np.abs(x[:3] - case.truth[:3]).max() < 0.01


def test_empty_map_and_many_tiles(oracle, flimo_lib):
    """Zero matches against an empty map (the second scan of a run): the update runs host-side with zero rows; and a scan
    with more tiles than resident CTAs (every CTA loops over several tiles inside the registration kernel)."""
    case = synth.make_case("tiny")
    m = mapper()
    m.set_scan(case.scan)
    x, P, passes = m.update(case.init, synth.default_P0(), 2, 0.0)
    assert passes == 2 and np.array_equal(x, case.init)      # zero rows: dx = 0 converges at once, two converged passes end the loop
    c1 = synth.make_case("c1")
    big_scan = np.concatenate([c1.scan] * 12)[:, :3].copy()           # 196 608 points = 1 536 tiles > 1 036 resident CTAs
    m.add(c1.map_pts)
    m.set_scan(big_scan)
    om = oracle.OracleMap()
    om.add(c1.map_pts)
    xo, Po, tr = om.update(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), c1.init, synth.default_P0(), 1, 0.0, big_scan)
    x, P, passes = m.update(c1.init, synth.default_P0(), 1, 0.0)
    assert passes == len(tr) == 2 and np.abs(x - xo).max() <= 1e-9


def _peer_worker(rank, world, box, out, mm):
    import time
    from fast_limo_b200.dist import shard_bounds
    case = synth.make_case("tiny")
    m = mapper(MAX_NUM_PC2MATCH=1500 if mm == 400 else BIG, MAX_NUM_MATCHES=mm)
    m.add(case.map_pts)
    box[rank] = m.peer_export()
    t0 = time.time()
    while len(box) < world:
        time.sleep(0.01)
        assert time.time() - t0 < 120
    m.peer_attach(rank, world, [box[r] for r in range(world)])
    m.set_scan(case.scan)
    n = min(len(case.scan), 1500 if mm == 400 else BIG)
    m.shard(*shard_bounds(n, rank, world))
    res = []
    for rep in range(3):
        x, P, passes = m.update_peer(case.init, synth.default_P0(), 3, 0.0)
        res.append((x, P, passes, m.update_trace()[:, 26:30].copy()))
    out[rank] = res
    m.close()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mm", [BIG, 400])
def test_peer_exchange_two_processes(flimo_lib, mm):
    """Scan sharded over two PROCESSES (sharing cuda:0 here, so their kernels time-slice), pass sums — and, with
    MAX_NUM_MATCHES = 400, the accepted-match bits for the first-N rule across shards — exchanged through each
    other's device inbox (CUDA IPC): both ranks end in the identical state, equal to the single-GPU update and to the
    golden fixture."""
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    box, out = mgr.dict(), mgr.dict()
    mp.spawn(_peer_worker, args=(world, box, out, mm), nprocs=world, join=True)
    case = synth.make_case("tiny")
    m = mapper(MAX_NUM_PC2MATCH=1500 if mm == 400 else BIG, MAX_NUM_MATCHES=mm)
    m.add(case.map_pts)
    m.set_scan(case.scan)
    x1, P1, p1 = m.update(case.init, synth.default_P0(), 3, 0.0)
    for rep in range(3):
        (xa, Pa, pa, ta), (xb, Pb, pb, tb) = out[0][rep], out[1][rep]
        assert pa == pb == p1 == 4
        assert np.array_equal(xa, xb) and np.array_equal(Pa, Pb) and np.array_equal(ta[:, :2], tb[:, :2])
        assert np.abs(xa - x1).max() <= 1e-10 and np.allclose(Pa, P1, rtol=1e-4, atol=1e-11)
        if mm == 400:
            assert np.all(ta[:, 1] == 400) and np.all(ta[:, 0] > 400)
    if mm == 400:
        g = np.load(os.path.join(G, "tiny_m400.npz"))
        assert np.allclose(out[0][0][0], g["x_final"], atol=1e-9)


def test_stalled_resident_update_is_redone_per_pass(oracle, flimo_lib):
    """The resident kernels of an update wait for each other (tiles for the next pose, the filter CTA for the pass sums); a
    watchdog ends a stalled pair and flimo_update then redoes the update with one launch per pass.  Fault injection
    (FLIMO_DEBUG_STALL_EVERY): every third update's tiles go deaf after their first pass.  Results must not change."""
    case = synth.make_case("tiny")
    P0 = synth.default_P0()
    m = mapper()
    m.add(case.map_pts, 0.0)
    m.set_scan(case.scan)
    x_ref, P_ref, p_ref = m.update(case.init, P0, 2, 0.0)
    mf = mapper(env={"FLIMO_DEBUG_STALL_EVERY": "3", "FLIMO_TIME_EVERY": "0"})
    mf.add(case.map_pts, 0.0)
    mf.set_scan(case.scan)
    for k in range(7):
        x, P, passes = mf.update(case.init, P0, 2, 0.0)
        assert passes == p_ref == 3
        assert np.abs(x - x_ref).max() <= 1e-9 and np.allclose(P, P_ref, rtol=1e-4, atol=1e-11), k      # device vs host filter step: libm last bits
    assert mf.stats()["update_stalls"] == 2 and m.stats()["update_stalls"] == 0
