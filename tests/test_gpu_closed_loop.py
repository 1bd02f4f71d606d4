"""GPU checks of the two-lanes-per-query scan variant and of the closed loop (promoted into `-m gpu` in round 2 after
their first run on a B200: 8 / 8 green).

FLIMO_KNN_PAIR=1 -- two lanes per query in the first scan (csrc/match_kernel.cu, pair_scan_round): tiny runs, dense cells,
caps, empty blocks, and the iterated update, all against the oracle.

fast_limo_b200.localizer -- the host mirror of Localizer::updateIMU / updatePointCloud in closed loop (IMU samples ->
prediction -> propagated frames -> device deskew -> update -> map add)."""
import os

import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu
BIG = 1 << 18


def pair_mapper(**kw):
    kw.setdefault("MAX_NUM_MATCHES", BIG)
    kw.setdefault("MAX_NUM_PC2MATCH", BIG)
    os.environ["FLIMO_KNN_PAIR"] = "1"
    try:
        return api.Mapper(api.MappingConfig(**kw), device=0)
    finally:
        del os.environ["FLIMO_KNN_PAIR"]


@pytest.mark.parametrize("name", ["tiny", "c1"])
@pytest.mark.parametrize("cell", [0.0, 0.15, 0.6])
def test_pair_scan_per_point_parity(oracle, flimo_lib, name, cell):
    case = synth.make_case(name)
    m = pair_mapper(knn_cell=cell)
    m.add(case.map_pts, 0.0)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), case.init[:14], case.scan)
    dbg = m.match_debug(case.init)
    assert np.array_equal(dbg["good"], ref["good"])
    g = ref["good"]
    assert np.array_equal(dbg["plane"][g], ref["plane"][g]) and np.array_equal(dbg["dist"][g], ref["dist"][g])
    close = ref["nn_d2"][:, -1] < 2.0
    assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])
    r = m.match(case.init)
    assert r.n_valid == ref["n_valid"] and np.allclose(r.HTH, ref["HTH"], rtol=1e-12, atol=1e-9)


def test_pair_scan_dense_sparse_and_update(oracle, flimo_lib):
    """Very dense cells (runs beyond the private cap), an almost empty map, and the iterated update."""
    rng = np.random.default_rng(1)
    dense = (rng.normal(0, [0.3, 0.3, 0.01], (60000, 3))).astype(np.float32)
    sparse = rng.uniform(-30, 30, (40, 3)).astype(np.float32)
    scan = np.concatenate([rng.normal(0, [0.5, 0.5, 0.02], (3000, 3)), rng.uniform(-30, 30, (500, 3))]).astype(np.float32)
    ident = synth.make_state([0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0])
    for pts in (dense, sparse, np.concatenate([dense, sparse])):
        m = pair_mapper()
        m.add(pts, 0.0)
        m.set_scan(scan)
        om = oracle.OracleMap()
        om.add(pts)
        ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), ident[:14], scan)
        dbg = m.match_debug(ident)
        assert np.array_equal(dbg["good"], ref["good"])
        close = ref["nn_d2"][:, -1] < 2.0
        assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])
    case = synth.make_case("tiny")
    m, m0 = pair_mapper(), api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    for mm in (m, m0):
        mm.add(case.map_pts, 0.0)
        mm.set_scan(case.scan)
    xa, Pa, pa = m.update(case.init, synth.default_P0(), 2, 0.0)
    xb, Pb, pb = m0.update(case.init, synth.default_P0(), 2, 0.0)
    assert pa == pb and np.array_equal(xa, xb) and np.array_equal(Pa, Pb)


def test_localizer_closed_loop(flimo_lib):
    """Twelve sweeps of the synthetic stream through the two callbacks: the first scan is dropped (no propagated state
    before it), the second initialises the map with zero matches, the following ones are registered."""
    from fast_limo_b200.localizer import Localizer, LocalizerConfig
    S = synth.Stream(rings=64, azimuths=512, imu_hz=200.0)    # same stream as tests/test_localizer_sequence.py (CPU, oracle stages)
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    filt = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=0.5, sensor_type=1)
    x0 = S.state(0.0)
    loc = Localizer(m, LocalizerConfig(filters=filt, MAX_NUM_ITERS=3), pos=x0[0:3], quat=x0[3:7], vel=x0[14:17])
    t_imu, outcome, errs = 0.0, [], []
    for k in range(12):
        raw, stamp = S.scan(k)
        t_need = stamp + S.dt + 1.0 / S.imu_hz               # the newest propagated state must not be older than the last point
        for smp in zip(*S.imu(t_imu, t_need, sigma_acc=0.05, sigma_gyro=0.002)):
            loc.updateIMU(smp[0], smp[1], smp[2] + np.float32([0.3, 0.0, 0.0]), smp[3])   # uncalibrated accelerometer bias
        t_imu = t_need
        outcome.append(loc.updatePointCloud(raw, stamp))
        truth = S.state(loc.imu_stamp)
        errs.append(float(np.linalg.norm(loc.x[0:3] - truth[0:3])))
        if k == 0:
            assert loc.last["null"].startswith("no frames") and m.size() == 0
        if k == 1:
            assert loc.last["passes"] >= 1 and m.size() > 0   # zero matches against the empty map, then Mapper::add
    assert outcome == [False] + [True] * 11
    assert max(errs) < 0.03, errs                             # dead reckoning alone drifts 0.5 * 0.3 * t^2 (22 cm at 1.2 s)
    assert m.size() > 20000


def test_map_receives_whole_cloud_beyond_pc2match_cap(oracle, flimo_lib):
    """MAX_NUM_PC2MATCH caps what Mapper::match queries (Mapper.cpp:63-69); the reference still transforms and maps the
    WHOLE pc2match (Localizer.cpp:361,377).  A 6144-point scan with a cap of 1500: the map must grow like the oracle's."""
    case = synth.make_case("tiny")
    assert len(case.scan) > 1500
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=1500), device=0)
    m.add(case.map_pts, 0.0)
    m.set_scan(case.scan)
    r = m.match(case.init)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ref = om.match(oracle.make_cfg(max_pc2match=1500, max_matches=BIG, num_threads=2), case.init[:14], case.scan)
    assert r.n_valid == ref["n_valid"]                         # only the first 1500 points were queried
    world = m.scan_to_world(case.truth)
    assert world.shape[0] == len(case.scan)                    # ... but all of them are transformed
    R = synth.quat_to_R(case.truth[3:7].astype(np.float32)).astype(np.float32)
    assert np.abs(world - (case.scan[:, :3] @ R.T + case.truth[0:3].astype(np.float32))).max() < 1e-4
    n0 = m.size()
    m.add_scan(case.truth, 1.0)
    om.add(world)
    assert m.size() == om.size() > n0
    assert np.array_equal(np.sort(m.points().view([("", np.float32)] * 3), axis=0), np.sort(om.points().view([("", np.float32)] * 3), axis=0))


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_tma_staged_scan_variant_parity(oracle, flimo_lib, name):
    """FLIMO_KNN_STAGE=1: the runs are staged in shared memory with cp.async.bulk + mbarrier (1-D TMA, UBLKCP in SASS) before
    they are scanned.  Slower than streaming on c2 (96.6 vs 35.5 us, profiles/README.md) and therefore off by default, but it
    has to stay exact: same accepted set, planes, distances and sums as the oracle."""
    case = synth.make_case(name)
    os.environ["FLIMO_KNN_STAGE"] = "1"
    try:
        m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    finally:
        del os.environ["FLIMO_KNN_STAGE"]
    m.add(case.map_pts, 0.0)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), case.init[:14], case.scan)
    dbg = m.match_debug(case.init)
    assert np.array_equal(dbg["good"], ref["good"])
    g = ref["good"]
    assert np.array_equal(dbg["plane"][g], ref["plane"][g]) and np.array_equal(dbg["dist"][g], ref["dist"][g])
    close = ref["nn_d2"][:, -1] < 2.0
    assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])
    r = m.match(case.init)
    assert r.n_valid == ref["n_valid"] and np.allclose(r.HTH, ref["HTH"], rtol=1e-12, atol=1e-9)
