"""GPU (-m gpu): device scan preparation (csrc/scan_prep.cu through the C ABI) against the CPU oracle.

Tolerances:
  filters + time sort : the kept indices and their order are IDENTICAL (points within 2e-6 rad of the
                        field-of-view limit are left out of the message: atan2f differs by an ulp between
                        CUDA and libm); point times bit-exact;
  deskew              : |delta| <= 2e-5 m per coordinate at ranges <= 150 m (sinf / cosf differ by an ulp
                        between CUDA and libm; everything else rounds like the oracle);
  voxel grid          : BIT-EXACT on identical input (stable sort + sequential per-voxel sums).
"""
import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu
BIG = 1 << 20


def mapper():
    return api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)


def message(n, sensor_type, seed=5, n_nan=60, fov=None):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-60, 60, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-2, 8, n)
    if fov is not None:      # keep clear of the FoV decision boundary
        ang = np.abs(np.arctan2(xyz[:, 1].astype(np.float64), xyz[:, 0].astype(np.float64)))
        xyz = xyz[np.abs(ang - fov) > 2e-6]
    return synth.make_raw_message(xyz, sensor_type=sensor_type, sweep=0.1, stamp=100.0, seed=seed, n_nan=n_nan)


def to_oracle_cfg(O, f: api.FilterConfig):
    crop = (f.cropBoxMin, f.cropBoxMax) if f.cropBoxMin is not None else None
    return O.make_prep_cfg(crop=crop, min_dist=f.min_dist, rate=f.rate_value, fov=f.fov_angle, sensor_type=f.sensor_type,
                           end_of_sweep=f.end_of_sweep, leaf=f.leafSize)


@pytest.mark.parametrize("sensor_type,eos", [(0, False), (0, True), (1, False), (1, True), (2, False), (3, False)])
def test_filter_sort_parity(oracle, flimo_lib, sensor_type, eos):
    O = oracle
    f = api.FilterConfig(cropBoxMin=(-1.5, -1.0, -1.0), cropBoxMax=(1.5, 1.0, 1.0), min_dist=4.0, rate_value=3, fov_angle=2.6,
                         sensor_type=sensor_type, end_of_sweep=eos)
    raw = message(50000, sensor_type, fov=2.6)
    m = mapper()
    n, t_last = m.prep_filter_sort(raw, 100.0, f)
    ocfg = to_oracle_cfg(O, f)
    order = O.prep_filter_sort(raw, ocfg, sort=True)
    assert n == len(order)
    assert np.array_equal(m.prep_get(0), order)
    assert t_last == O.prep_times(raw, order, ocfg, 100.0)[-1]
    # no filter at all, empty message
    n2, _ = m.prep_filter_sort(raw, 100.0, api.FilterConfig(sensor_type=sensor_type, end_of_sweep=eos))
    assert n2 == int((np.isfinite(raw["x"]) & np.isfinite(raw["y"]) & np.isfinite(raw["z"])).sum())
    assert m.prep_filter_sort(raw[:0], 100.0, f) == (0, 0.0)


@pytest.mark.parametrize("sensor_type", [0, 1, 2, 3])
def test_deskew_parity(oracle, flimo_lib, sensor_type):
    O = oracle
    f = api.FilterConfig(min_dist=2.0, sensor_type=sensor_type)
    raw = message(40000, sensor_type, seed=9)
    stamp = 100.0
    m = mapper()
    n, t_last = m.prep_filter_sort(raw, stamp, f)
    frames = synth.make_frames(stamp, stamp + 0.1, rate_hz=400.0, speed=12.0, yaw_rate=1.7)       # 12 m/s, ~100 deg/s
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.27, -0.03, 0.4]
    T[:3, :3] = synth.quat_to_R(synth.quat_from_rpy(0.01, -0.02, 0.03)).astype(np.float32)
    last_q, last_p = frames["q"][-2], frames["p"][-2]
    offset = -3.0e-4
    assert m.prep_deskew(frames, last_q, last_p, T, offset) == n
    ocfg = to_oracle_cfg(O, f)
    order = O.prep_filter_sort(raw, ocfg, sort=True)
    ow, ob = O.prep_deskew(raw, order, ocfg, stamp, offset, frames, last_q, last_p, T)
    gw, gb, pc = m.prep_get(1), m.prep_get(2), m.prep_get(3)
    assert gw.shape == ow.shape
    assert np.abs(gw - ow).max() <= 2e-5 and np.abs(gb - ob).max() <= 2e-5
    assert np.array_equal(pc, gb)                                  # no voxel grid: pc2match is the deskewed cloud
    # the deskewed cloud is bound as the scan: scan_to_world with the last pose gives the world cloud back
    st = synth.make_state(last_p.astype(np.float64), last_q.astype(np.float64))
    back = m.scan_to_world(st)
    assert np.abs(back - gw[:, :3]).max() <= 1e-4


def test_voxel_grid_bit_exact_and_pipeline(oracle, flimo_lib):
    O = oracle
    rng = np.random.default_rng(3)
    pts = np.ones((120000, 4), np.float32)
    pts[:, :3] = rng.uniform(-80, 80, (120000, 3))
    pts[:, 2] = rng.uniform(-2, 10, 120000)
    m = mapper()
    for leaf in (0.5, 1.0):
        assert np.array_equal(m.voxel_grid(pts, leaf), O.prep_voxel(pts, leaf))
    tiny = np.ones((4, 4), np.float32)
    tiny[:, :3] = [[0, 0, 0], [5000, 5000, 5000], [1, 2, 3], [-4000, 100, 7]]
    assert np.array_equal(m.voxel_grid(tiny, 0.001), tiny)          # "leaf size too small": input returned
    # full pipeline with voxel grid: same voxel population as the oracle up to deskew round-off
    f = api.FilterConfig(min_dist=3.0, rate_value=2, leafSize=1.0, sensor_type=1)
    raw = message(60000, 1, seed=21)
    n, _ = m.prep_filter_sort(raw, 100.0, f)
    frames = synth.make_frames(100.0, 100.1, rate_hz=200.0)
    last_q, last_p = frames["q"][-2], frames["p"][-2]
    nv = m.prep_deskew(frames, last_q, last_p, np.eye(4), 0.0)
    ocfg = to_oracle_cfg(O, f)
    order = O.prep_filter_sort(raw, ocfg, sort=True)
    _, ob = O.prep_deskew(raw, order, ocfg, 100.0, 0.0, frames, last_q, last_p, np.eye(4))
    ov = O.prep_voxel(ob, 1.0)
    assert abs(nv - len(ov)) <= max(3, len(ov) // 2000)            # a point within 1e-5 m of a voxel face may change voxel
    assert np.array_equal(m.voxel_grid(m.prep_get(2), 1.0), m.prep_get(3))   # stage consistency on the device cloud


def test_stream_replay_parity(oracle, flimo_lib):
    """BASELINE config c3 in small: a motion-distorted stream through the whole per-scan sequence of
    Localizer::updatePointCloud (filters -> deskew -> voxel grid -> iterated update -> world cloud ->
    Mapper::add with the down-sampling rule), device against oracle, scan after scan."""
    O = oracle
    S = synth.Stream(azimuths=256, rings=32)
    m = mapper()
    om = O.OracleMap()
    ocfg = O.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=2)
    f = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=0.5, sensor_type=1)
    opc = to_oracle_cfg(O, f)
    P0, lim = synth.default_P0(), np.full(23, 0.001)
    T = np.eye(4, dtype=np.float32)
    rng = np.random.default_rng(2)
    prev_end = 0.0
    for k in range(6):
        raw, stamp = S.scan(k)
        n, t_last = m.prep_filter_sort(raw, stamp, f)
        frames = S.frames(prev_end, t_last)
        pred = S.state(t_last)
        pred[:3] += rng.normal(0, 0.02, 3)
        lq, lp = pred[3:7].astype(np.float32), pred[:3].astype(np.float32)
        npc = m.prep_deskew(frames, lq, lp, T, 0.0)
        order = O.prep_filter_sort(raw, opc, sort=True)
        assert np.array_equal(m.prep_get(0), order)
        # the oracle continues from the DEVICE's pc2match (bit-identical voxel stage, deskew within 2e-5 m)
        pc = np.ascontiguousarray(m.prep_get(3)[:, :3])
        _, ob = O.prep_deskew(raw, order, opc, stamp, 0.0, frames, lq, lp, T)
        assert np.abs(m.prep_get(2) - ob).max() <= 2e-5
        if k == 0:
            xg = xo = pred.copy()
        else:
            xg, Pg, pg = m.update(pred, P0, 3, lim)
            xo, Po, tr = om.update(ocfg, pred, P0, 3, lim, pc)
            assert pg == len(tr)
            assert np.abs(xg[:3] - xo[:3]).max() <= 1e-9 and np.abs(xg[3:7] - xo[3:7]).max() <= 1e-10
        world = m.scan_to_world(xg)
        m.add(world, t_last)
        om.add(world)
        assert m.size() == om.size() and npc == len(pc)
        prev_end = t_last


def test_prep_against_golden_fixture(flimo_lib):
    """Device scan preparation against the committed fixture tests/golden/prep_velodyne.npz."""
    import importlib.util
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    raw, frames, T = mod.prep_inputs()
    g = np.load(os.path.join(G, "prep_velodyne.npz"))
    m = mapper()
    f = api.FilterConfig(cropBoxMin=(-1.5, -1.0, -1.0), cropBoxMax=(1.5, 1.0, 1.0), min_dist=4.0, rate_value=2, fov_angle=2.6,
                         leafSize=1.0, sensor_type=1)
    n, t_last = m.prep_filter_sort(raw, 100.0, f)
    assert n == len(g["order"]) and t_last == float(g["t_last"])
    assert np.array_equal(m.prep_get(0), g["order"])
    m.prep_deskew(frames, frames["q"][-2], frames["p"][-2], T, -2.0e-4)
    assert np.abs(m.prep_get(1) - g["world"]).max() <= 2e-5 and np.abs(m.prep_get(2) - g["xt2"]).max() <= 2e-5
    assert np.array_equal(m.voxel_grid(g["xt2"], 1.0), g["voxel"])


@pytest.mark.parametrize("sensor_type,dtype_code,field,np_type", [(0, 6, "t", "<u4"), (1, 7, "time", "<f4"), (2, 8, "timestamp", "<f8")])
def test_pointcloud2_wire_format_decode(oracle, flimo_lib, sensor_type, dtype_code, field, np_type):
    """sensor_msgs/PointCloud2 payloads with driver-style layouts (fields at other offsets, unaligned
    float64, padding, extra fields) decode on the device to the same cloud pcl::fromROSMsg builds on the host:
    identical filtered order and last-point time as the canonical 32-byte path."""
    raw = message(30000, sensor_type, seed=13)
    n = len(raw)
    # an Ouster / Velodyne / Hesai flavoured point: x y z | intensity | time | ring(u16) | pad
    step = {6: 26, 7: 22, 8: 30}[dtype_code] + 6           # deliberately odd point_step values
    wire = np.dtype({"names": ["x", "y", "z", "intensity", "tf", "ring"], "formats": ["<f4", "<f4", "<f4", "<f4", np_type, "<u2"],
                     "offsets": [0, 4, 8, 14, 18, 18 + np.dtype(np_type).itemsize], "itemsize": step})
    msg = np.zeros(n, wire)
    for k in ("x", "y", "z", "intensity"):
        msg[k] = raw[k]
    msg["tf"] = raw[field]
    msg["ring"] = np.arange(n) % 64
    f = api.FilterConfig(cropBoxMin=(-1.5, -1.0, -1.0), cropBoxMax=(1.5, 1.0, 1.0), min_dist=4.0, rate_value=3,
                         sensor_type=sensor_type)
    m = mapper()
    n1, t1 = m.prep_filter_sort(raw, 100.0, f)
    o1 = m.prep_get(0).copy()
    n2, t2 = m.prep_filter_sort_msg(msg.tobytes(), step, 100.0, f, off_xyz=(0, 4, 8), off_intensity=14, off_time=18,
                                    time_datatype=dtype_code)
    assert (n1, t1) == (n2, t2) and np.array_equal(m.prep_get(0), o1)
    order = oracle.prep_filter_sort(raw, to_oracle_cfg(oracle, f), sort=True)
    assert np.array_equal(o1, order)
    with pytest.raises(api.FlimoError, match="invalid pointcloud structure"):
        m.prep_filter_sort_msg(msg.tobytes(), step, 100.0, f, off_xyz=(0, 4, step - 2), off_time=18, time_datatype=dtype_code)
