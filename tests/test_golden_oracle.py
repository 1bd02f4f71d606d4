"""CPU: the oracle reproduces the committed golden fixtures (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from fast_limo_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("tiny", 1 << 18, 1 << 18), ("tiny", 1500, 400), ("c1", 1 << 18, 1 << 18)]


def load(name, mm):
    return np.load(os.path.join(G, f"{name}_m{mm}.npz"))


@pytest.mark.parametrize("name,pc2,mm", CASES)
def test_oracle_matches_golden(oracle, name, pc2, mm):
    g = load(name, mm)
    case = synth.make_case(name)
    assert case.scan.astype(np.float64).sum() == float(g["scan_checksum"])      # generators are seed-stable
    assert case.map_pts.astype(np.float64).sum() == float(g["map_checksum"])
    om = oracle.OracleMap()
    om.add(case.map_pts)
    cfg = oracle.make_cfg(max_pc2match=pc2, max_matches=mm, num_threads=2)
    r = om.match(cfg, case.init[:14], case.scan)
    good = np.unpackbits(g["good"])[: int(g["n_q"])].astype(bool)
    assert np.array_equal(r["good"], good)
    assert np.array_equal(r["plane"][good], g["plane"]) and np.array_equal(r["dist"][good], g["dist"])
    assert np.array_equal(r["nn_d2"], g["nn_d2"])
    assert r["n_valid"] == int(g["n_valid"]) and r["rows"] == int(g["rows"])
    assert np.allclose(r["HTH"], g["HTH"], rtol=1e-13, atol=1e-12)
    x, P, tr = om.update(cfg, case.init, synth.default_P0(), int(g["max_iter"]), 0.0, case.scan)
    assert np.allclose(x, g["x_final"], rtol=0, atol=1e-12)
    assert np.allclose(P, g["P_final"], rtol=1e-9, atol=1e-13)
    assert [t["rows"] for t in tr] == g["trace_rows"].tolist()


def _prep_inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.prep_inputs()


def test_oracle_prep_matches_golden(oracle):
    """Scan preparation (filters, time sort, deskew, voxel grid) against tests/golden/prep_velodyne.npz."""
    O = oracle
    g = np.load(os.path.join(G, "prep_velodyne.npz"))
    raw, frames, T = _prep_inputs()
    assert np.nansum(raw["x"].astype(np.float64)) == float(g["raw_checksum"])
    cfg = O.make_prep_cfg(crop=([-1.5, -1.0, -1.0], [1.5, 1.0, 1.0]), min_dist=4.0, rate=2, fov=2.6, sensor_type=1, leaf=1.0)
    order = O.prep_filter_sort(raw, cfg, sort=True)
    assert np.array_equal(order, g["order"])
    assert O.prep_times(raw, order, cfg, 100.0)[-1] == float(g["t_last"])
    w, b = O.prep_deskew(raw, order, cfg, 100.0, -2.0e-4, frames, frames["q"][-2], frames["p"][-2], T)
    # sinf / cosf come from libm: allow an ulp-level difference between machines
    assert np.abs(w - g["world"]).max() <= 2e-5 and np.abs(b - g["xt2"]).max() <= 2e-5
    assert np.array_equal(O.prep_voxel(g["xt2"], 1.0), g["voxel"])


def _golden_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_imu_predict_matches_golden(oracle, flimo_lib):
    """esekf::predict x 120 + the frames of one scan window: oracle AND product (host algebra of libflimo_cuda)
    against tests/golden/imu_predict.npz."""
    from fast_limo_b200 import api
    mod = _golden_module()
    g = np.load(os.path.join(G, "imu_predict.npz"))
    x0, P0, stamps, dts, acc, gyr = mod.imu_inputs()
    pr = oracle.Propagator(x0, P0)
    m = api.Mapper(device=-1)
    x, P = x0, P0
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        pr.propagate(t, dt, a, w, mod.IMU_COV)
        x, P = m.ekf_predict(x, P, t, dt, a, w, mod.IMU_COV)
    gold_fr = np.frombuffer(g["frames"].tobytes(), oracle.FRAME)
    assert len(gold_fr) == int(g["n_frames"]) == 42
    for (xx, PP, fr) in ((*pr.get(), pr.frames(*mod.IMU_WINDOW)), (x, P, m.propagated_frames(*mod.IMU_WINDOW))):
        # sin / cos / atan come from libm: allow last-bit differences between machines
        assert np.allclose(xx, g["x"], rtol=0, atol=1e-11)
        assert np.allclose(PP, g["P"], rtol=1e-10, atol=1e-16)
        assert np.array_equal(fr["time"], gold_fr["time"]) and np.array_equal(fr["a"], gold_fr["a"]) and np.array_equal(fr["w"], gold_fr["w"])
        for k in ("q", "p", "v", "bg", "ba", "g"):
            assert np.allclose(fr[k], gold_fr[k], rtol=3e-7, atol=1e-9)
