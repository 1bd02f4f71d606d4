"""CPU: the oracle's scan preparation (oracle/prep.hpp: filters, time sort, deskew, voxel grid) against
independent numpy / scipy formulations of the same reference code (Localizer.cpp:262-321,733-853,
State.cpp:76-119).  The oracle is float32 in the reference's operation order; the cross-checks are
float64, so agreement is to float32 round-off."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from fast_limo_b200 import synth


def _message(n=6000, sensor_type=1, seed=5, n_nan=40):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-40, 40, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-2, 6, n)
    return synth.make_raw_message(xyz, sensor_type=sensor_type, sweep=0.1, stamp=100.0, seed=seed, n_nan=n_nan)


@pytest.mark.parametrize("sensor_type,eos", [(0, False), (0, True), (1, False), (1, True), (2, False), (3, False)])
def test_filters_and_time_sort(oracle, sensor_type, eos):
    O = oracle
    raw = _message(sensor_type=sensor_type)
    cfg = O.make_prep_cfg(crop=([-1.5, -1.0, -1.0], [1.5, 1.0, 1.0]), min_dist=4.0, rate=3, fov=2.6, sensor_type=sensor_type,
                          end_of_sweep=eos)
    kept = O.prep_filter_sort(raw, cfg, sort=False)
    # numpy formulation of Localizer.cpp:262-302
    x, y, z = raw["x"], raw["y"], raw["z"]
    fin = np.isfinite(x) & np.isfinite(y) & np.isfinite(z)
    inside = (x >= -1.5) & (x <= 1.5) & (y >= -1.0) & (y <= 1.0) & (z >= -1.0) & (z <= 1.0)
    stage1 = fin & ~inside
    idx1 = np.cumsum(stage1) - 1
    nrm = np.sqrt(x.astype(np.float32) ** 2 + (y.astype(np.float32) ** 2 + z.astype(np.float32) ** 2))
    with np.errstate(invalid="ignore"):
        ok = stage1 & (nrm > np.float32(4.0)) & (idx1 % 3 == 0) & (np.abs(np.arctan2(y, x)) < np.float32(2.6))
    assert np.array_equal(kept, np.nonzero(ok)[0])
    order = O.prep_filter_sort(raw, cfg, sort=True)
    assert sorted(order.tolist()) == kept.tolist()
    t = O.prep_times(raw, order, cfg, 100.0)
    key = {0: raw["t"].astype(np.float64), 1: raw["time"].astype(np.float64)}.get(sensor_type, raw["timestamp"])[order]
    assert np.all(np.diff(key) <= 0) if (eos and sensor_type < 2) else np.all(np.diff(key) >= 0)
    # extract_point_time (Localizer.cpp:747-777)
    if sensor_type == 0:
        rel = (raw["t"][order].astype(np.float32) * np.float32(1e-9)).astype(np.float64)
        exp = 100.0 - rel if eos else 100.0 + rel
    elif sensor_type == 1:
        exp = 100.0 - raw["time"][order].astype(np.float64) if eos else 100.0 + raw["time"][order].astype(np.float64)
    elif sensor_type == 2:
        exp = raw["timestamp"][order]
    else:
        exp = raw["timestamp"][order] * np.float64(np.float32(1e-9))
    assert np.array_equal(t, exp)


def _deskew_f64(raw_xyz, t, frames, last_q, last_p, T):
    """float64 formulation of Localizer.cpp:822-843 + State::update."""
    out_w, out_b = np.zeros((len(t), 3)), np.zeros((len(t), 3))
    ft = frames["time"]
    Rl = Rotation.from_quat(np.asarray(last_q, np.float64)).as_matrix()
    for k in range(len(t)):
        i = max(int(np.searchsorted(ft, t[k], side="right")) - 1, 0)
        f = frames[i]
        dt = t[k] - f["time"]
        w = f["w"].astype(np.float64) - f["bg"]
        Rq = Rotation.from_quat(f["q"].astype(np.float64))
        R = (Rq * Rotation.from_rotvec(w * dt)).as_matrix()
        a0 = Rq.as_matrix() @ (f["a"].astype(np.float64) - f["ba"]) + f["g"]
        p = f["p"] + f["v"] * dt + 0.5 * a0 * dt * dt
        Tw = np.eye(4); Tw[:3, :3] = R; Tw[:3, 3] = p
        wv = Tw @ T @ np.append(raw_xyz[k], 1.0)
        out_w[k] = wv[:3]
        out_b[k] = Rl.T @ (wv[:3] - np.asarray(last_p, np.float64))
    return out_w, out_b


def test_deskew_matches_float64_formulation(oracle):
    O = oracle
    raw = _message(n=3000, sensor_type=1, n_nan=0)
    cfg = O.make_prep_cfg(sensor_type=1)
    order = O.prep_filter_sort(raw, cfg, sort=True)
    t = O.prep_times(raw, order, cfg, 100.0)
    frames = synth.make_frames(100.0, 100.1, rate_hz=200.0)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = Rotation.from_euler("xyz", [0.01, -0.02, 0.03]).as_matrix()
    T[:3, 3] = [0.2, -0.1, 0.3]
    last_q = frames["q"][-2]
    last_p = frames["p"][-2]
    w, b = O.prep_deskew(raw, order, cfg, 100.0, -0.0005, frames, last_q, last_p, T)
    xyz = np.stack([raw["x"], raw["y"], raw["z"]], 1)[order].astype(np.float64)
    ew, eb = _deskew_f64(xyz, t - 0.0005, frames, last_q, last_p, T.astype(np.float64))
    assert np.abs(w[:, :3] - ew).max() < 5e-5 and np.abs(b[:, :3] - eb).max() < 5e-5
    assert np.all(w[:, 3] == 1.0) and np.all(b[:, 3] == 1.0)


def test_voxel_grid_centroids(oracle):
    O = oracle
    rng = np.random.default_rng(11)
    pts = np.ones((20000, 4), np.float32)
    pts[:, :3] = rng.uniform(-30, 30, (20000, 3))
    pts[:, 2] = rng.uniform(-2, 4, 20000)
    leaf = 1.0
    out = O.prep_voxel(pts, leaf)
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    div = ijk.max(0) + 1
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uk, inv_idx, cnt = np.unique(key, return_inverse=True, return_counts=True)
    assert len(out) == len(uk)
    sums = np.zeros((len(uk), 3))
    np.add.at(sums, inv_idx, pts[:, :3].astype(np.float64))
    assert np.abs(out[:, :3] - sums / cnt[:, None]).max() < 1e-4        # ascending voxel index, float32 sums
    # a leaf so small that the voxel count overflows int32: pcl returns the input cloud
    big = np.ones((4, 4), np.float32)
    big[:, :3] = [[0, 0, 0], [5000, 5000, 5000], [1, 2, 3], [-4000, 100, 7]]
    assert np.array_equal(O.prep_voxel(big, 0.001), big)
