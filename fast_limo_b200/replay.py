"""Synthetic stream replay through the public API (BASELINE configs c3 / c5): raw message -> filters / time sort / deskew /
voxel grid (device) -> iterated update -> world cloud -> Mapper::add, scan after scan (the per-scan sequence of
Localizer::updatePointCloud, Localizer.cpp:326-377, with the prediction handed in the way IMU propagation would).
Used by bench.py (`streams` field) and tools/stream_replay.py; generation of the synthetic messages is not pipeline time."""
import time

import numpy as np

from . import api, synth


def _predicted(S, frames, t_last, rng):
    """Prediction at the end of the sweep (truth + a small drift) and the IMU frames moved rigidly onto it."""
    from scipy.spatial.transform import Rotation
    truth = S.state(t_last)
    pred = truth.copy()
    pred[:3] += rng.normal(0, 0.02, 3)
    dq = synth.quat_from_rpy(*rng.normal(0, 0.002, 3))
    x, y, z, w = truth[3:7]
    a, b, c, d = dq
    q2 = np.array([w * a + x * d + y * c - z * b, w * b - x * c + y * d + z * a, w * c + x * b - y * a + z * d, w * d - x * a - y * b - z * c])
    pred[3:7] = q2 / np.linalg.norm(q2)
    Rt, Rp = synth.quat_to_R(truth[3:7]), synth.quat_to_R(pred[3:7])
    dR = Rp @ Rt.T
    dt_ = pred[:3] - dR @ truth[:3]
    fR = np.stack([dR @ synth.quat_to_R(q) for q in frames["q"].astype(np.float64)])
    frames["q"] = Rotation.from_matrix(fR).as_quat().astype(np.float32)
    frames["p"] = (frames["p"].astype(np.float64) @ dR.T + dt_).astype(np.float32)
    frames["v"] = (frames["v"].astype(np.float64) @ dR.T).astype(np.float32)
    return truth, pred


def replay(n_scans, rings=64, az=2048, dt=0.1, imu_hz=200.0, speed=10.0, leaf=0.5, max_iter=3, premap=0, warm=10, device=0, on_scan=None):
    """Runs the stream; returns a dict of per-stage times (ms per scan after `warm` scans), scans/s, latency percentiles of
    raw message -> pose on the host (prep + update), map size, pose error vs ground truth and the index statistics."""
    S = synth.Stream(azimuths=az, rings=rings, scan_dt=dt, imu_hz=imu_hz, speed=speed)
    big = 1 << 20
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=big, MAX_NUM_PC2MATCH=big), device=device)
    filt = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=leaf if leaf > 0 else None, sensor_type=1)
    P0, lim, T_l2b = synth.default_P0(), np.full(23, 0.001), np.eye(4, dtype=np.float32)
    rng = np.random.default_rng(1)
    if premap:
        m.add(synth.sample_map(S.world, premap, 1005), 0.0)
    t_gen = t_prep = t_upd = t_add = 0.0
    lat, adds, errs, n_pc = [], [], [], []
    prev_end, timed = 0.0, 0
    for k in range(n_scans):
        g0 = time.perf_counter()
        raw, stamp = S.scan(k)
        t_gen += time.perf_counter() - g0
        a0 = time.perf_counter()
        _, t_last = m.prep_filter_sort(raw, stamp, filt)
        h0 = time.perf_counter()                      # harness work (synthetic IMU frames, prediction) is not pipeline time
        frames = S.frames(prev_end, t_last)
        truth, pred = _predicted(S, frames, t_last, rng)
        lq, lp = pred[3:7].astype(np.float32), pred[:3].astype(np.float32)
        a0 += time.perf_counter() - h0
        n_pc2 = m.prep_deskew(frames, lq, lp, T_l2b, 0.0)
        a1 = time.perf_counter()
        if k == 0 and not premap:
            x_est, passes = truth.copy(), 0           # the first mapped scan initialises the map (zero matches); anchored at the truth
        else:
            x_est, _, passes = m.update(pred, P0, max_iter, lim)
        a2 = time.perf_counter()
        m.add_scan(x_est, t_last)                     # transformPointCloud + Mapper::add without leaving the device
        a3 = time.perf_counter()
        if k >= warm:
            t_prep += a1 - a0
            t_upd += a2 - a1
            t_add += a3 - a2
            lat.append(a2 - a0)
            adds.append(a3 - a2)
            timed += 1
        errs.append(float(np.linalg.norm(x_est[:3] - truth[:3])))
        n_pc.append(n_pc2)
        prev_end = t_last
        if on_scan is not None:
            on_scan(k, dict(m=m, raw=raw, stamp=stamp, frames=frames, lq=lq, lp=lp, pred=pred, truth=truth, x_est=x_est, passes=passes, n_pc2match=n_pc2,
                            t_last=t_last, map_size=m.size(), pose_err=errs[-1], max_err=max(errs), timed=timed, t_prep=t_prep, t_upd=t_upd, t_add=t_add,
                            t_gen=t_gen))
    st = m.stats()
    n = max(timed, 1)
    l_ms, a_ms = np.array(lat or [0.0]) * 1e3, np.array(adds or [0.0]) * 1e3
    out = {"scans": n_scans, "timed_scans": timed, "scans_per_s": n / max(t_prep + t_upd + t_add, 1e-9), "prep_ms": t_prep / n * 1e3, "update_ms": t_upd / n * 1e3,
           "add_ms": t_add / n * 1e3, "add_ms_p50": float(np.percentile(a_ms, 50)), "add_ms_max": float(a_ms.max()),
           "latency_ms_p50": float(np.percentile(l_ms, 50)), "latency_ms_p99": float(np.percentile(l_ms, 99)), "latency_ms_max": float(l_ms.max()),
           "map_points": int(m.size()), "pc2match_mean": float(np.mean(n_pc)), "pose_err_m_max": float(max(errs)), "pose_err_m_last": errs[-1],
           "index_builds": int(st["index_builds"]), "index_updates": int(st["index_updates"]), "index_rows_moved": int(st["index_rows_moved"]),
           "generation_ms": t_gen / max(n_scans, 1) * 1e3}
    m.close()
    return out
