"""BASELINE config c3: replay of a motion-distorted HDL-64E-shaped stream through the public API —
raw message -> filters / time sort / deskew / voxel grid (device) -> iterated update -> world cloud ->
Mapper::add (map growth with the reference's down-sampling rule).  Prints per-stage times, map size and
pose error; with --oracle K the first K scans are also run through the CPU oracle (pose parity + CPU time).

usage: python tools/stream_replay.py [n_scans] [--az 2048] [--leaf 0.5] [--oracle 0] [--imu]

--imu: closed loop through fast_limo_b200.localizer (the mirror of Localizer::updateIMU / updatePointCloud): IMU samples
-> esekf::predict -> propagated frames -> deskew -> update with the carried covariance -> map add, no ground truth on
the way in (NOT yet run on a GPU: written after the round-1 GPU budget was spent).
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("n_scans", type=int, nargs="?", default=100)
ap.add_argument("--az", type=int, default=2048)
ap.add_argument("--leaf", type=float, default=0.5)
ap.add_argument("--oracle", type=int, default=0)
ap.add_argument("--max-iter", type=int, default=3)          # kitti.yaml MAX_NUM_ITERS
ap.add_argument("--rings", type=int, default=64)
ap.add_argument("--dt", type=float, default=0.1, help="sweep duration = scan period (s)")
ap.add_argument("--imu-hz", type=float, default=200.0)
ap.add_argument("--speed", type=float, default=10.0)
ap.add_argument("--imu", action="store_true", help="closed loop: the filter is driven by synthetic IMU samples")
ap.add_argument("--premap", type=int, default=0, help="points of a pre-built map of the world (config c5: 2 000 000)")
args = ap.parse_args()

S = synth.Stream(azimuths=args.az, rings=args.rings, scan_dt=args.dt, imu_hz=args.imu_hz, speed=args.speed)
BIG = 1 << 20
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
filt = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=args.leaf if args.leaf > 0 else None, sensor_type=1)
P0, lim = synth.default_P0(), np.full(23, 0.001)
T_l2b = np.eye(4, dtype=np.float32)
rng = np.random.default_rng(1)
om = None
if args.oracle:
    from oracle import oracle as O
    om = O.OracleMap()
    ocfg = O.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=O.max_threads())
    opc = O.make_prep_cfg(crop=([-1, -1, -1], [1, 1, 1]), min_dist=3.0, leaf=args.leaf if args.leaf > 0 else None, sensor_type=1)

if args.premap:
    m.add(synth.sample_map(S.world, args.premap, 1005), 0.0)

if args.imu:
    from fast_limo_b200.localizer import Localizer, LocalizerConfig
    x0 = S.state(0.0)
    loc = Localizer(m, LocalizerConfig(filters=filt, MAX_NUM_ITERS=args.max_iter), pos=x0[0:3], quat=x0[3:7], vel=x0[14:17])
    t_imu, t_pred, t_scan, n_timed, errs, lat = 0.0, 0.0, 0.0, 0, [], []
    for k in range(args.n_scans):
        raw, stamp = S.scan(k)
        samples = list(zip(*S.imu(t_imu, stamp + args.dt + 1.0 / args.imu_hz)))
        t_imu = stamp + args.dt + 1.0 / args.imu_hz
        a0 = time.perf_counter()
        for smp in samples:
            loc.updateIMU(*smp)
        a1 = time.perf_counter()
        ok = loc.updatePointCloud(raw, stamp)
        a2 = time.perf_counter()
        if k >= 10:
            t_pred += a1 - a0; t_scan += a2 - a1; n_timed += 1
            lat.append(a2 - a1)
        errs.append(float(np.linalg.norm(loc.x[0:3] - S.state(loc.imu_stamp)[0:3])))
        if (k + 1) % 25 == 0 or k + 1 == args.n_scans:
            print(f"scan {k+1}: {'registered' if ok else loc.last.get('null')}, map {m.size()} pts, pc2match {loc.last.get('n_pc2match')}, "
                  f"passes {loc.last.get('passes')}, pose err {errs[-1]*1e3:.1f} mm (max {max(errs)*1e3:.1f}); per scan: IMU side "
                  f"{t_pred/max(n_timed,1)*1e3:.2f} ms ({len(samples)} predictions), LiDAR callback {t_scan/max(n_timed,1)*1e3:.2f} ms", flush=True)
    if lat:
        l_ = np.array(lat) * 1e3
        print(f"LiDAR callback latency: p50 {np.percentile(l_, 50):.2f} ms, p99 {np.percentile(l_, 99):.2f} ms")
    sys.exit(0)
lat_pose = []
t_gen = t_prep = t_upd = t_add = 0.0
WARM = 10                                      # scans left out of the averages (allocations, lazy module loading)
errs, sizes, n_pc = [], [], []
t_cpu = 0.0
x_est = None
prev_end = 0.0
for k in range(args.n_scans):
    g0 = time.perf_counter()
    raw, stamp = S.scan(k)
    t_gen += time.perf_counter() - g0
    a0 = time.perf_counter()
    n_kept, t_last = m.prep_filter_sort(raw, stamp, filt)
    h0 = time.perf_counter()                      # harness work (synthetic IMU frames, prediction) is not pipeline time
    frames = S.frames(prev_end, t_last)
    truth = S.state(t_last)
    # prediction at the end of the sweep: truth + a small drift (what IMU propagation would hand over)
    pred = truth.copy()
    pred[:3] += rng.normal(0, 0.02, 3)
    dq = synth.quat_from_rpy(*rng.normal(0, 0.002, 3))
    x, y, z, w = truth[3:7]; a, b, c, d = dq
    q2 = np.array([w * a + x * d + y * c - z * b, w * b - x * c + y * d + z * a, w * c + x * b - y * a + z * d, w * d - x * a - y * b - z * c])
    pred[3:7] = q2 / np.linalg.norm(q2)
    # the IMU frames come from the same (drifted) filter: move them rigidly so that the frame at t_last IS the prediction
    Rt, Rp = synth.quat_to_R(truth[3:7]), synth.quat_to_R(pred[3:7])
    dR = Rp @ Rt.T
    dt_ = pred[:3] - dR @ truth[:3]
    fq = frames["q"].astype(np.float64)
    fR = np.stack([dR @ synth.quat_to_R(q) for q in fq])
    from scipy.spatial.transform import Rotation
    frames["q"] = Rotation.from_matrix(fR).as_quat().astype(np.float32)
    frames["p"] = (frames["p"].astype(np.float64) @ dR.T + dt_).astype(np.float32)
    frames["v"] = (frames["v"].astype(np.float64) @ dR.T).astype(np.float32)
    frames["g"] = (frames["g"].astype(np.float64) @ dR.T).astype(np.float32) * 0 + frames["g"]   # gravity stays world-fixed
    lq, lp = pred[3:7].astype(np.float32), pred[:3].astype(np.float32)
    a0 += time.perf_counter() - h0
    n_pc2 = m.prep_deskew(frames, lq, lp, T_l2b, 0.0)
    a1 = time.perf_counter()
    if k == 0 and not args.premap:
        x_est, passes = truth.copy(), 0           # the first mapped scan initialises the map (zero matches); anchored at the truth
    else:
        x_est, Pn, passes = m.update(pred, P0, args.max_iter, lim)
    a2 = time.perf_counter()
    if om is not None and k < args.oracle:
        world = m.scan_to_world(x_est)            # the oracle arm below needs the world cloud on the host
    m.add_scan(x_est, t_last)                     # transformPointCloud + Mapper::add without leaving the device
    a3 = time.perf_counter()
    if k >= WARM:
        t_prep += a1 - a0; t_upd += a2 - a1; t_add += a3 - a2
        lat_pose.append(a2 - a0)
    errs.append(float(np.linalg.norm(x_est[:3] - truth[:3]))); sizes.append(m.size()); n_pc.append(n_pc2)
    if om is not None and k < args.oracle:
        c0 = time.perf_counter()
        order = O.prep_filter_sort(raw, opc, sort=True)
        ow, ob = O.prep_deskew(raw, order, opc, stamp, 0.0, frames, lq, lp, T_l2b)
        pc = O.prep_voxel(ob, args.leaf)[:, :3] if args.leaf > 0 else ob[:, :3]
        if k == 0:
            xo = truth.copy()
        else:
            xo, Po, tr = om.update(ocfg, pred, P0, args.max_iter, lim, np.ascontiguousarray(pc))
        wo = om.match(ocfg, xo[:14], np.ascontiguousarray(pc))["world"] if om.size() else None
        if wo is None:     # empty map: transform with the GPU helper's arithmetic (bit-identical to the oracle's)
            wo = world
        om.add(np.ascontiguousarray(wo))
        t_cpu += time.perf_counter() - c0
        print(f"  scan {k}: oracle |dp| = {np.abs(xo[:3] - x_est[:3]).max():.2e} m, map {om.size()} vs {m.size()}, pc2match {len(pc)} vs {n_pc2}", flush=True)
    prev_end = t_last
    if (k + 1) % 25 == 0 or k + 1 == args.n_scans:
        print(f"scan {k+1}: map {sizes[-1]} pts, pc2match {n_pc[-1]}, pose err {errs[-1]*1e3:.1f} mm (max so far {max(errs)*1e3:.1f}), "
              f"per scan after {WARM} warm-up scans: prep {t_prep/max(k+1-WARM,1)*1e3:.2f} ms, update {t_upd/max(k+1-WARM,1)*1e3:.2f} ms, "
              f"to_world+add {t_add/max(k+1-WARM,1)*1e3:.2f} ms => {max(k+1-WARM,1)/max(t_prep+t_upd+t_add,1e-9):.1f} scans/s "
              f"(generation {t_gen/(k+1)*1e3:.0f} ms/scan not counted)", flush=True)
if lat_pose:
    lp_ = np.array(lat_pose) * 1e3
    print(f"latency raw message -> pose on the host (prep + update): p50 {np.percentile(lp_, 50):.2f} ms, p99 {np.percentile(lp_, 99):.2f} ms, max {lp_.max():.2f} ms")
if args.oracle:
    print(f"CPU oracle: {t_cpu/args.oracle*1e3:.1f} ms per scan over the first {args.oracle} scans ({O.max_threads()} threads)")
