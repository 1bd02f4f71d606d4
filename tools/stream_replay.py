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

if args.imu:
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    if args.premap:
        m.add(synth.sample_map(S.world, args.premap, 1005), 0.0)
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
# ---- open loop (prediction = truth + drift): the loop lives in fast_limo_b200/replay.py (bench.py runs the same code) ----
from fast_limo_b200 import replay as R

state = {"xo": None, "t_cpu": 0.0}


def on_scan(k, c):
    if om is not None and k < args.oracle:
        c0 = time.perf_counter()
        raw, frames = c["raw"], c["frames"]
        order = O.prep_filter_sort(raw, opc, sort=True)
        ow, ob = O.prep_deskew(raw, order, opc, c["stamp"], 0.0, frames, c["lq"], c["lp"], T_l2b)
        pc = np.ascontiguousarray(O.prep_voxel(ob, args.leaf)[:, :3] if args.leaf > 0 else ob[:, :3])
        if k == 0 and not args.premap:
            xo = c["truth"].copy()
        else:
            xo, Po, tr = om.update(ocfg, c["pred"], P0, args.max_iter, lim, pc)
        om.add(O.scan_to_world(xo[:14], pc))          # the oracle's own transformPointCloud
        state["t_cpu"] += time.perf_counter() - c0
        print(f"  scan {k}: oracle |dp| = {np.abs(xo[:3] - c['x_est'][:3]).max():.2e} m, map {om.size()} vs {c['map_size']}, pc2match {len(pc)} vs {c['n_pc2match']}", flush=True)
    if (k + 1) % 25 == 0 or k + 1 == args.n_scans:
        n = max(c["timed"], 1)
        print(f"scan {k+1}: map {c['map_size']} pts, pc2match {c['n_pc2match']}, pose err {c['pose_err']*1e3:.1f} mm (max so far {c['max_err']*1e3:.1f}), "
              f"per scan after 10 warm-up scans: prep {c['t_prep']/n*1e3:.2f} ms, update {c['t_upd']/n*1e3:.2f} ms, "
              f"to_world+add {c['t_add']/n*1e3:.2f} ms => {n/max(c['t_prep']+c['t_upd']+c['t_add'],1e-9):.1f} scans/s "
              f"(generation {c['t_gen']/(k+1)*1e3:.0f} ms/scan not counted)", flush=True)


res = R.replay(args.n_scans, rings=args.rings, az=args.az, dt=args.dt, imu_hz=args.imu_hz, speed=args.speed, leaf=args.leaf, max_iter=args.max_iter,
               premap=args.premap, on_scan=on_scan)
print(f"latency raw message -> pose on the host (prep + update): p50 {res['latency_ms_p50']:.2f} ms, p99 {res['latency_ms_p99']:.2f} ms, max {res['latency_ms_max']:.2f} ms")
print(f"Mapper::add (to-world + insert rule + index): mean {res['add_ms']:.2f} ms, p50 {res['add_ms_p50']:.2f} ms, max {res['add_ms_max']:.2f} ms; "
      f"index: {res['index_builds']} full builds, {res['index_updates']} row merges, {res['index_rows_moved']} rows moved")
if args.oracle:
    print(f"CPU oracle: {state['t_cpu']/args.oracle*1e3:.1f} ms per scan over the first {args.oracle} scans ({O.max_threads()} threads)")
