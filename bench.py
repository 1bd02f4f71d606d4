#!/usr/bin/env python
"""bench.py — headline benchmark of the registration hot path (BASELINE.json).

  metric   : scans/s, 64-beam 131 072-point synthetic scan against a 5 000 000-point map,
             3 IKFoM passes per scan (BASELINE.json configs[1] = "c2")
  step     : one whole scan registration = esekf::update_iterated_dyn_share_modified on one scan
             (per pass: fused kNN + plane fit + Jacobian + H^T H / H^T h tiles, filter step in a resident
             filter CTA on the device; the host only forms the final covariance), pose re-seeded each step
  value    : scans/s with the scans already resident in HBM (flimo_scan_set_device + flimo_update)
  e2e      : scans/s through the C ABI with HOST buffers (pinned scan -> H2D inside the timed region,
             updated state + covariance back on the host)
  roofline : algorithmic bytes (528 B per point-match) / mean device time of the fused
             kernel (CUDA events on its launch stream, collected by the library) vs the measured HBM peak
  cpu_baseline / --impl reference : the CPU restatement of the reference path (oracle/, OpenMP at
             the reference's three loops) on the box's host cores, bounded sample

  repeats  : the timed loop of --steps steps is run --repeats times (default 9) after ONE warm-up; the median repeat is
             reported (value, ms_per_step, e2e), min / max beside it — a 20-step loop is 3 ms long and a single one is noisy
  caps_kitti : e2e scans/s with the reference's shipped caps (config/kitti.yaml: 10 000 queried points, 5 000 rows, 4 passes)
  streams  : BASELINE configs c3 / c5 in small (--stream-scans each): the whole per-scan sequence raw message -> filters / deskew /
             voxel grid -> update -> Mapper::add with map growth; scans/s, latency p50 / p99, Mapper::add time, index statistics
  c4       : (at --gpus 8, or --c4) 300 000-point rosette scan against a 20 M-point map, scan sharded over the ranks

N > 1 (torchrun): the scan is sharded across ranks (replicated map); every rank runs the whole update on its device, the 96
doubles of a pass travel over NVLink peer memory between the ranks' filter CTAs (--exchange peer; shm / nccl: round-1 paths).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec (128k-pt scan vs 5M-pt map, 3 IKFoM passes)"
WORKLOAD = "c2: 64x2048-pt synthetic spinning scan vs 5M-pt planar-world map, 3 passes, caps raised"
A_PM = 528          # algorithmic bytes per point-match
N_SCANS = 4         # distinct scans rotated through the steps
MAX_ITER = 2        # MAX_NUM_ITERS = 2  => 3 passes with LIMITS = 0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def _traffic():
    """DRAM bytes per launch of the fused kernel from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "k1_dram_bytes.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(n_scans):
    """The c2 case plus n_scans scans taken from slightly different true poses (all seeded)."""
    from fast_limo_b200 import synth
    case = synth.make_case("c2")
    world = synth.make_world(1002, 120.0, 70.0, 25.0, 60)
    scans, inits = [case.scan], [case.init]
    for k in range(1, n_scans):
        pos = np.array([0.3 + 0.7 * k, -0.2 - 0.4 * k, 1.8])
        quat = synth.quat_from_rpy(0.01, -0.02, 0.4 + 0.15 * k)
        scans.append(synth.spinning_scan(world, pos, quat, 64, 2048, 1002 + 10 * k))
        st = synth.make_state(pos + np.array([0.03, -0.03, 0.025]), quat)
        inits.append(st)
    return case, scans, inits


def host_threads():
    """Host threads the CPU arm may use: the cores this process is allowed on.  (torchrun exports
    OMP_NUM_THREADS=1 to its workers; the oracle's loops take their thread count as an argument, like the
    reference's `num_threads` parameter, so that setting does not bind them.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """CPU arm: the restated reference path (oracle/) on the host cores.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")     # before libgomp starts
    from fast_limo_b200 import synth
    from oracle import oracle as O
    case, scans, inits = make_inputs(min(N_SCANS, 2))
    threads = host_threads()
    om = O.OracleMap()
    om.add(case.map_pts)
    cfg = O.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20, num_threads=threads)
    P0 = synth.default_P0()
    # one full scan to size the per-step sample: the whole run should stay within ~100 s of CPU wall time
    om.update(cfg, inits[0], P0, MAX_ITER, 0.0, scans[0])
    t0 = time.perf_counter()
    om.update(cfg, inits[0], P0, MAX_ITER, 0.0, scans[0])
    t_full = time.perf_counter() - t0
    stride = max(1, int(np.ceil((args.steps + args.warmup) * t_full / 100.0)))
    sub = [np.ascontiguousarray(s[::stride]) for s in scans]
    for w in range(args.warmup):
        om.update(cfg, inits[w % len(sub)], P0, MAX_ITER, 0.0, sub[w % len(sub)])
    t0 = time.perf_counter()
    for s in range(args.steps):
        om.update(cfg, inits[s % len(sub)], P0, MAX_ITER, 0.0, sub[s % len(sub)])
    dt = time.perf_counter() - t0
    frac = sub[0].shape[0] / scans[0].shape[0]
    v = args.steps * frac / dt                     # scans/s: each step registered `frac` of a scan
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(),
        "details": {"parallelism": f"openmp x{threads}",
                    "sample": f"every {stride}-th scan point per step ({sub[0].shape[0]} of {scans[0].shape[0]}), value scaled to whole scans"},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 3 passes on every {stride}-th point of the c2 scans; full scan measured at {1.0 / t_full:.3f} scans/s"},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _config():
    """`config` is identical in both arms (the driver compares them); everything arm-specific lives beside it."""
    return {"workload": WORKLOAD, "scan_points": 131072, "map_points": 5000000, "passes_per_scan": MAX_ITER + 1,
            "MAX_NUM_PC2MATCH": 1 << 20, "MAX_NUM_MATCHES": 1 << 20,
            "l2_policy": "inputs larger than L2 (map index of several GB, 4 rotating scans)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cell", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--repeats", type=int, default=9, help="the timed loop of --steps steps is repeated; the MEDIAN repeat is reported")
    ap.add_argument("--no-extra", action="store_true", help="skip the caps_kitti / c4 legs")
    ap.add_argument("--stream-scans", type=int, default=70, help="scans of the c3 / c5 stream replays (streams field)")
    ap.add_argument("--c4", action="store_true", help="also run config c4 (300k-pt rosette scan, 20M-pt map); default at --gpus 8")
    ap.add_argument("--exchange", choices=["peer", "shm", "nccl"], default="peer",
                    help="N>1: how the 96 doubles per pass are summed over ranks (peer: NVLink peer stores from inside the "
                         "filter kernel, filter step on the device | shm: fused host-segment exchange | nccl: NCCL all-reduce)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fast_limo_b200 import api, synth
    from fast_limo_b200.dist import attach_exchange, shard_bounds, sharded_update

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)
    P0 = synth.default_P0()
    lim = np.zeros(23)

    def bcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    class Arm:
        """One workload on this rank's GPU: map, rotating scans (device + pinned host copies), timed loops."""

        def __init__(self, map_pts, scans, inits, truths, max_iter, pc2match=1 << 20, matches=1 << 20):
            self.scans, self.inits, self.truths, self.max_iter = scans, inits, truths, max_iter
            self.n_pts = scans[0].shape[0]
            self.m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=matches, MAX_NUM_PC2MATCH=pc2match, knn_cell=args.cell), device=local)
            self.m.add(map_pts, 0.0)
            self.d_scans, self.h_scans = [], []
            for s in scans:
                s4 = np.zeros((self.n_pts, 4), np.float32)
                s4[:, :3] = s[:, :3]
                self.d_scans.append(torch.from_numpy(s4).cuda())
                self.h_scans.append(torch.from_numpy(s4).pin_memory())
            self.n_q = min(self.n_pts, pc2match)                       # points the passes query (first-N rule)
            self.lo, self.hi = shard_bounds(self.n_q, rank, world)
            self.red = torch.zeros(96, dtype=torch.float64, device="cuda")
            self.shm = None
            if world > 1 and args.exchange == "shm":
                self.shm = attach_exchange(self.m, rank, world, bcast)
            if world > 1 and args.exchange == "peer":
                handles = [None] * world
                dist.all_gather_object(handles, self.m.peer_export())
                self.m.peer_attach(rank, world, handles)
            if world > 1:
                dist.barrier()
            self.h_stream = torch.cuda.ExternalStream(self.m.stream())

        def register(self, k, from_host):
            """One scan registration.  Returns (state, passes)."""
            m, lo, hi = self.m, self.lo, self.hi
            if from_host:
                # this step's scan: H2D from pinned memory (already under way if the previous step prefetched it); then
                # start the copy of the NEXT step's scan so that it overlaps this registration.  N > 1: every rank
                # uploads only ITS slice of the queried points and binds it as its whole scan.
                m.set_scan_host(self.h_scans[k].data_ptr() + 16 * lo, hi - lo, 16)
                m.prefetch_scan_host(self.h_scans[(k + 1) % len(self.scans)].data_ptr() + 16 * lo, hi - lo, 16)
            else:
                m.set_scan_device(self.d_scans[k].data_ptr(), self.n_pts, 16)
            if world == 1:
                x, P, passes = m.update(self.inits[k], P0, self.max_iter, lim)
                return x, passes
            if not from_host:
                m.shard(lo, hi)
            if args.exchange == "peer":
                x, P, passes = m.update_peer(self.inits[k], P0, self.max_iter, lim)
                return x, passes
            if self.shm is not None:
                x, P, passes = m.update_exchange(self.inits[k], P0, self.max_iter, lim)
                return x, passes
            # NCCL: everything of the pass is ordered on the HANDLE's stream (stream NULL in the ABI = the handle's stream)
            with torch.cuda.stream(self.h_stream):
                def local_pass(state):
                    m.match_async(state, self.red.data_ptr(), None)
                    return self.red

                def all_reduce(t):
                    dist.all_reduce(t)                     # 96 doubles: HTH tri + HTh + counters (NCCL)
                    return t.cpu().numpy()

                x, P, passes = sharded_update(m, self.inits[k], P0, self.max_iter, lim, local_pass, all_reduce)
            return x, passes

        def timed(self, steps, from_host):
            """`steps` registrations between two events on the handle's stream; max over ranks.  Returns (ms, x, passes)."""
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.h_stream)
            x, passes = None, 0
            for s in range(steps):
                x, passes = self.register(s % len(self.scans), from_host)
            e1.record(self.h_stream)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()), x, passes

        def measure(self, steps, warmup, repeats, from_host):
            """W warm-up steps, then `repeats` timed loops of `steps` steps each: (median ms, all ms, x, passes)."""
            self.timed(warmup, from_host)
            runs = []
            for _ in range(max(1, repeats)):
                ms, x, passes = self.timed(steps, from_host)
                runs.append(ms)
            return float(np.median(runs)), runs, x, passes

        def pose_err(self, x, steps):
            k_last = (steps - 1) % len(self.scans)
            return float(np.abs(x[:3] - self.truths[k_last][:3]).max())

        def close(self):
            self.m.close()
            if self.shm is not None:
                self.shm.close()
                if rank == 0:
                    self.shm.unlink()

    case, scans, inits = make_inputs(N_SCANS)
    truths = [st.copy() for st in inits]
    for t in truths:
        t[:3] -= np.array([0.03, -0.03, 0.025])
    arm = Arm(case.map_pts, scans, inits, truths, MAX_ITER)
    m = arm.m

    sampler = ClockSampler(local)
    sampler.start()
    st0 = m.stats()
    ms, runs, x, passes = arm.measure(args.steps, args.warmup, args.repeats, False)
    st1 = m.stats()
    ms_e2e, runs_e2e, x2, _ = arm.measure(args.steps, args.warmup, args.repeats, True)
    clocks = sampler.stop()
    # Every 8th flimo_update runs with one launch per pass and CUDA events around the kernel.  A run with very few
    # steps may have none of those inside the timed region: time the kernel over 8 further scans right after it
    # (all ranks take the same branch: the call count is the same on every rank).
    st_k0, st_k1, kernel_where = st0, st1, "inside the timed region"
    if st1["match_timed"] - st0["match_timed"] == 0:
        st_k0 = m.stats()
        if world == 1:
            arm.timed(8, False)
            kernel_where = "over 8 further scans right after the timed region (steps * repeats < 8)"
        else:
            # the peer path never leaves the device: time the kernel of this rank's shard with one launch per pass right after
            # the timed region — per scan one pass at the initial pose and two at the registered one, as an update sees them
            for k in range(2 * len(scans)):
                kk = k % len(scans)
                m.set_scan_device(arm.d_scans[kk].data_ptr(), arm.n_pts, 16)
                m.shard(arm.lo, arm.hi)
                xk, _, _ = m.update_peer(inits[kk], P0, MAX_ITER, lim)
                for st in (inits[kk], xk, xk):
                    m.match(st)
            kernel_where = "one launch per pass on this rank's shard right after the timed region (8 scans x 3 poses)"
        st_k1 = m.stats()
    st_end = m.stats()
    pose_err = arm.pose_err(x, args.steps)
    n_pts, lo, hi = arm.n_pts, arm.lo, arm.hi
    map_gb = st1["map_bytes"] / 1e9
    arm.close()

    extra = {}
    if not args.no_extra and world == 1:
        # (1) the reference's shipped caps (config/kitti.yaml:76-78): 10 000 queried points, 5 000 Jacobian rows, 4 passes max.
        #     n_valid > MAX_NUM_MATCHES at every pass: the first-N rule is resolved on the device (pass repeated with a row limit).
        k_arm = Arm(case.map_pts, scans, inits, truths, 3, pc2match=10000, matches=5000)
        s0 = k_arm.m.stats()
        k_ms, k_runs, kx, k_passes = k_arm.measure(min(args.steps, 500), args.warmup, min(args.repeats, 5), True)
        s1 = k_arm.m.stats()
        n_upd = (args.warmup + min(args.steps, 500) * max(1, min(args.repeats, 5)))
        extra["caps_kitti"] = {"value": min(args.steps, 500) / (k_ms / 1e3), "unit": "scans/s", "measured": "e2e (pinned host scan -> state on host)",
                               "MAX_NUM_PC2MATCH": 10000, "MAX_NUM_MATCHES": 5000, "MAX_NUM_ITERS": 3, "passes": k_passes,
                               "kernel_launches_per_scan": (s1["kernel_launches"] - s0["kernel_launches"]) / n_upd,
                               "passes_incl_repeats_per_scan": (s1["match_launches"] - s0["match_launches"]) / n_upd,
                               "pose_err_m": k_arm.pose_err(kx, min(args.steps, 500))}
        k_arm.close()
    if not args.no_extra and world == 1:
        # (1b) BASELINE configs c3 / c5 in small: stream replays through the whole per-scan sequence (raw message -> filters / deskew /
        #      voxel grid -> update -> Mapper::add with map growth); generation of the synthetic messages is not counted
        from fast_limo_b200 import replay as R
        c3 = R.replay(args.stream_scans, rings=64, az=2048, dt=0.1, imu_hz=200.0, speed=10.0, leaf=0.5, max_iter=3)
        c5 = R.replay(args.stream_scans, rings=32, az=1024, dt=0.02, imu_hz=400.0, speed=12.0, leaf=0.5, max_iter=3, premap=2_000_000)
        extra["streams"] = {
            "c3": dict(c3, workload="c3: 64x2048-pt HDL-64E-shaped sweeps at 10 m/s, map grows from empty (kitti-style pipeline, caps raised)"),
            "c5": dict(c5, workload="c5: 32x1024-pt sweeps at 50 Hz with deskew against a pre-built 2M-pt map that keeps growing; "
                                    "latency = raw message -> pose on the host"),
            "note": "scans_per_s = 1 / (prep + update + Mapper::add) per scan through the public API, host buffers in, pose out; first 10 scans untimed"}
    if not args.no_extra and (args.c4 or world == 8):
        if True:
            # (2) BASELINE config c4: 300 000-point rosette scan against a 20 M-point map (index >> L2), scan sharded over the ranks
            # rank 0 builds the case (20 M map points take ~15 s of host time); the others receive it over NCCL
            if rank == 0:
                c4 = synth.make_case("c4")
                blobs = [np.ascontiguousarray(c4.map_pts, np.float32), np.ascontiguousarray(c4.scan, np.float32),
                         np.ascontiguousarray(c4.init, np.float64), np.ascontiguousarray(c4.truth, np.float64)]
            else:
                blobs = [np.zeros((20_000_000, 3), np.float32), np.zeros((300_000, 3), np.float32), np.zeros(26), np.zeros(26)]
            if world > 1:
                for i, b in enumerate(blobs):
                    t = torch.from_numpy(b).cuda()
                    dist.broadcast(t, src=0)
                    blobs[i] = t.cpu().numpy()
                    del t
                torch.cuda.empty_cache()
            c4_map, c4_scan, c4_init, c4_truth = blobs
            c_arm = Arm(c4_map, [c4_scan], [c4_init], [c4_truth], MAX_ITER)
            c_ms, c_runs, cx, c_passes = c_arm.measure(min(args.steps, 200), args.warmup, min(args.repeats, 5), False)
            extra["c4"] = {"value": min(args.steps, 200) / (c_ms / 1e3), "unit": "scans/s", "workload": "c4: 300k-pt rosette scan vs 20M-pt map, 3 passes",
                           "ms_per_scan": c_ms / min(args.steps, 200), "passes": c_passes, "n_gpus": world,
                           "pose_err_m": float(np.abs(cx[:3] - c4_truth[:3]).max()), "index_gb": c_arm.m.stats()["map_bytes"] / 1e9}
            c_arm.close()

    if rank == 0:
        value = args.steps / (ms / 1e3)
        launches = st1["match_launches"] - st0["match_launches"]
        n_timed = st_k1["match_timed"] - st_k0["match_timed"]
        if n_timed == 0:
            raise SystemExit("no event-timed launch of the fused kernel (FLIMO_TIME_EVERY=0?)")
        k1_ms = (st_k1["match_ms_total"] - st_k0["match_ms_total"]) / n_timed
        # in-kernel time per pass of the resident kernels, over the timed region of the `value` arm (command posted -> sums complete
        # -> filter step -> next command, i.e. everything between two poses; N > 1: includes the wait for the slowest rank)
        pp = st1["persist_passes"] - st0["persist_passes"]
        k1_in_ms = (st1["persist_ms_total"] - st0["persist_ms_total"]) / pp if pp else None
        xch_us = (st1["exchange_ms_total"] - st0["exchange_ms_total"]) / pp * 1e3 if pp else None
        peak, peak_src = _peaks()
        achieved = (hi - lo) * A_PM / (k1_ms * 1e-3) / 1e9
        exch = {"peer": "NVLink peer stores inside the filter kernel", "shm": "fused host-segment exchange", "nccl": "NCCL all-reduce"}[args.exchange]
        out = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": _config(),
            "repeats": {"n": len(runs), "reported": "median", "value_min": args.steps / (max(runs) / 1e3), "value_max": args.steps / (min(runs) / 1e3),
                        "e2e_min": args.steps / (max(runs_e2e) / 1e3), "e2e_max": args.steps / (min(runs_e2e) / 1e3)},
            "details": {"arithmetic": "float32 per point (kNN, plane fit, residual, Jacobian row), float64 normal equations and filter algebra",
                        "parallelism": "scan-shard x%d (replicated map, 96 doubles summed per pass via %s)" % (world, exch if world > 1 else "no exchange"),
                        "update": "device-resident: tiles kernel + filter CTA, no host between passes; every 8th scan one launch per pass (event-timed)",
                        "passes_per_scan": passes, "map_index_gb": map_gb, "pose_err_m": pose_err, "knn_cell": st1["knn_cell"], "levels": st1["n_levels"],
                        "update_stalls": int(st_end["update_stalls"]),
                        # per pass on rank 0: own tiles complete -> sums of all ranks in hand (collect; N > 1: + NVLink peer stores + wait for the slowest rank)
                        "exchange_us_per_pass": xch_us},
            "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": "scans/s", "h2d_bytes_per_step": n_pts * 16,   # summed over ranks
                    "d2h_bytes_per_step": 160 * 16},
            # kernels launched during ONE timed loop of `steps` steps (counted over warm-up + all repeats, scaled)
            "gpu_launches": int(round((st1["kernel_launches"] - st0["kernel_launches"]) * args.steps / (args.warmup + args.steps * len(runs)))),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "peak_source": peak_src, "kernel": "match_reduce_kernel",
                         "kernel_ms": k1_ms, "passes_all_repeats": int(launches), "passes_event_timed": int(n_timed), "kernel_timed": kernel_where,
                         "kernel_ms_in_persistent_kernel": k1_in_ms, "bytes_per_launch": (hi - lo) * A_PM,
                         "note": "kernel_ms = CUDA-event time of one-launch-per-pass executions (every 8th scan); the other scans run all "
                                 "passes inside the resident tiles kernel, timed in-kernel with %globaltimer (command posted -> pass sums complete, "
                                 "plus the filter step)"},
            "clocks": clocks,
        }
        out.update(extra)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(case, scans, inits)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(case, scans, inits):
    """The oracle (CPU restatement of the reference path) on the host cores, bounded sample."""
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")     # before libgomp starts
    from fast_limo_b200 import synth
    from oracle import oracle as O
    threads = host_threads()
    om = O.OracleMap()
    om.add(case.map_pts)
    cfg = O.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20, num_threads=threads)
    P0 = synth.default_P0()
    om.update(cfg, inits[0], P0, MAX_ITER, 0.0, scans[0])          # warm-up (thread pool, page faults)
    n, t0 = 0, time.perf_counter()
    while n < 3 or (time.perf_counter() - t0 < 6.0 and n < 12):
        om.update(cfg, inits[n % len(scans)], P0, MAX_ITER, 0.0, scans[n % len(scans)])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "scans/s", "cores": threads, "kind": "port",
            "sample": f"{n} full c2 scans (3 passes each), oracle update_iterated_dyn_share_modified, OpenMP x{threads}"}


if __name__ == "__main__":
    main()
