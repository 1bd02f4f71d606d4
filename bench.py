#!/usr/bin/env python
"""bench.py — headline benchmark of the registration hot path (BASELINE.json).

  metric   : scans/s, 64-beam 131 072-point synthetic scan against a 5 000 000-point map,
             3 IKFoM passes per scan (BASELINE.json configs[1] = "c2")
  step     : one whole scan registration = esekf::update_iterated_dyn_share_modified on one scan
             (per pass: fused kNN + plane fit + Jacobian + H^T H / H^T h kernel, 96 doubles to the
             host, 23x23 filter algebra on the host), pose re-seeded each step
  value    : scans/s with the scans already resident in HBM (flimo_scan_set_device + flimo_update)
  e2e      : scans/s through the C ABI with HOST buffers (pinned scan -> H2D inside the timed region,
             updated state + covariance back on the host)
  roofline : algorithmic bytes (528 B per point-match) / mean device time of the fused
             kernel (CUDA events on its launch stream, collected by the library) vs the measured HBM peak
  cpu_baseline / --impl reference : the CPU restatement of the reference path (oracle/, OpenMP at
             the reference's three loops) on the box's host cores, bounded sample

N > 1 (torchrun): the scan is sharded across ranks (replicated map), each pass all-reduces the 96
doubles of the normal equations over NCCL, every rank runs the identical filter algebra; strong scaling.
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
        "config": {"workload": WORKLOAD, "parallelism": f"openmp x{threads}",
                   "sample": f"every {stride}-th scan point per step ({sub[0].shape[0]} of {scans[0].shape[0]}), value scaled to whole scans"},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x 3 passes on every {stride}-th point of the c2 scans; full scan measured at {1.0 / t_full:.3f} scans/s"},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cell", type=float, default=0.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exchange", choices=["peer", "shm", "nccl"], default="peer",
                    help="N>1: how the 96 doubles per pass are summed over ranks (peer: NVLink peer stores from inside the "
                         "registration kernel, filter step on the device | shm: fused host-segment exchange | nccl: NCCL all-reduce)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from fast_limo_b200 import api, synth
    from fast_limo_b200.dist import attach_exchange, shard_bounds, sharded_update, sharded_update_exchange

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)

    case, scans, inits = make_inputs(N_SCANS)
    n_pts = scans[0].shape[0]
    cfg = api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=args.cell)
    m = api.Mapper(cfg, device=local)
    m.add(case.map_pts, 0.0)
    P0 = synth.default_P0()
    lim = np.zeros(23)

    # device-resident scans (float4 rows) and pinned host copies
    d_scans, h_scans = [], []
    for s in scans:
        s4 = np.zeros((n_pts, 4), np.float32)
        s4[:, :3] = s
        d_scans.append(torch.from_numpy(s4).cuda())
        h_scans.append(torch.from_numpy(s4).pin_memory())
    lo, hi = shard_bounds(n_pts, rank, world)
    red = torch.zeros(96, dtype=torch.float64, device="cuda")
    shm = None
    if world > 1 and args.exchange == "shm":
        def bcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        shm = attach_exchange(m, rank, world, bcast)
        dist.barrier()
    if world > 1 and args.exchange == "peer":
        handles = [None] * world
        dist.all_gather_object(handles, m.peer_export())
        m.peer_attach(rank, world, handles)
        dist.barrier()
    h_stream = torch.cuda.ExternalStream(m.stream())
    cur = torch.cuda.current_stream()

    def register(k, from_host):
        """One scan registration.  Returns the updated state."""
        if from_host:
            # this step's scan: H2D from pinned memory (already under way if the previous step prefetched it);
            # then start the copy of the NEXT step's scan so that it overlaps this registration
            # (N > 1: every rank uploads only ITS slice of the scan and binds it as its whole scan)
            m.set_scan_host(h_scans[k].data_ptr() + 16 * lo, hi - lo, 16)
            m.prefetch_scan_host(h_scans[(k + 1) % N_SCANS].data_ptr() + 16 * lo, hi - lo, 16)
        else:
            m.set_scan_device(d_scans[k].data_ptr(), n_pts, 16)
        if world == 1:
            x, P, passes = m.update(inits[k], P0, MAX_ITER, lim)
            return x, passes
        if not from_host:
            m.shard(lo, hi)
        if args.exchange == "peer":
            x, P, passes = m.update_peer(inits[k], P0, MAX_ITER, lim)
            return x, passes
        if shm is not None:
            x, P, passes = m.update_exchange(inits[k], P0, MAX_ITER, lim)
            return x, passes
        # Everything of the pass is ordered on the HANDLE's stream: the kernel is launched there
        # (stream NULL in the ABI = the handle's stream; the legacy default stream cannot be named) and
        # torch enqueues the NCCL all-reduce relative to the current stream, which we make that stream.
        with torch.cuda.stream(h_stream):
            def local_pass(state):
                m.match_async(state, red.data_ptr(), None)
                return red

            def all_reduce(t):
                dist.all_reduce(t)                     # 96 doubles: HTH tri + HTh + counters (NCCL)
                return t.cpu().numpy()

            x, P, passes = sharded_update(m, inits[k], P0, MAX_ITER, lim, local_pass, all_reduce)
        return x, passes

    def timed(steps, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(h_stream)
        x, passes = None, 0
        for s in range(steps):
            x, passes = register(s % N_SCANS, from_host)
        e1.record(h_stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), x, passes

    timed(args.warmup, False)
    timed(min(args.warmup, 3), True)
    st0 = m.stats()
    sampler = ClockSampler(local)
    sampler.start()
    ms, x, passes = timed(args.steps, False)
    st1 = m.stats()
    ms_e2e, x2, _ = timed(args.steps, True)
    clocks = sampler.stop()
    # Every 8th flimo_update runs with one launch per pass and CUDA events around the kernel.  A run with very few
    # steps may have none of those inside the timed region: time the kernel over 8 further scans right after it
    # (all ranks take the same branch: the call count is the same on every rank).
    st_k0, st_k1, kernel_where = st0, st1, "inside the timed region"
    if st1["match_timed"] - st0["match_timed"] == 0:
        st_k0 = m.stats()
        timed(8, False)
        st_k1 = m.stats()
        kernel_where = "over 8 further scans right after the timed region (steps < 8)"

    # sanity: the registration really converged onto the true pose of the last scan
    k_last = (args.steps - 1) % N_SCANS
    true_pos = inits[k_last][:3] - np.array([0.03, -0.03, 0.025])
    pose_err = float(np.abs(x[:3] - true_pos).max())

    if rank == 0:
        value = args.steps / (ms / 1e3)
        launches = st1["match_launches"] - st0["match_launches"]
        n_timed = st_k1["match_timed"] - st_k0["match_timed"]
        if n_timed == 0:
            raise SystemExit("no event-timed launch of the fused kernel (FLIMO_TIME_EVERY=0?)")
        k1_ms = (st_k1["match_ms_total"] - st_k0["match_ms_total"]) / n_timed
        pp = st_k1["persist_passes"] - st_k0["persist_passes"]
        k1_in_ms = (st_k1["persist_ms_total"] - st_k0["persist_ms_total"]) / pp if pp else None
        peak, peak_src = _peaks()
        achieved = (hi - lo) * A_PM / (k1_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arithmetic": "float32 per point (kNN, plane fit, residual, Jacobian row), float64 normal equations and filter algebra",
                       "parallelism": "scan-shard x%d (replicated map, 96 doubles summed per pass via %s)"
                       % (world, {"peer": "NVLink peer stores inside the registration kernel", "shm": "fused host-segment exchange",
                                   "nccl": "NCCL all-reduce"}[args.exchange] if world > 1 else "no exchange"),
                       "passes_per_scan": passes, "l2_policy": "inputs larger than L2 (multi-level map index %.1f GB, 4 rotating scans)"
                       % (st1["map_bytes"] / 1e9), "pose_err_m": pose_err, "knn_cell": st1["knn_cell"], "levels": st1["n_levels"]},
            "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": "scans/s", "h2d_bytes_per_step": n_pts * 16,   # summed over ranks
                    "d2h_bytes_per_step": passes * 96 * 8 + (26 + 529) * 8},
            "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "peak_source": peak_src, "kernel": "match_reduce_kernel",
                         "kernel_ms": k1_ms, "passes": int(launches), "passes_event_timed": int(n_timed), "kernel_timed": kernel_where,
                         "kernel_ms_in_persistent_kernel": k1_in_ms, "bytes_per_launch": (hi - lo) * A_PM,
                         "note": "kernel_ms = CUDA-event time of one-launch-per-pass executions (every 8th scan); the other scans run all "
                                 "passes inside one persistent launch, timed in-kernel with %globaltimer"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(case, scans, inits)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        m.close()
        if shm is not None:
            shm.close()
            if rank == 0:
                shm.unlink()
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
