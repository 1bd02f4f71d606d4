"""Tuning sweep of the index ladder (finest cell, level ratio, tau) on one case; prints the fused-kernel time
per configuration and checks every configuration's per-point results against the first one (and,
with --check, against the CPU oracle)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
specs = sys.argv[2] if len(sys.argv) > 2 else "0.15:1.4142:24"
check = "--check" in sys.argv
case = synth.make_case(name)
ref = None
if check:
    from oracle import oracle as O
    om = O.OracleMap(); om.add(case.map_pts)
    ocfg = O.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20, num_threads=O.max_threads())
    ref = om.match(ocfg, case.init[:14], case.scan)
first = None
for spec in specs.split(","):
    cell, ratio, tau = spec.split(":")[:3]
    for kv in spec.split(":")[3:]:
        k, v = kv.split("=")
        os.environ["FLIMO_KNN_" + k.upper()] = v
    cfg = api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=float(cell), knn_level_ratio=float(ratio), knn_tau=int(tau))
    m = api.Mapper(cfg, device=0)
    t = time.time(); m.add(case.map_pts, 0.0); tb = time.time() - t
    m.set_scan(case.scan)
    ts = []
    for i in range(8):
        r = m.match(case.init); ts.append(m.stats()["last_match_ms"])
    x, Pm, passes = m.update(case.init, synth.default_P0(), 2, 0.0)
    tc = []
    for i in range(6):
        m.match(x); tc.append(m.stats()["last_match_ms"])
    conv = min(tc[1:])
    s0 = m.stats()
    for i in range(20): m.update(case.init, synth.default_P0(), 2, 0.0)
    s1 = m.stats()
    avg3 = (s1["match_ms_total"] - s0["match_ms_total"]) / max(s1["match_timed"] - s0["match_timed"], 1)
    st = m.stats()
    n = case.scan.shape[0]
    best, med = min(ts[1:]), float(np.median(ts[1:]))
    dbg = m.match_debug(case.init)
    sig = (int(dbg["good"].sum()), float(dbg["dist"][dbg["good"]].astype(np.float64).sum()))
    if first is None: first = dbg
    same = np.array_equal(dbg["good"], first["good"]) and np.array_equal(dbg["plane"][dbg["good"]], first["plane"][first["good"]])
    msg = ""
    if ref is not None:
        g = ref["good"]
        close = ref["nn_d2"][:, 4] < 2.0
        msg = " oracle: good=%s plane=%s dist=%s nn=%s" % (np.array_equal(dbg["good"], g), np.array_equal(dbg["plane"][g], ref["plane"][g]),
              np.array_equal(dbg["dist"][g], ref["dist"][g]), np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close]))
    fl, ll = dbg["levels"] % 16, dbg["levels"] // 16
    diag = (f" first_lvl_hist={np.bincount(fl, minlength=st['n_levels']).tolist()} rescanned={float((ll != fl).mean()):.3f} "
            f"cnt q10/50/90/99={np.quantile(dbg['first_count'], [0.1, 0.5, 0.9, 0.99]).astype(int).tolist()} over_cap={float((dbg['first_count'] > 96).mean()):.4f}")
    print(f"cell={cell} ratio={ratio} tau={tau} levels={st['n_levels']} build={tb*1e3:.0f}ms map={st['map_bytes']/1e9:.1f}GB table={st['table_bytes']/1e9:.2f}GB "
          f"match init best={best*1e3:.1f}us med={med*1e3:.1f}us conv={conv*1e3:.1f}us avg3={avg3*1e3:.1f}us n_valid={r.n_valid} same_as_first={same}{msg}{diag}", flush=True)
    m.close()
