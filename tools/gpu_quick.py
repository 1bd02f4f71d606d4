"""Scratch GPU check: builds a case on the device, times the match pass, compares with the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cells = [float(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0.25]
check = "--check" in sys.argv
t = time.time(); case = synth.make_case(name); print("gen", name, round(time.time() - t, 2), "s", case.map_pts.shape, case.scan.shape, flush=True)
for sort in (False, True):
    for cell in cells:
        cfg = api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=cell, sort_scan=sort)
        m = api.Mapper(cfg, device=0)
        t = time.time(); m.add(case.map_pts, 0.0); tb = time.time() - t
        m.set_scan(case.scan)
        ts = []
        for i in range(6):
            r = m.match(case.init); ts.append(m.stats()["last_match_ms"])
        st = m.stats()
        n = case.scan.shape[0]
        best = min(ts[1:])
        print(f"cell={cell} sort={sort} build={tb*1e3:.1f}ms grid={st['grid_nx']}x{st['grid_ny']}x{st['grid_nz']} table={st['table_bytes']/1e6:.0f}MB "
              f"match_ms={best:.4f} (all {['%.3f'%x for x in ts]}) n_valid={r.n_valid} alg_GBs={n*528/best/1e6:.0f}", flush=True)
        t = time.time(); x, P, passes = m.update(case.init, synth.default_P0(), 2, 0.0); tu = time.time() - t
        print(f"   update passes={passes} wall={tu*1e3:.2f}ms pose_err={np.abs(x[:3]-case.truth[:3]).max():.5f}", flush=True)
        m.close()
if check:
    from oracle import oracle as O
    om = O.OracleMap(); t = time.time(); om.add(case.map_pts); print("oracle build", round(time.time() - t, 2), flush=True)
    th = O.max_threads()
    ocfg = O.make_cfg(max_pc2match=1 << 20, max_matches=1 << 20, num_threads=th)
    t = time.time(); ref = om.match(ocfg, case.init[:14], case.scan); to = time.time() - t
    print(f"oracle match {to*1e3:.1f} ms on {th} threads n_valid={ref['n_valid']}", flush=True)
    cfg = api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=cells[0])
    m = api.Mapper(cfg, device=0); m.add(case.map_pts, 0.0); m.set_scan(case.scan)
    dbg = m.match_debug(case.init)
    g = ref["good"]
    print("good equal:", np.array_equal(dbg["good"], g), "mismatch", int((dbg["good"] != g).sum()))
    both = g & dbg["good"]
    print("plane bit-equal:", np.array_equal(dbg["plane"][both], ref["plane"][both]), "dist bit-equal:", np.array_equal(dbg["dist"][both], ref["dist"][both]))
    close = ref["nn_d2"][:, 4] < 2.0
    print("nn_d2 bit-equal (where d5<2):", np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close]), "world equal:", np.array_equal(dbg["world"], ref["world"]))
    r = m.match(case.init)
    print("HTH rel err", np.abs(r.HTH - ref["HTH"]).max() / np.abs(ref["HTH"]).max(), "n_valid", r.n_valid, ref["n_valid"])
