"""Diagnostic for tests/test_gpu_configs.py::test_c3_stream_independent_oracle_chain: per-scan pose difference between the
device chain and the oracle's own chain, for several scan densities."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth
from oracle import oracle as O
BIG = 1 << 20

def rot_angle(qa, qb):
    return 2.0 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))

def run(rings, az, leaf, n_scans, handoff):
    S = synth.Stream(azimuths=az, rings=rings)
    m, om = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0), O.OracleMap()
    ocfg = O.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=O.max_threads())
    f = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=leaf, sensor_type=1)
    opc = O.make_prep_cfg(crop=([-1, -1, -1], [1, 1, 1]), min_dist=3.0, leaf=leaf, sensor_type=1)
    P0, lim, T = synth.default_P0(), np.full(23, 0.001), np.eye(4, dtype=np.float32)
    rng = np.random.default_rng(5)
    prev_end = 0.0
    for k in range(n_scans):
        raw, stamp = S.scan(k)
        n, t_last = m.prep_filter_sort(raw, stamp, f)
        frames = S.frames(prev_end, t_last)
        pred = S.state(t_last)
        pred[:3] += rng.normal(0, 0.02, 3)
        lq, lp = pred[3:7].astype(np.float32), pred[:3].astype(np.float32)
        m.prep_deskew(frames, lq, lp, T, 0.0)
        dev_pc = np.ascontiguousarray(m.prep_get(3)[:, :3])
        order = O.prep_filter_sort(raw, opc, sort=True)
        _, ob = O.prep_deskew(raw, order, opc, stamp, 0.0, frames, lq, lp, T)
        opc2 = np.ascontiguousarray(O.prep_voxel(ob, leaf)[:, :3])
        nd = len(dev_pc)
        if nd == len(opc2):
            dd = np.abs(dev_pc - opc2).max(axis=1)
            nd = "%d (same n; centroid diff max %.1e, #>1e-4: %d)" % (nd, dd.max(), int((dd > 1e-4).sum()))
        if k == 0:
            xg = xo = pred.copy()
            msg = ""
        else:
            xg, Pg, pg = m.update(pred, P0, 3, lim)
            xo, Po, tr = om.update(ocfg, pred, P0, 3, lim, opc2)
            msg = "dp %.2e dq %.2e  err_dev %.4f err_orc %.4f passes %d/%d" % (np.abs(xg[:3] - xo[:3]).max(), rot_angle(xg[3:7], xo[3:7]),
                  np.abs(xg[:3] - S.state(t_last)[:3]).max(), np.abs(xo[:3] - S.state(t_last)[:3]).max(), pg, len(tr))
        m.add_scan(xg, t_last)
        om.add(O.scan_to_world(xo[:14], opc2))        # the oracle's own transformPointCloud (reference float order)
        print(f"[{rings}x{az} leaf {leaf}] scan {k}: pc2match oracle {len(opc2)} dev {nd} map dev {m.size()} orc {om.size()} {msg}", flush=True)
        prev_end = t_last
    m.close()

for spec in (sys.argv[1] if len(sys.argv) > 1 else "32x512x0.5,64x1024x0.5,64x1024x0.25").split(","):
    r, a, l = spec.split("x")
    run(int(r), int(a), float(l), 10, False)
