"""Regenerates tests/golden/*.npz from the CPU oracle (run in the dev container):

    python tests/golden/make_golden.py

The reference itself ships no golden vectors and cannot be built here (Eigen/PCL/Boost absent), so
these fixtures pin the ORACLE's outputs on seeded inputs ("parity unpinned" w.r.t. the reference,
see DESIGN.md).  Inputs are regenerated from seeds by fast_limo_b200.synth; only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fast_limo_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def one(name, max_iter, **cfgkw):
    case = synth.make_case(name)
    om = O.OracleMap()
    om.add(case.map_pts)
    cfg = O.make_cfg(num_threads=4, **cfgkw)
    r = om.match(cfg, case.init[:14], case.scan, want_rows=True)
    x, P, tr = om.update(cfg, case.init, synth.default_P0(), max_iter, 0.0, case.scan)
    np.savez_compressed(
        os.path.join(HERE, f"{name}_m{cfgkw['max_matches']}.npz"),
        scan_checksum=np.float64(case.scan.astype(np.float64).sum()), map_checksum=np.float64(case.map_pts.astype(np.float64).sum()),
        good=np.packbits(r["good"]), n_q=np.int64(r["good"].shape[0]), plane=r["plane"][r["good"]], dist=r["dist"][r["good"]],
        nn_d2=r["nn_d2"], HTH=r["HTH"], HTh=r["HTh"], n_valid=np.int64(r["n_valid"]), rows=np.int64(r["rows"]),
        x_final=x, P_final=P, trace_states=np.stack([t["state"] for t in tr]), trace_rows=np.array([t["rows"] for t in tr]),
        max_iter=np.int64(max_iter))
    print(name, cfgkw, "n_valid", r["n_valid"], "rows", r["rows"], "passes", len(tr))


if __name__ == "__main__":
    one("tiny", 2, max_pc2match=1 << 18, max_matches=1 << 18)
    one("tiny", 3, max_pc2match=1500, max_matches=400)       # both first-N caps active
    one("c1", 0, max_pc2match=1 << 18, max_matches=1 << 18)  # BASELINE configs[0]: 16k scan, 100k map, 1 pass
