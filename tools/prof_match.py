"""Runs a few match passes of a synthetic case (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sort = len(sys.argv) > 4 and sys.argv[4] == "sort"
case = synth.make_case(name)
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=cell, sort_scan=sort), device=0)
conv = "conv" in sys.argv
m.add(case.map_pts, 0.0)
m.set_scan(case.scan)
pose = case.init
if conv:   # the pose after a full update (what two of the three passes of a registration see)
    pose, _, _ = m.update(case.init, synth.default_P0(), 2, 0.0)
for i in range(reps):
    r = m.match(pose)
print("n_valid", r.n_valid, "ms", m.stats()["last_match_ms"])
