"""Per-phase device timeline of the filter step inside the registration kernel (FLIMO_PROFILE=1 prints it)."""
import os, sys
os.environ["FLIMO_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
case = synth.make_case(name)
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20), device=0)
m.add(case.map_pts, 0.0)
m.set_scan(case.scan)
for i in range(5):
    x, P, passes = m.update(case.init, synth.default_P0(), 2, 0.0)
tr = m.update_trace()
print("passes", passes, "pass ns", tr[:, 28], "pose err", np.abs(x[:3] - case.truth[:3]).max())
