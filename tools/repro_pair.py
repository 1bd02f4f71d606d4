import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth
BIG = 1 << 18
case = synth.make_case("tiny")
for pair in sys.argv[1:]:
    os.environ["FLIMO_KNN_PAIR"] = pair
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    m.add(case.map_pts, 0.0); m.set_scan(case.scan)
    for rep in range(3):
        t = time.time()
        try:
            x, P, p = m.update(case.init, synth.default_P0(), 2, 0.0)
            print("pair", pair, "rep", rep, "passes", p, "dt %.3f" % (time.time() - t), "err", np.abs(x[:3] - case.truth[:3]).max(), flush=True)
        except Exception as e:
            print("pair", pair, "rep", rep, "FAILED", e, "dt %.3f" % (time.time() - t), flush=True)
    print(m.update_trace()[:, 26:30])
