"""Per-warp timeline of one match pass (uses the flimo_debug_timing profiling hook)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
cell = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
sort = len(sys.argv) > 3 and sys.argv[3] == "sort"
case = synth.make_case(name)
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20, knn_cell=cell, sort_scan=sort), device=0)
m.add(case.map_pts, 0.0); m.set_scan(case.scan)
pose = case.init
if "conv" in sys.argv:
    pose, _, _ = m.update(case.init, synth.default_P0(), 2, 0.0)
frac = [int(a.split("=")[1]) for a in sys.argv if a.startswith("shard=")]
if frac:                      # keep only 1/frac of the scan: the per-warp latency chain at low occupancy
    m.shard(0, case.scan.shape[0] // frac[0])
for i in range(3): m.match(pose)
L = _lib.load()
nw = ((case.scan.shape[0] // frac[0] if frac else case.scan.shape[0]) + 127) // 128 * 4
buf = np.zeros((nw + 1, 8), np.uint64)
L.flimo_debug_timing.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
L.flimo_debug_timing(m._h, 1, None, nw)
m.match(pose)
L.flimo_debug_timing(m._h, 1, buf.ctypes.data, nw)
print("pass ms", m.stats()["last_match_ms"])
t_final = buf[nw, 0]
buf = buf[:nw]
print("final CTA done at us", (t_final - buf[:, 1].min()) / 1e3)
t0 = buf[:, 1].min()
st, knn, qr, acc = (buf[:, 1] - t0) / 1e3, (buf[:, 2] - buf[:, 1]) / 1e3, (buf[:, 3] - buf[:, 2]) / 1e3, (buf[:, 4] - buf[:, 3]) / 1e3
priv = (buf[:, 6] - buf[:, 1]) / 1e3
team = (buf[:, 2] - buf[:, 6]) / 1e3
maxcnt = (buf[:, 7] & np.uint64(0xFFFF)).astype(float)
probe = ((buf[:, 7] >> np.uint64(16)).astype(np.int64) - (buf[:, 1] & np.uint64((1 << 48) - 1)).astype(np.int64)) / 1e3   # lane 0's probe done
end = (buf[:, 4] - t0) / 1e3
q = lambda a: np.round(np.quantile(a, [0, 0.1, 0.5, 0.9, 0.99, 1.0]), 2)
print("start us  ", q(st)); print("knn us    ", q(knn)); print(" probe    ", q(probe)); print(" private  ", q(priv)); print(" teams    ", q(team)); print(" max cnt  ", q(maxcnt)); print("corr(priv,maxcnt)", np.corrcoef(priv, maxcnt)[0, 1]); print("qr us     ", q(qr)); print("acc us    ", q(acc)); print("end us    ", q(end))
print("escalated lanes per warp", q(buf[:, 5].astype(float)))
sm = buf[:, 0].astype(int)
per_sm_end = np.array([end[sm == s].max() if (sm == s).any() else 0 for s in range(148)])
per_sm_cnt = np.bincount(sm, minlength=148)
print("per-SM end us", q(per_sm_end), "warps per SM", q(per_sm_cnt.astype(float)))
slow = np.argsort(-end)[:8]
for w in slow: print(" slow warp", w, "sm", sm[w], "start %.1f knn %.1f qr %.1f acc %.1f esc %d" % (st[w], knn[w], qr[w], acc[w], buf[w, 5]))
print("corr(knn, esc)", np.corrcoef(knn, buf[:, 5].astype(float))[0, 1])
