"""Wall-clock breakdown of one scan registration through the C ABI (host timers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fast_limo_b200 import api, synth
case = synth.make_case("c2")
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20), device=0)
m.add(case.map_pts, 0.0)
s4 = np.zeros((case.scan.shape[0], 4), np.float32); s4[:, :3] = case.scan
d = torch.from_numpy(s4).cuda(); hp = torch.from_numpy(s4).pin_memory()
P0 = synth.default_P0(); lim = np.zeros(23)
def T(f, n=200):
    for _ in range(10): f()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
n = case.scan.shape[0]
print("set_scan_device (pack+sort)  us", round(T(lambda: m.set_scan_device(d.data_ptr(), n, 16)), 1))
print("set_scan host pinned         us", round(T(lambda: m._ck(m._L.flimo_scan_set(m._h, hp.data_ptr(), n, 16))), 1))
m.set_scan_device(d.data_ptr(), n, 16)
print("match (blocking, 1 pass)     us", round(T(lambda: m.match(case.init)), 1), " kernel us", round(m.stats()["last_match_ms"] * 1e3, 1))
HTH = np.eye(12) * 1e4; HTh = np.ones(12)
def ekf():
    m.ekf_begin(case.init, P0, 2, lim); m.ekf_step(HTH, HTh, 1000); m.ekf_end()
print("ekf begin+1 step+end (host)  us", round(T(ekf), 1))
def ekf3():
    m.ekf_begin(case.init, P0, 2, lim)
    for _ in range(3): m.ekf_step(HTH, HTh, 1000)
    m.ekf_end()
print("ekf begin+3 steps+end (host) us", round(T(ekf3), 1))
print("update (3 passes)            us", round(T(lambda: m.update(case.init, P0, 2, lim)), 1))
def full():
    m.set_scan_device(d.data_ptr(), n, 16); m.update(case.init, P0, 2, lim)
print("set_scan_device + update     us", round(T(full), 1))
