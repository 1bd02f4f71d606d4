"""Why are cudaMalloc / cudaFree slow inside the library's process (tools/micro/alloc_latency.cu measures 0.3 ms in a bare
process, the index rebuild saw 100+ ms)?  Times a 32 MB cudaMalloc + cudaFree through the runtime the library uses, at several
points of a session."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth, _lib

L = _lib.load()
rt = None
for name in ("libcudart.so.12", "libcudart.so"):
    try:
        rt = C.CDLL(name); break
    except OSError:
        pass
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaFree.argtypes = [C.c_void_p]

def probe(tag):
    ts = []
    for _ in range(3):
        p = C.c_void_p()
        t0 = time.perf_counter(); rc1 = rt.cudaMalloc(C.byref(p), 32 << 20); t1 = time.perf_counter(); rc2 = rt.cudaFree(p); t2 = time.perf_counter()
        ts.append(((t1 - t0) * 1e6, (t2 - t1) * 1e6, rc1, rc2))
    print(f"{tag:40s} cudaMalloc/cudaFree us:", " ".join(f"{a:.0f}/{b:.0f}" for a, b, _, _ in ts), flush=True)

probe("fresh process (library loaded)")
case = synth.make_case("c1")
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20), device=0)
probe("after flimo_create")
m.add(case.map_pts[:50000], 0.0)
probe("after first Mapper::add")
m.set_scan(case.scan)
probe("after scan_set")
m.match(case.init)
probe("after one match (per-pass kernel)")
m.update(case.init, synth.default_P0(), 2, 0.0)
probe("after one update (resident kernels)")
m.add(case.map_pts[50000:52000], 1.0)
probe("after incremental add")
import torch
torch.zeros(1, device="cuda")
probe("after torch cuda init")
