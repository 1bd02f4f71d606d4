"""Device scan preparation (filters + time sort + deskew + voxel grid) timing next to the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fast_limo_b200 import api, synth
from oracle import oracle as O
case = synth.make_case("c2")
sub = int(sys.argv[1]) if len(sys.argv) > 1 else 1            # keep every sub-th point (4 -> the 32 k points of config c5)
raw = synth.make_raw_message(case.scan[::sub], sensor_type=1, rings=64)
import torch
raw_pinned = torch.from_numpy(raw.view(np.uint8).reshape(-1)).pin_memory().numpy().view(raw.dtype)
m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=1 << 20, MAX_NUM_PC2MATCH=1 << 20), device=0)
frames = synth.make_frames(100.0, 100.1, rate_hz=400.0)
lq, lp = frames["q"][-2], frames["p"][-2]
for leaf in (None, 0.5):
    f = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, fov_angle=3.0, leafSize=leaf, sensor_type=1)
    for _ in range(3):
        n, tl = m.prep_filter_sort(raw, 100.0, f); nv = m.prep_deskew(frames, lq, lp, np.eye(4), 0.0)
    tp0 = time.perf_counter()
    for _ in range(50):
        n, tl = m.prep_filter_sort(raw_pinned, 100.0, f)
    tp1 = time.perf_counter()
    t0 = time.perf_counter()
    for _ in range(50):
        n, tl = m.prep_filter_sort(raw, 100.0, f)
    t1 = time.perf_counter()
    for _ in range(50):
        nv = m.prep_deskew(frames, lq, lp, np.eye(4), 0.0)
    t2 = time.perf_counter()
    oc = O.make_prep_cfg(crop=([-1, -1, -1], [1, 1, 1]), min_dist=3.0, fov=3.0, sensor_type=1, leaf=leaf)
    c0 = time.perf_counter()
    order = O.prep_filter_sort(raw, oc, sort=True)
    c1 = time.perf_counter()
    w, b = O.prep_deskew(raw, order, oc, 100.0, 0.0, frames, lq, lp, np.eye(4))
    if leaf: v = O.prep_voxel(b, leaf)
    c2 = time.perf_counter()
    print(f"leaf={leaf}: raw {len(raw)} -> kept {n} -> pc2match {nv} | GPU filter+sort {(t1-t0)/50*1e6:.0f} us (incl. {len(raw)*32/1e6:.1f} MB H2D from pageable memory; {(tp1-tp0)/50*1e6:.0f} us from pinned), "
          f"deskew{'+voxel' if leaf else ''} {(t2-t1)/50*1e6:.0f} us | CPU oracle (1 thread) filter+sort {(c1-c0)*1e3:.1f} ms, deskew{'+voxel' if leaf else ''} {(c2-c1)*1e3:.1f} ms")
