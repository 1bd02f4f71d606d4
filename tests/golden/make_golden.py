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


def prep_inputs():
    """Seeded inputs of the scan-preparation fixture (shared with the tests)."""
    rng = np.random.default_rng(77)
    xyz = rng.uniform(-50, 50, (6000, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(-2, 7, 6000)
    ang = np.abs(np.arctan2(xyz[:, 1].astype(np.float64), xyz[:, 0].astype(np.float64)))
    xyz = xyz[np.abs(ang - 2.6) > 2e-6]                     # clear of the field-of-view decision boundary
    raw = synth.make_raw_message(xyz, sensor_type=1, sweep=0.1, stamp=100.0, seed=77, n_nan=30)
    frames = synth.make_frames(100.0, 100.1, rate_hz=200.0, speed=12.0, yaw_rate=1.2)
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = [0.27, -0.03, 0.4]
    T[:3, :3] = synth.quat_to_R(synth.quat_from_rpy(0.01, -0.02, 0.03)).astype(np.float32)
    return raw, frames, T


def prep():
    raw, frames, T = prep_inputs()
    cfg = O.make_prep_cfg(crop=([-1.5, -1.0, -1.0], [1.5, 1.0, 1.0]), min_dist=4.0, rate=2, fov=2.6, sensor_type=1, leaf=1.0)
    order = O.prep_filter_sort(raw, cfg, sort=True)
    t = O.prep_times(raw, order, cfg, 100.0)
    w, b = O.prep_deskew(raw, order, cfg, 100.0, -2.0e-4, frames, frames["q"][-2], frames["p"][-2], T)
    v = O.prep_voxel(b, 1.0)
    np.savez_compressed(os.path.join(HERE, "prep_velodyne.npz"), raw_checksum=np.float64(np.nansum(raw["x"].astype(np.float64))),
                        order=order, t_last=np.float64(t[-1]), world=w, xt2=b, voxel=v)
    print("prep", len(raw), "->", len(order), "->", len(v))


def imu_inputs():
    """Seeded inputs of the IMU-rate fixture (shared with the tests): state, covariance, 120 samples at 400 Hz."""
    from scipy.spatial.transform import Rotation as Rot
    rng = np.random.default_rng(4242)
    g = np.array([0.4, -0.3, -9.7])
    x0 = synth.make_state(rng.normal(0, 20, 3), Rot.from_rotvec([0.1, -0.2, 1.1]).as_quat(), Rot.from_rotvec([0.01, 0.02, -0.03]).as_quat(),
                          [0.27, -0.03, 0.4], [9.0, 2.0, -0.3], [0.002, -0.001, 0.003], [0.05, 0.02, -0.04], g * 9.809 / np.linalg.norm(g))
    A = rng.normal(size=(23, 23))
    P0 = 1e-3 * (A @ A.T / 23 + np.eye(23))
    n = 120
    stamps = 50.0 + (1 + np.arange(n)) / 400.0
    acc = (rng.normal(0, 2.0, (n, 3)) + [0, 0, 9.8]).astype(np.float32)
    gyr = rng.normal(0, 0.7, (n, 3)).astype(np.float32)
    return x0, P0, stamps, np.full(n, 1 / 400.0), acc, gyr


IMU_COV = (6.e-4, 1.e-2, 1.e-5, 3.e-4)
IMU_WINDOW = (50.1001, 50.2001)


def imu():
    x0, P0, stamps, dts, acc, gyr = imu_inputs()
    pr = O.Propagator(x0, P0)
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        pr.propagate(t, dt, a, w, IMU_COV)
    x, P = pr.get()
    fr = pr.frames(*IMU_WINDOW)
    np.savez_compressed(os.path.join(HERE, "imu_predict.npz"), x=x, P=P, frames=fr.view(np.uint8), n_frames=np.int32(len(fr)))
    print("imu", len(stamps), "samples ->", len(fr), "frames in the window")


def ref_octree_inputs():
    """Seeded batches (dense enough to trigger the drop rule and leaf splits, one batch grows the root) and
    queries for the reference-octree fixture."""
    rng = np.random.default_rng(4242)
    case = synth.make_case("tiny")
    pts = case.map_pts
    batches = [pts[:8000]]
    centres = [pts[rng.integers(len(pts))] for _ in range(3)]
    for i in range(8):
        centre = centres[i % 3]                                       # revisited patches: the drop rule fires
        dense = (centre + rng.normal(0, [1.5, 1.5, 0.02], (2500, 3))).astype(np.float32)          # a dense patch: > 4 points per min-level cell
        sparse = pts[rng.choice(len(pts), 1500, replace=False)] + rng.normal(0, 0.03, (1500, 3)).astype(np.float32)
        b = np.concatenate([dense, sparse]).astype(np.float32)
        if i == 3:
            b = (b + np.array([70.0, -45.0, 2.0], np.float32)).astype(np.float32)                    # outside the first root
        batches.append(b)
    batches.append(np.concatenate([batches[2][:50], np.full((3, 3), np.nan, np.float32)]))        # NaN points are skipped
    queries = np.concatenate([case.scan[:3000] + np.float32(0.05), batches[4][:500] + np.float32(0.01)]).astype(np.float32)
    return batches, queries


def contents_checksum(pts):
    """Order-independent fingerprint of a point set, EXACT in any summation order: the count and wrap-around
    uint64 sums of the coordinate bit patterns and of a per-point mix."""
    p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    b = p.view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        mix = b[:, 0] * np.uint64(0x9E3779B185EBCA87) + b[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) + b[:, 2] * np.uint64(0x165667B19E3779F9)
        return np.array([np.uint64(len(p)), b[:, 0].sum(dtype=np.uint64), b[:, 1].sum(dtype=np.uint64), b[:, 2].sum(dtype=np.uint64),
                         mix.sum(dtype=np.uint64)], np.uint64)


def ref_octree():
    """Golden vectors produced by the REFERENCE's own octree (oracle/_ref, Octree.hpp compiled as it lies):
    map size after every batch, the final contents, exact 5-NN distances / neighbours of the queries."""
    if not O.ref_available():
        print("oracle/_ref/libref_octree.so missing (make -C oracle ref; needs /root/reference): fixture not regenerated")
        return
    for tag, me, ds in (("a", 0.2, True), ("b", 0.35, True), ("c", 0.2, False)):
        batches, queries = ref_octree_inputs()
        ro = O.RefOctree(bucket=2, min_extent=me, downsample=ds)        # bucket_size 2 from the YAML: ignored by the reference (D5)
        sizes = []
        for b in batches:
            ro.add(b)
            sizes.append(ro.size())
        pts = ro.points()
        pts = pts[np.lexsort(pts.T)]
        xyz, d2, cnt = ro.knn(queries, 5)
        np.savez_compressed(os.path.join(HERE, f"ref_octree_{tag}.npz"), min_extent=np.float32(me), downsample=np.int32(ds),
                            sizes=np.array(sizes, np.int64), contents_checksum=contents_checksum(pts), knn_d2=d2,
                            knn_xyz_checksum=np.float64(xyz.astype(np.float64).sum()), knn_cnt=cnt.astype(np.int8),
                            in_checksum=np.float64(sum(float(np.nansum(b.astype(np.float64))) for b in batches)))
        print("ref_octree", tag, sizes)


if __name__ == "__main__":
    ref_octree()
    prep()
    imu()
    one("tiny", 2, max_pc2match=1 << 18, max_matches=1 << 18)
    one("tiny", 3, max_pc2match=1500, max_matches=400)       # both first-N caps active
    one("c1", 0, max_pc2match=1 << 18, max_matches=1 << 18)  # BASELINE configs[0]: 16k scan, 100k map, 1 pass
