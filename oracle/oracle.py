"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over ``liboracle.so`` (the CPU restatement of the reference's registration
path: i-Octree kNN, plane fit, Jacobian, HtH, IKFoM iterated update).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product package never does.

Parity status: the octree (kNN, insert) is PINNED against the reference's own Octree.hpp compiled
unmodified (oracle/_ref, `make -C oracle ref`, tests/test_ref_octree.py); everything that needs real
Eigen / PCL (plane QR, Jacobian, IKFoM update, deskew algebra, voxel grid) is UNPINNED — the reference
ships no tests/golden vectors and cannot be compiled here as a whole; see the headers of ioctree.hpp /
plane_match.hpp / ekf.hpp / prep.hpp for the reference file:line each function follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TRACE_STRIDE = 26 + 23 + 1 + 144 + 12


class OrcCfg(C.Structure):
    _fields_ = [
        ("k", C.c_int32),
        ("estimate_extrinsics", C.c_int32),
        ("num_threads", C.c_int32),
        ("_pad", C.c_int32),
        ("max_pc2match", C.c_int64),
        ("max_matches", C.c_int64),
        ("max_dist_plane", C.c_double),
        ("plane_threshold", C.c_double),
    ]


def build(force=False):
    """Compile liboracle.so from the sources in this directory (g++, a few seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "ioctree.hpp", "plane_match.hpp", "ekf.hpp", "prep.hpp", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        f32p, f64p, u8p, i32p, i64p = (C.POINTER(t) for t in (C.c_float, C.c_double, C.c_uint8, C.c_int32, C.c_int64))
        L.orc_max_threads.restype = C.c_int
        L.orc_map_new.restype = C.c_void_p
        L.orc_map_new.argtypes = [C.c_int, C.c_float, C.c_int]
        L.orc_map_free.argtypes = [C.c_void_p]
        L.orc_map_add.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_size_t]
        L.orc_map_size.restype = C.c_size_t
        L.orc_map_size.argtypes = [C.c_void_p]
        L.orc_map_dump.restype = C.c_size_t
        L.orc_map_dump.argtypes = [C.c_void_p, f32p, C.c_size_t]
        L.orc_knn.argtypes = [C.c_void_p, f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_int, f32p, f32p, i32p]
        L.orc_plane_fit.argtypes = [f32p, C.c_int, f32p]
        L.orc_match.restype = C.c_long
        L.orc_match.argtypes = [C.c_void_p, C.POINTER(OrcCfg), f64p, f32p, C.c_size_t, C.c_size_t,
                                u8p, f32p, f32p, f32p, f32p, f64p, f64p, f64p, f64p, i64p]
        L.orc_scan_to_world.argtypes = [f64p, f32p, C.c_size_t, C.c_size_t, f32p]
        L.orc_update.restype = C.c_int
        L.orc_update.argtypes = [C.c_void_p, C.POINTER(OrcCfg), f64p, f64p, C.c_int, f64p, C.c_double, C.c_double,
                                 f32p, C.c_size_t, C.c_size_t, f64p, C.c_int]
        L.orc_update_fixed.restype = C.c_int
        L.orc_update_fixed.argtypes = [f64p, f64p, C.c_int, f64p, C.c_double, C.c_double, f64p, f64p, C.c_long, f64p, C.c_int]
        L.orc_boxplus.argtypes = [f64p, f64p]
        L.orc_boxminus.argtypes = [f64p, f64p, f64p]
        L.orc_invert.restype = C.c_int
        L.orc_invert.argtypes = [f64p, C.c_int]
        _LIB = L
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def max_threads():
    return int(lib().orc_max_threads())


def make_cfg(k=5, max_pc2match=10000, max_matches=2000, max_dist_plane=2.0, plane_threshold=0.05,
             estimate_extrinsics=True, num_threads=1):
    return OrcCfg(k, int(bool(estimate_extrinsics)), int(num_threads), 0, int(max_pc2match), int(max_matches),
                  float(max_dist_plane), float(plane_threshold))


def _xyz(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


class OracleMap:
    """The reference's Mapper + i-Octree (bucket 32 effective)."""

    def __init__(self, min_extent=0.2, downsample=True, bucket=32):
        self._h = lib().orc_map_new(int(bucket), float(min_extent), int(bool(downsample)))

    def __del__(self):
        try:
            if self._h:
                lib().orc_map_free(self._h)
                self._h = None
        except Exception:
            pass

    def add(self, pts):
        pts = _xyz(pts)
        lib().orc_map_add(self._h, _p(pts, C.c_float), pts.shape[0], pts.shape[1])

    def size(self):
        return int(lib().orc_map_size(self._h))

    def points(self):
        n = self.size()
        out = np.empty((max(n, 1), 3), np.float32)
        got = lib().orc_map_dump(self._h, _p(out, C.c_float), out.shape[0])
        return out[:got].copy()

    def knn(self, q, k=5, threads=1):
        q = _xyz(q)
        nq = q.shape[0]
        d2 = np.empty((nq, k), np.float32)
        nb = np.empty((nq, k, 3), np.float32)
        cnt = np.empty(nq, np.int32)
        lib().orc_knn(self._h, _p(q, C.c_float), nq, q.shape[1], k, threads, _p(d2, C.c_float), _p(nb, C.c_float), _p(cnt, C.c_int32))
        return d2, nb, cnt

    def match(self, cfg, state14, scan, want_rows=False):
        """One h_share_model pass.  Returns dict with per-query arrays, HTH, HTh, n_valid, rows."""
        scan = _xyz(scan)
        n = scan.shape[0]
        nq = min(n, cfg.max_pc2match)
        st = np.ascontiguousarray(state14, np.float64)
        good = np.zeros(nq, np.uint8)
        plane = np.zeros((nq, 4), np.float32)
        dist = np.zeros(nq, np.float32)
        world = np.zeros((nq, 3), np.float32)
        nn = np.zeros((nq, cfg.k), np.float32)
        HTH = np.zeros((12, 12), np.float64)
        HTh = np.zeros(12, np.float64)
        nv = C.c_int64(0)
        H = h = None
        if want_rows:
            cap = min(nq, cfg.max_matches)
            H = np.zeros((cap + 1, 12), np.float64)
            h = np.zeros(cap + 1, np.float64)
        rows = lib().orc_match(self._h, C.byref(cfg), _p(st, C.c_double), _p(scan, C.c_float), n, scan.shape[1],
                               _p(good, C.c_uint8), _p(plane, C.c_float), _p(dist, C.c_float), _p(world, C.c_float),
                               _p(nn, C.c_float), _p(H, C.c_double), _p(h, C.c_double), _p(HTH, C.c_double),
                               _p(HTh, C.c_double), C.byref(nv))
        out = dict(good=good.astype(bool), plane=plane, dist=dist, world=world, nn_d2=nn, HTH=HTH, HTh=HTh,
                   n_valid=int(nv.value), rows=int(rows))
        if want_rows:
            out["H"] = H[:rows]
            out["h"] = h[:rows]
        return out

    def update(self, cfg, state26, P, max_iter, limits, scan, R=0.001, D=5.0):
        """esekf::update_iterated_dyn_share_modified with the reference measurement model."""
        scan = _xyz(scan)
        st = np.array(state26, np.float64).copy()
        Pm = np.array(P, np.float64).reshape(23, 23).copy()
        lim = np.ascontiguousarray(np.broadcast_to(np.asarray(limits, np.float64), (23,)))
        trace = np.zeros((max_iter + 2, TRACE_STRIDE), np.float64)
        passes = lib().orc_update(self._h, C.byref(cfg), _p(st, C.c_double), _p(Pm, C.c_double), int(max_iter),
                                  _p(lim, C.c_double), float(R), float(D), _p(scan, C.c_float), scan.shape[0],
                                  scan.shape[1], _p(trace, C.c_double), trace.shape[0])
        return st, Pm, unpack_trace(trace[:passes])


def scan_to_world(state14, scan):
    """pcl::transformPointCloud(pc2match, final_scan, state.get_RT()) (Localizer.cpp:361): (n, 3) float32 world points."""
    scan = _xyz(scan)
    st = np.ascontiguousarray(state14, np.float64)
    out = np.zeros((scan.shape[0], 3), np.float32)
    lib().orc_scan_to_world(_p(st, C.c_double), _p(scan, C.c_float), scan.shape[0], scan.shape[1], _p(out, C.c_float))
    return out


def unpack_trace(tr):
    return [dict(state=t[:26].copy(), dx=t[26:49].copy(), rows=int(t[49]), HTH=t[50:194].reshape(12, 12).copy(),
                 HTh=t[194:206].copy()) for t in tr]


def update_fixed(state26, P, max_iter, limits, H, h, R=0.001, D=5.0):
    st = np.array(state26, np.float64).copy()
    Pm = np.array(P, np.float64).reshape(23, 23).copy()
    lim = np.ascontiguousarray(np.broadcast_to(np.asarray(limits, np.float64), (23,)))
    H = np.ascontiguousarray(H, np.float64).reshape(-1, 12)
    h = np.ascontiguousarray(h, np.float64)
    trace = np.zeros((max_iter + 2, TRACE_STRIDE), np.float64)
    passes = lib().orc_update_fixed(_p(st, C.c_double), _p(Pm, C.c_double), int(max_iter), _p(lim, C.c_double), float(R),
                                    float(D), _p(H, C.c_double), _p(h, C.c_double), H.shape[0], _p(trace, C.c_double),
                                    trace.shape[0])
    return st, Pm, unpack_trace(trace[:passes])


def plane_fit(pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros(4, np.float32)
    lib().orc_plane_fit(_p(pts, C.c_float), pts.shape[0], _p(out, C.c_float))
    return out


def boxplus(state26, d23):
    st = np.array(state26, np.float64).copy()
    d = np.ascontiguousarray(d23, np.float64)
    lib().orc_boxplus(_p(st, C.c_double), _p(d, C.c_double))
    return st


def boxminus(a26, b26):
    a = np.ascontiguousarray(a26, np.float64)
    b = np.ascontiguousarray(b26, np.float64)
    d = np.zeros(23, np.float64)
    lib().orc_boxminus(_p(a, C.c_double), _p(b, C.c_double), _p(d, C.c_double))
    return d


def invert(A):
    A = np.array(A, np.float64).copy()
    rc = lib().orc_invert(_p(A, C.c_double), A.shape[0])
    assert rc == 0
    return A


# ---- scan preparation (prep.hpp): filters, time sort, deskew, voxel grid ---------------------------
RAW_POINT = np.dtype({"names": ["x", "y", "z", "intensity", "t", "time", "timestamp"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<f4", "<f8"],
                      "offsets": [0, 4, 8, 16, 24, 24, 24], "itemsize": 32})      # fast_limo::Point (Common.hpp:100-113)
FRAME = np.dtype([("time", "<f8"), ("q", "<f4", 4), ("p", "<f4", 3), ("v", "<f4", 3), ("w", "<f4", 3), ("a", "<f4", 3),
                  ("bg", "<f4", 3), ("ba", "<f4", 3), ("g", "<f4", 3)], align=True)  # fast_limo::State members used


class PrepCfg(C.Structure):
    _fields_ = [("crop_active", C.c_int32), ("dist_active", C.c_int32), ("rate_active", C.c_int32), ("fov_active", C.c_int32),
                ("crop_min", C.c_float * 3), ("crop_max", C.c_float * 3), ("min_dist", C.c_double), ("rate_value", C.c_int32),
                ("fov_angle", C.c_float), ("sensor_type", C.c_int32), ("end_of_sweep", C.c_int32), ("voxel_active", C.c_int32),
                ("leaf", C.c_float)]


def make_prep_cfg(crop=None, min_dist=None, rate=None, fov=None, sensor_type=1, end_of_sweep=False, leaf=None):
    c = PrepCfg()
    if crop is not None:
        c.crop_active = 1
        c.crop_min[:] = [float(v) for v in crop[0]]
        c.crop_max[:] = [float(v) for v in crop[1]]
    if min_dist is not None:
        c.dist_active, c.min_dist = 1, float(min_dist)
    c.rate_value = 1
    if rate is not None:
        c.rate_active, c.rate_value = 1, int(rate)
    if fov is not None:
        c.fov_active, c.fov_angle = 1, float(fov)
    c.sensor_type, c.end_of_sweep = int(sensor_type), int(bool(end_of_sweep))
    if leaf is not None:
        c.voxel_active, c.leaf = 1, float(leaf)
    return c


def prep_filter_sort(raw, cfg, sort=True):
    L = lib()
    L.orc_prep_filter_sort.restype = C.c_size_t
    L.orc_prep_filter_sort.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(PrepCfg), C.c_int, C.c_void_p]
    raw = np.ascontiguousarray(raw)
    out = np.zeros(max(len(raw), 1), np.uint32)
    m = L.orc_prep_filter_sort(raw.ctypes.data, len(raw), C.byref(cfg), int(sort), out.ctypes.data)
    return out[:m].copy()


def prep_times(raw, order, cfg, sweep_ref_time):
    L = lib()
    L.orc_prep_times.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(PrepCfg), C.c_double, C.c_void_p]
    raw, order = np.ascontiguousarray(raw), np.ascontiguousarray(order, np.uint32)
    t = np.zeros(len(order), np.float64)
    L.orc_prep_times(raw.ctypes.data, order.ctypes.data, len(order), C.byref(cfg), float(sweep_ref_time), t.ctypes.data)
    return t


def prep_deskew(raw, order, cfg, sweep_ref_time, offset, frames, last_q, last_p, T_l2b):
    """Returns (world xyz1, Xt2-frame xyz1), float32 (m, 4) each."""
    L = lib()
    L.orc_prep_deskew.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(PrepCfg), C.c_double, C.c_double, C.c_void_p,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    raw, order = np.ascontiguousarray(raw), np.ascontiguousarray(order, np.uint32)
    frames = np.ascontiguousarray(frames, FRAME)
    lq, lp = np.ascontiguousarray(last_q, np.float32), np.ascontiguousarray(last_p, np.float32)
    T = np.ascontiguousarray(T_l2b, np.float32).reshape(16)
    w = np.zeros((len(order), 4), np.float32)
    b = np.zeros((len(order), 4), np.float32)
    L.orc_prep_deskew(raw.ctypes.data, order.ctypes.data, len(order), C.byref(cfg), float(sweep_ref_time), float(offset),
                      frames.ctypes.data, len(frames), lq.ctypes.data, lp.ctypes.data, T.ctypes.data, w.ctypes.data, b.ctypes.data)
    return w, b


def prep_voxel(pts4, leaf):
    L = lib()
    L.orc_prep_voxel.restype = C.c_size_t
    L.orc_prep_voxel.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
    pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
    out = np.zeros((max(len(pts4), 1), 4), np.float32)
    m = L.orc_prep_voxel(pts4.ctypes.data, len(pts4), float(leaf), out.ctypes.data)
    return out[:m].copy()


# ---- IMU rate (predict.hpp) ------------------------------------------------------------------------------
class Propagator:
    """Filter state + propagated_buffer: Localizer::propagateImu / integrateImu restated (oracle/predict.hpp)."""

    def __init__(self, state26, P):
        L = lib()
        L.orc_prop_new.restype = C.c_void_p
        L.orc_prop_new.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_prop_free.argtypes = [C.c_void_p]
        L.orc_prop_propagate.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_prop_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_prop_frames.restype = C.c_long
        L.orc_prop_frames.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t]
        x = np.ascontiguousarray(state26, np.float64).reshape(26)
        Pm = np.ascontiguousarray(np.asarray(P, np.float64).reshape(23, 23))
        self._L = L
        self._p = L.orc_prop_new(x.ctypes.data, Pm.ctypes.data)

    def __del__(self):
        if getattr(self, "_p", None):
            self._L.orc_prop_free(self._p)
            self._p = None

    def propagate(self, stamp, dt, lin_accel, ang_vel, cov=(6.e-4, 1.e-2, 1.e-5, 3.e-4)):
        a, w = np.ascontiguousarray(lin_accel, np.float32), np.ascontiguousarray(ang_vel, np.float32)
        c4 = np.ascontiguousarray(cov, np.float64)
        self._L.orc_prop_propagate(self._p, float(stamp), float(dt), a.ctypes.data, w.ctypes.data, c4.ctypes.data)

    def get(self):
        x, Pm = np.zeros(26, np.float64), np.zeros((23, 23), np.float64)
        self._L.orc_prop_get(self._p, x.ctypes.data, Pm.ctypes.data)
        return x, Pm

    def frames(self, start_time, end_time):
        """None: the reference would block waiting for IMU data; else FRAME records (possibly empty), oldest first."""
        n = self._L.orc_prop_frames(self._p, float(start_time), float(end_time), None, 0)
        if n < 0:
            return None
        out = np.zeros(n, FRAME)
        if n:
            self._L.orc_prop_frames(self._p, float(start_time), float(end_time), out.ctypes.data, n)
        return out


def process_model(state26, acc, gyro):
    """get_f (24), df_dx (24 x 23), df_dw (24 x 12) of use-ikfom.cpp:46-91."""
    L = lib()
    L.orc_process_model.argtypes = [C.c_void_p] * 6
    x = np.ascontiguousarray(state26, np.float64).reshape(26)
    a, w = np.ascontiguousarray(acc, np.float64), np.ascontiguousarray(gyro, np.float64)
    f, fx, fw = np.zeros(24), np.zeros((24, 23)), np.zeros((24, 12))
    L.orc_process_model(x.ctypes.data, a.ctypes.data, w.ctypes.data, f.ctypes.data, fx.ctypes.data, fw.ctypes.data)
    return f, fx, fw


# ---- oracle/_ref: the REFERENCE's own octree, compiled where it lies (make -C oracle ref) -----------------
REF_LIB = os.path.join(_HERE, "_ref", "libref_octree.so")


def ref_available():
    return os.path.exists(REF_LIB)


class RefOctree:
    """fast_limo::octree::Octree of the reference (Octree.hpp, unmodified), driven like Mapper drives it."""

    def __init__(self, bucket=2, min_extent=0.2, downsample=True):
        L = self.L = C.CDLL(REF_LIB)
        L.ref_octree_new.restype = C.c_void_p
        L.ref_octree_new.argtypes = [C.c_int, C.c_float, C.c_int]
        L.ref_octree_free.argtypes = [C.c_void_p]
        L.ref_octree_update.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_octree_size.restype = C.c_size_t
        L.ref_octree_size.argtypes = [C.c_void_p]
        L.ref_octree_dump.restype = C.c_size_t
        L.ref_octree_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ref_octree_knn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.ref_octree_new(int(bucket), float(min_extent), int(bool(downsample)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_octree_free(self.h)
            self.h = None

    def add(self, pts):
        p = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
        self.L.ref_octree_update(self.h, p.ctypes.data, len(p))

    def size(self):
        return int(self.L.ref_octree_size(self.h))

    def points(self):
        out = np.zeros((max(self.size(), 1), 3), np.float32)
        n = self.L.ref_octree_dump(self.h, out.ctypes.data, len(out))
        return out[:n]

    def knn(self, queries, k=5):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, 3)
        xyz = np.zeros((len(q), k, 3), np.float32)
        d2 = np.zeros((len(q), k), np.float32)
        cnt = np.zeros(len(q), np.int32)
        self.L.ref_octree_knn(self.h, q.ctypes.data, len(q), int(k), xyz.ctypes.data, d2.ctypes.data, cnt.ctypes.data)
        return xyz, d2, cnt
