"""Host-side mirror of the reference's call surface for the registration hot path.

The reference is a C++ library (fast_limo::Mapper / fast_limo::Localizer, singletons; see
INTEGRATION.md for the C++ binding).  This module exposes the same operations, with the same
names and argument meaning, over the C ABI in ``include/flimo.h`` so that the parity tests and
``bench.py`` read like calls into the reference:

    Mapper.set_config / exists / size / last_time / add / match      (Modules/Mapper.hpp:49-60)
    Localizer-side iterated update: ``Registration.update``           (Localizer.cpp:326-353 ->
                                                                       esekfom.hpp:1620-1823)

No computation happens in Python and nothing here falls back to a CPU path.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import FlimoCfg, FlimoError, FlimoMsgLayout, FlimoPrepCfg, FlimoStats


@dataclass
class MappingConfig:
    """fast_limo::Config::iKFoM::Mapping (+ the two iKFoM flags the path reads)."""
    NUM_MATCH_POINTS: int = 5
    MAX_NUM_MATCHES: int = 2000
    MAX_NUM_PC2MATCH: int = 10000
    MAX_DIST_PLANE: float = 2.0
    PLANE_THRESHOLD: float = 5.0e-2
    estimate_extrinsics: bool = True
    octree_bucket_size: int = 2          # ignored, exactly like the reference
    octree_min_extent: float = 0.2
    octree_downsampling: bool = True
    knn_cell: float = 0.0                # extension: device grid cell (0 = auto)
    sort_scan: bool = False              # extension: Morton-sort the scan on upload (default: in-kernel scatter instead)
    knn_level_ratio: float = 0.0         # extension: cell growth between index levels (0 = 1.5)
    knn_tau: int = 0                     # extension: candidates-per-block threshold of the level choice (0 = 8)


# fast_limo::Point (Common.hpp:100-113) and the fast_limo::State members the deskew reads, as numpy records
RAW_POINT = np.dtype({"names": ["x", "y", "z", "intensity", "t", "time", "timestamp"],
                      "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<f4", "<f8"],
                      "offsets": [0, 4, 8, 16, 24, 24, 24], "itemsize": 32})
FRAME = np.dtype([("time", "<f8"), ("q", "<f4", 4), ("p", "<f4", 3), ("v", "<f4", 3), ("w", "<f4", 3), ("a", "<f4", 3),
                  ("bg", "<f4", 3), ("ba", "<f4", 3), ("g", "<f4", 3)], align=True)


@dataclass
class FilterConfig:
    """fast_limo::Config::filters + sensor_type / end_of_sweep (Config.hpp:43-55,83-91)."""
    cropBoxMin: tuple = None             # crop_active when both boxes are given
    cropBoxMax: tuple = None
    min_dist: float = None               # dist_active when not None
    rate_value: int = None               # rate_active when not None
    fov_angle: float = None              # fov_active when not None (radians)
    leafSize: float = None               # voxel_active when not None
    sensor_type: int = 1
    end_of_sweep: bool = False

    def to_c(self):
        c = FlimoPrepCfg()
        if self.cropBoxMin is not None and self.cropBoxMax is not None:
            c.crop_active = 1
            c.cropBoxMin[:] = [float(v) for v in self.cropBoxMin]
            c.cropBoxMax[:] = [float(v) for v in self.cropBoxMax]
        if self.min_dist is not None:
            c.dist_active, c.min_dist = 1, float(self.min_dist)
        c.rate_value = 1
        if self.rate_value is not None:
            c.rate_active, c.rate_value = 1, int(self.rate_value)
        if self.fov_angle is not None:
            c.fov_active, c.fov_angle = 1, float(self.fov_angle)
        if self.leafSize is not None:
            c.voxel_active, c.leafSize = 1, float(self.leafSize)
        c.sensor_type, c.end_of_sweep = int(self.sensor_type), int(bool(self.end_of_sweep))
        return c


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class PassResult:
    HTH: np.ndarray
    HTh: np.ndarray
    n_valid: int
    n_rows: int
    sum_sq_res: float


class Mapper:
    """fast_limo::Mapper over one GPU (the reference's is a process-wide singleton)."""

    def __init__(self, config: MappingConfig = None, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.device = device
        self.set_config(config or MappingConfig())

    # -- lifecycle ---------------------------------------------------------------------------
    def set_config(self, cfg: MappingConfig):
        c = FlimoCfg()
        self._L.flimo_cfg_default(C.byref(c))
        c.NUM_MATCH_POINTS = int(cfg.NUM_MATCH_POINTS)
        c.MAX_NUM_MATCHES = int(cfg.MAX_NUM_MATCHES)
        c.MAX_NUM_PC2MATCH = int(cfg.MAX_NUM_PC2MATCH)
        c.MAX_DIST_PLANE = float(cfg.MAX_DIST_PLANE)
        c.PLANE_THRESHOLD = float(cfg.PLANE_THRESHOLD)
        c.estimate_extrinsics = int(bool(cfg.estimate_extrinsics))
        c.octree_bucket_size = int(cfg.octree_bucket_size)
        c.octree_min_extent = float(cfg.octree_min_extent)
        c.octree_downsampling = int(bool(cfg.octree_downsampling))
        c.knn_cell = float(cfg.knn_cell)
        c.sort_scan = int(bool(cfg.sort_scan))
        c.knn_level_ratio = float(cfg.knn_level_ratio)
        c.knn_tau = int(cfg.knn_tau)
        if self._h:
            self._L.flimo_destroy(self._h)
            self._h = C.c_void_p()
        rc = self._L.flimo_create(C.byref(c), int(self.device), C.byref(self._h))
        if rc != 0:
            raise FlimoError(f"flimo_create failed ({rc}): {self._L.flimo_last_error(None).decode()}")
        self.config = cfg

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self._L.flimo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise FlimoError(f"libflimo_cuda error {rc}: {self._L.flimo_last_error(self._h).decode()}")

    # -- Mapper API --------------------------------------------------------------------------
    def exists(self):
        return bool(self._L.flimo_map_exists(self._h))

    def size(self):
        n = C.c_size_t(0)
        self._ck(self._L.flimo_map_size(self._h, C.byref(n)))
        return int(n.value)

    def last_time(self):
        return float(self._L.flimo_map_last_time(self._h))

    def add(self, points, time=0.0):
        """Mapper::add(pc, time): world-frame points, (n, >=3) float32 (row stride = itemsize*cols)."""
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, np.shape(points)[-1] if np.ndim(points) == 2 else 3)
        self._ck(self._L.flimo_map_add(self._h, pts.ctypes.data, pts.shape[0], 4 * pts.shape[1], float(time)))

    def add_device(self, dptr, n, stride_bytes, time=0.0):
        self._ck(self._L.flimo_map_add_device(self._h, C.c_void_p(dptr), n, stride_bytes, float(time)))

    def points(self):
        n = self.size()
        out = np.empty((max(n, 1), 3), np.float32)
        got = C.c_size_t(0)
        self._ck(self._L.flimo_map_get_points(self._h, _fp(out), n, C.byref(got)))
        return out[:n]

    def set_scan(self, scan):
        """Binds Localizer::pc2match (body-frame points of the current scan)."""
        sc = np.ascontiguousarray(scan, dtype=np.float32).reshape(-1, np.shape(scan)[-1] if np.ndim(scan) == 2 else 3)
        self._ck(self._L.flimo_scan_set(self._h, sc.ctypes.data, sc.shape[0], 4 * sc.shape[1]))
        self._scan_n = min(sc.shape[0], self.config.MAX_NUM_PC2MATCH)
        self._scan_n_full = sc.shape[0]

    def set_scan_device(self, dptr, n, stride_bytes):
        self._ck(self._L.flimo_scan_set_device(self._h, C.c_void_p(dptr), n, stride_bytes))
        self._scan_n = min(n, self.config.MAX_NUM_PC2MATCH)
        self._scan_n_full = n

    def set_scan_host(self, hptr, n, stride_bytes):
        """flimo_scan_set on a raw host pointer (e.g. pinned memory); binds a prefetched copy if there is one."""
        self._ck(self._L.flimo_scan_set(self._h, C.c_void_p(hptr), n, stride_bytes))
        self._scan_n = min(n, self.config.MAX_NUM_PC2MATCH)
        self._scan_n_full = n

    def prefetch_scan_host(self, hptr, n, stride_bytes):
        """flimo_scan_prefetch: start the H2D copy of the NEXT scan on the copy stream."""
        self._ck(self._L.flimo_scan_prefetch(self._h, C.c_void_p(hptr), n, stride_bytes))

    # -- scan preparation (Localizer::updatePointCloud before the hot path) -----------------------
    def prep_filter_sort(self, raw, sweep_ref_time, filters: "FilterConfig"):
        """Localizer.cpp:262-302 + the time sort of deskewPointCloud.  raw: RAW_POINT records.
        Returns (n_kept, t_last)."""
        raw = np.ascontiguousarray(raw, RAW_POINT)
        c = filters.to_c()
        n, t = C.c_size_t(0), C.c_double(0)
        self._ck(self._L.flimo_prep_filter_sort(self._h, raw.ctypes.data, len(raw), float(sweep_ref_time), C.byref(c),
                                                C.byref(n), C.byref(t)))
        return int(n.value), float(t.value)

    def prep_filter_sort_msg(self, data, point_step, sweep_ref_time, filters: "FilterConfig", off_xyz=(0, 4, 8), off_intensity=-1,
                             off_time=-1, time_datatype=7):
        """Same on a sensor_msgs/PointCloud2 payload (bytes / uint8 array): decode on the device, then filters + sort."""
        buf = np.frombuffer(data, np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        n_pts = len(buf) // int(point_step)
        lay = FlimoMsgLayout(int(off_xyz[0]), int(off_xyz[1]), int(off_xyz[2]), int(off_intensity), int(off_time), int(time_datatype))
        c = filters.to_c()
        n, t = C.c_size_t(0), C.c_double(0)
        self._ck(self._L.flimo_prep_filter_sort_msg(self._h, buf.ctypes.data, n_pts, int(point_step), C.byref(lay), float(sweep_ref_time),
                                                    C.byref(c), C.byref(n), C.byref(t)))
        return int(n.value), float(t.value)

    def prep_deskew(self, frames, last_q, last_p, T_lidar2baselink, offset=0.0):
        """deskewPointCloud's per-point loop (+ voxel grid if configured); binds pc2match as the scan."""
        fr = np.ascontiguousarray(frames, FRAME)
        lq, lp = np.ascontiguousarray(last_q, np.float32), np.ascontiguousarray(last_p, np.float32)
        T = np.ascontiguousarray(T_lidar2baselink, np.float32).reshape(16)
        n = C.c_size_t(0)
        self._ck(self._L.flimo_prep_deskew(self._h, fr.ctypes.data, len(fr), _fp(lq), _fp(lp), _fp(T), float(offset), C.byref(n)))
        self._scan_n = min(int(n.value), self.config.MAX_NUM_PC2MATCH)
        self._scan_n_full = int(n.value)
        return int(n.value)

    def prep_get(self, what):
        """0: sorted indices into the raw message, 1: deskewed world cloud, 2: deskewed body cloud, 3: pc2match."""
        n = C.c_size_t(0)
        self._ck(self._L.flimo_prep_get(self._h, int(what), None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.uint32) if what == 0 else np.zeros((max(n.value, 1), 4), np.float32)
        self._ck(self._L.flimo_prep_get(self._h, int(what), out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]

    def voxel_grid(self, pts4, leaf):
        """pcl::VoxelGrid centroids (leaf, leaf, leaf) of xyz1 points; ascending voxel index."""
        p = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
        out = np.zeros((max(len(p), 1), 4), np.float32)
        n = C.c_size_t(0)
        self._ck(self._L.flimo_voxel_grid(self._h, _fp(p), len(p), float(leaf), _fp(out), len(out), C.byref(n)))
        return out[:n.value]

    def shard(self, begin, end):
        self._ck(self._L.flimo_scan_shard(self._h, begin, end))

    def match(self, state, scan=None) -> PassResult:
        """Mapper::match(State, pc) + Localizer::calculate_H + H^T H / H^T h, reduced form."""
        if scan is not None:
            self.set_scan(scan)
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        HTH = np.zeros((12, 12), np.float64)
        HTh = np.zeros(12, np.float64)
        nv, nr, ss = C.c_int64(0), C.c_int64(0), C.c_double(0)
        self._ck(self._L.flimo_match_reduce(self._h, _dp(st), _dp(HTH), _dp(HTh), C.byref(nv), C.byref(nr), C.byref(ss)))
        return PassResult(HTH, HTh, int(nv.value), int(nr.value), float(ss.value))

    def exchange_attach(self, buf, rank, world):
        """Attach a host segment shared by all ranks (flimo_exchange_attach); `buf` is a writable buffer."""
        self._xch_keep = buf
        addr = C.addressof(C.c_char.from_buffer(buf))
        self._ck(self._L.flimo_exchange_attach(self._h, C.c_void_p(addr), len(buf), int(rank), int(world)))

    def match_exchange(self, state) -> PassResult:
        """One pass on this rank's shard, summed over all ranks through the shared segment."""
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        HTH = np.zeros((12, 12), np.float64)
        HTh = np.zeros(12, np.float64)
        nv, nr, ss = C.c_int64(0), C.c_int64(0), C.c_double(0)
        self._ck(self._L.flimo_match_reduce_exchange(self._h, _dp(st), _dp(HTH), _dp(HTh), C.byref(nv), C.byref(nr), C.byref(ss)))
        return PassResult(HTH, HTh, int(nv.value), int(nr.value), float(ss.value))

    def match_async(self, state, d_out_ptr, stream=None):
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        self._ck(self._L.flimo_match_reduce_async(self._h, _dp(st), C.c_void_p(d_out_ptr), C.c_void_p(stream or 0)))

    def match_debug(self, state):
        """Per-point records of one pass: dict of arrays in original scan order."""
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        n = self._scan_n
        out = np.zeros((max(n, 1), 16), np.float32)
        got = C.c_size_t(0)
        self._ck(self._L.flimo_match_debug(self._h, _dp(st), _fp(out), n, C.byref(got)))
        out = out[:n]
        return dict(world=out[:, 0:3], plane=out[:, 3:7], dist=out[:, 7], good=out[:, 8] > 0.5, nn_d2=out[:, 9:14],
                    levels=out[:, 14].astype(np.int32), first_count=out[:, 15].astype(np.int32))

    def scan_to_world(self, state):
        """pcl::transformPointCloud(pc2match -> world) (Localizer.cpp:361): the WHOLE bound cloud, original order."""
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        n = self._scan_n_full
        out = np.zeros((max(n, 1), 3), np.float32)
        got = C.c_size_t(0)
        self._ck(self._L.flimo_scan_to_world(self._h, _dp(st), _fp(out), n, C.byref(got)))
        return out[:n]

    def add_scan(self, state, time=0.0):
        """Localizer.cpp:361 + :377 on the device: transform the bound cloud with `state` and Mapper::add it."""
        st = np.ascontiguousarray(np.asarray(state, np.float64)[:14])
        self._ck(self._L.flimo_map_add_scan(self._h, _dp(st), float(time)))

    def stats(self):
        s = FlimoStats()
        self._ck(self._L.flimo_get_stats(self._h, C.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    def stream(self):
        return int(self._L.flimo_stream(self._h) or 0)

    # -- iterated update ---------------------------------------------------------------------
    def _update_buffers(self):
        """Reusable argument buffers of update(): building ctypes pointers costs more than a whole pass."""
        b = getattr(self, "_ubuf", None)
        if b is None:
            x, Pm, lim = np.zeros(26, np.float64), np.zeros((23, 23), np.float64), np.zeros(23, np.float64)
            b = self._ubuf = (x, Pm, lim, _dp(x), _dp(Pm), _dp(lim), C.c_int(0))
        return b

    def _update(self, fn, state26, P, max_iter, limits, R, D):
        x, Pm, lim, px, pP, pl, passes = self._update_buffers()
        x[:] = state26
        Pm[:] = np.asarray(P, np.float64).reshape(23, 23)
        lim[:] = limits
        self._ck(fn(self._h, px, pP, int(max_iter), pl, float(R), float(D), C.byref(passes)))
        return x.copy(), Pm.copy(), int(passes.value)

    def update(self, state26, P, max_iter, limits, R=0.001, D=5.0):
        """esekf::update_iterated_dyn_share_modified on the bound scan.  Returns (x, P, passes)."""
        return self._update(self._L.flimo_update, state26, P, max_iter, limits, R, D)

    def update_exchange(self, state26, P, max_iter, limits, R=0.001, D=5.0):
        """`update` with every pass summed over all ranks through the attached exchange segment."""
        return self._update(self._L.flimo_update_exchange, state26, P, max_iter, limits, R, D)

    def update_peer(self, state26, P, max_iter, limits, R=0.001, D=5.0):
        """`update` of a scan sharded over the ranks attached with peer_attach: the whole update runs on the device,
        pass sums travel over NVLink peer memory (flimo_update_peer)."""
        return self._update(self._L.flimo_update_peer, state26, P, max_iter, limits, R, D)

    def peer_export(self):
        """This rank's inbox as a CUDA IPC handle (64 bytes) for the other ranks."""
        buf = (C.c_ubyte * 64)()
        self._ck(self._L.flimo_peer_export(self._h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peer_attach(self, rank, world, handles):
        """handles: the 64-byte IPC handles of all ranks, in rank order."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._ck(self._L.flimo_peer_attach(self._h, int(rank), int(world), C.cast(buf, C.c_void_p)))

    def update_trace(self):
        """Per-pass records of the last device-resident update: (passes, 32) = state26, n_valid, n_rows, ns, row limit."""
        out = np.zeros((16, 32), np.float64)
        n = C.c_size_t(0)
        self._ck(self._L.flimo_update_trace(self._h, _dp(out), 16, C.byref(n)))
        return out[:n.value]

    def ekf_begin(self, state26, P, max_iter, limits, R=0.001, D=5.0):
        x = np.ascontiguousarray(state26, np.float64)
        Pm = np.ascontiguousarray(np.asarray(P, np.float64).reshape(23, 23))
        lim = np.ascontiguousarray(np.broadcast_to(np.asarray(limits, np.float64), (23,)))
        self._ck(self._L.flimo_ekf_begin(self._h, _dp(x), _dp(Pm), int(max_iter), _dp(lim), float(R), float(D)))

    def ekf_state(self):
        x = np.zeros(26, np.float64)
        self._ck(self._L.flimo_ekf_state(self._h, _dp(x)))
        return x

    def ekf_step(self, HTH, HTh, n_rows):
        HTH = np.ascontiguousarray(HTH, np.float64)
        HTh = np.ascontiguousarray(HTh, np.float64)
        done = C.c_int(0)
        self._ck(self._L.flimo_ekf_step(self._h, _dp(HTH), _dp(HTh), int(n_rows), C.byref(done)))
        return bool(done.value)

    def ekf_end(self):
        x = np.zeros(26, np.float64)
        Pm = np.zeros((23, 23), np.float64)
        self._ck(self._L.flimo_ekf_end(self._h, _dp(x), _dp(Pm)))
        return x, Pm


    # ---- IMU rate (host algebra; works on a handle created with device=-1) ----
    def ekf_predict(self, state26, P, stamp, dt, lin_accel, ang_vel, cov=(6.e-4, 1.e-2, 1.e-5, 3.e-4)):
        """Localizer::propagateImu(imu): esekf::predict + push of State(x, stamp, a, w) on the propagated ring.
        cov = Config::iKFoM (cov_gyro, cov_acc, cov_bias_gyro, cov_bias_acc), defaults of src/main.cpp:159-162."""
        x = np.array(state26, np.float64).reshape(26)
        Pm = np.array(P, np.float64).reshape(23, 23)
        imu = _lib.FlimoImu(float(stamp), float(dt))
        imu.ang_vel[:] = [float(np.float32(v)) for v in ang_vel]
        imu.lin_accel[:] = [float(np.float32(v)) for v in lin_accel]
        c4 = np.ascontiguousarray(cov, np.float64).reshape(4)
        self._ck(self._L.flimo_ekf_predict(self._h, _dp(x), _dp(Pm), C.byref(imu), _dp(c4)))
        return x, Pm

    def propagated_frames(self, start_time, end_time):
        """Localizer::integrateImu(start_time, end_time): the frames prep_deskew takes (FRAME records, oldest first)."""
        n = C.c_size_t(0)
        self._ck(self._L.flimo_propagated_frames(self._h, float(start_time), float(end_time), None, 0, C.byref(n)))
        out = np.zeros(n.value, FRAME)
        if n.value:
            self._ck(self._L.flimo_propagated_frames(self._h, float(start_time), float(end_time), out.ctypes.data, n.value, C.byref(n)))
        return out

    def propagated_clear(self):
        self._ck(self._L.flimo_propagated_clear(self._h))


def unpack96(packed):
    L = _lib.load()
    p = np.ascontiguousarray(packed, np.float64)
    HTH = np.zeros((12, 12), np.float64)
    HTh = np.zeros(12, np.float64)
    nv, nr, ss = C.c_int64(0), C.c_int64(0), C.c_double(0)
    L.flimo_unpack96(_dp(p), _dp(HTH), _dp(HTh), C.byref(nv), C.byref(nr), C.byref(ss))
    return PassResult(HTH, HTh, int(nv.value), int(nr.value), float(ss.value))
