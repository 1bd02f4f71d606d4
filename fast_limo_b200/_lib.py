"""ctypes binding of libflimo_cuda.so (the C ABI declared in include/flimo.h).

The library is the product; there is no Python or CPU fallback.  If the shared object is
missing (not built) or no sm_100 GPU is present, importing is fine but the first call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("FLIMO_LIB_NAME", "libflimo_cuda.so"))   # FLIMO_LIB_NAME: A/B builds of tools/

# every symbol include/flimo.h declares (tests check that the .so exports all of them)
SYMBOLS = [
    "flimo_cfg_default", "flimo_create", "flimo_destroy", "flimo_last_error", "flimo_version",
    "flimo_map_add", "flimo_map_add_device", "flimo_map_size", "flimo_map_exists", "flimo_map_last_time",
    "flimo_map_get_points", "flimo_scan_set", "flimo_scan_set_device", "flimo_scan_prefetch", "flimo_scan_shard",
    "flimo_match_reduce", "flimo_match_reduce_async", "flimo_unpack96", "flimo_match_debug",
    "flimo_update", "flimo_ekf_begin", "flimo_ekf_state", "flimo_ekf_step", "flimo_ekf_end",
    "flimo_scan_to_world", "flimo_map_add_scan", "flimo_prep_filter_sort", "flimo_prep_filter_sort_msg", "flimo_prep_deskew", "flimo_prep_get", "flimo_voxel_grid", "flimo_get_stats", "flimo_stream", "flimo_exchange_attach", "flimo_match_reduce_exchange", "flimo_update_exchange",
    "flimo_ekf_predict", "flimo_propagated_frames", "flimo_propagated_clear",
    "flimo_peer_export", "flimo_peer_attach", "flimo_update_peer", "flimo_update_trace",
]


class FlimoCfg(C.Structure):
    _fields_ = [
        ("NUM_MATCH_POINTS", C.c_int32),
        ("MAX_NUM_MATCHES", C.c_int32),
        ("MAX_NUM_PC2MATCH", C.c_int32),
        ("estimate_extrinsics", C.c_int32),
        ("MAX_DIST_PLANE", C.c_double),
        ("PLANE_THRESHOLD", C.c_double),
        ("octree_bucket_size", C.c_int32),
        ("octree_downsampling", C.c_int32),
        ("octree_min_extent", C.c_float),
        ("knn_cell", C.c_float),
        ("sort_scan", C.c_int32),
        ("knn_level_ratio", C.c_float),
        ("knn_tau", C.c_int32),
    ]


class FlimoPrepCfg(C.Structure):
    """flimo_prep_cfg: Config::filters + the sensor flags of deskewPointCloud."""
    _fields_ = [
        ("crop_active", C.c_int32), ("dist_active", C.c_int32), ("rate_active", C.c_int32), ("fov_active", C.c_int32),
        ("voxel_active", C.c_int32),
        ("cropBoxMin", C.c_float * 3), ("cropBoxMax", C.c_float * 3),
        ("min_dist", C.c_double),
        ("rate_value", C.c_int32),
        ("fov_angle", C.c_float),
        ("leafSize", C.c_float),
        ("sensor_type", C.c_int32),
        ("end_of_sweep", C.c_int32),
    ]


class FlimoMsgLayout(C.Structure):
    """flimo_msg_layout: byte offsets of the fast_limo::Point fields inside a PointCloud2 point."""
    _fields_ = [("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_intensity", C.c_int32),
                ("off_time", C.c_int32), ("time_datatype", C.c_int32)]


class FlimoStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("match_launches", C.c_uint64),
        ("last_match_ms", C.c_float),
        ("match_ms_total", C.c_double),
        ("match_timed", C.c_uint64),
        ("knn_cell", C.c_float),
        ("grid_nx", C.c_int32),
        ("grid_ny", C.c_int32),
        ("grid_nz", C.c_int32),
        ("n_levels", C.c_int32),
        ("table_bytes", C.c_uint64),
        ("map_bytes", C.c_uint64),
        ("persist_ms_total", C.c_double),
        ("persist_passes", C.c_uint64),
        ("index_builds", C.c_uint64),
        ("index_updates", C.c_uint64),
        ("index_rows_moved", C.c_uint64),
        ("exchange_ms_total", C.c_double),
        ("update_stalls", C.c_uint64),
    ]


class FlimoImu(C.Structure):
    """flimo_imu: fast_limo::IMUmeas members propagateImu reads (Common.hpp:126-132)."""
    _fields_ = [("stamp", C.c_double), ("dt", C.c_double), ("ang_vel", C.c_float * 3), ("lin_accel", C.c_float * 3)]


_lib = None


class FlimoError(RuntimeError):
    pass


def load():
    """Load libflimo_cuda.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlimoError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, sz, dbl, i64 = C.c_void_p, C.c_size_t, C.c_double, C.c_int64
    pd, pf = C.POINTER(C.c_double), C.POINTER(C.c_float)
    L.flimo_cfg_default.argtypes = [C.POINTER(FlimoCfg)]
    L.flimo_cfg_default.restype = None
    L.flimo_create.argtypes = [C.POINTER(FlimoCfg), C.c_int, C.POINTER(vp)]
    L.flimo_destroy.argtypes = [vp]
    L.flimo_destroy.restype = None
    L.flimo_last_error.argtypes = [vp]
    L.flimo_last_error.restype = C.c_char_p
    L.flimo_version.restype = C.c_char_p
    L.flimo_map_add.argtypes = [vp, vp, sz, sz, dbl]
    L.flimo_map_add_device.argtypes = [vp, vp, sz, sz, dbl]
    L.flimo_map_size.argtypes = [vp, C.POINTER(sz)]
    L.flimo_map_exists.argtypes = [vp]
    L.flimo_map_last_time.argtypes = [vp]
    L.flimo_map_last_time.restype = dbl
    L.flimo_map_get_points.argtypes = [vp, pf, sz, C.POINTER(sz)]
    L.flimo_scan_set.argtypes = [vp, vp, sz, sz]
    L.flimo_scan_set_device.argtypes = [vp, vp, sz, sz]
    L.flimo_scan_prefetch.argtypes = [vp, vp, sz, sz]
    L.flimo_scan_shard.argtypes = [vp, sz, sz]
    L.flimo_match_reduce.argtypes = [vp, pd, pd, pd, C.POINTER(i64), C.POINTER(i64), pd]
    L.flimo_match_reduce_async.argtypes = [vp, pd, vp, vp]
    L.flimo_exchange_attach.argtypes = [vp, vp, sz, C.c_int, C.c_int]
    L.flimo_match_reduce_exchange.argtypes = [vp, pd, pd, pd, C.POINTER(i64), C.POINTER(i64), pd]
    L.flimo_unpack96.argtypes = [pd, pd, pd, C.POINTER(i64), C.POINTER(i64), pd]
    L.flimo_unpack96.restype = None
    L.flimo_match_debug.argtypes = [vp, pd, pf, sz, C.POINTER(sz)]
    L.flimo_update.argtypes = [vp, pd, pd, C.c_int, pd, dbl, dbl, C.POINTER(C.c_int)]
    L.flimo_update_exchange.argtypes = [vp, pd, pd, C.c_int, pd, dbl, dbl, C.POINTER(C.c_int)]
    L.flimo_update_peer.argtypes = [vp, pd, pd, C.c_int, pd, dbl, dbl, C.POINTER(C.c_int)]
    L.flimo_peer_export.argtypes = [vp, vp]
    L.flimo_peer_attach.argtypes = [vp, C.c_int, C.c_int, vp]
    L.flimo_update_trace.argtypes = [vp, pd, sz, C.POINTER(sz)]
    L.flimo_ekf_begin.argtypes = [vp, pd, pd, C.c_int, pd, dbl, dbl]
    L.flimo_ekf_state.argtypes = [vp, pd]
    L.flimo_ekf_step.argtypes = [vp, pd, pd, i64, C.POINTER(C.c_int)]
    L.flimo_ekf_end.argtypes = [vp, pd, pd]
    L.flimo_scan_to_world.argtypes = [vp, pd, pf, sz, C.POINTER(sz)]
    L.flimo_map_add_scan.argtypes = [vp, pd, dbl]
    L.flimo_get_stats.argtypes = [vp, C.POINTER(FlimoStats)]
    L.flimo_prep_filter_sort.argtypes = [vp, vp, sz, dbl, C.POINTER(FlimoPrepCfg), C.POINTER(sz), pd]
    L.flimo_prep_filter_sort_msg.argtypes = [vp, vp, sz, sz, C.POINTER(FlimoMsgLayout), dbl, C.POINTER(FlimoPrepCfg), C.POINTER(sz), pd]
    L.flimo_prep_deskew.argtypes = [vp, vp, C.c_int, pf, pf, pf, dbl, C.POINTER(sz)]
    L.flimo_prep_get.argtypes = [vp, C.c_int, vp, sz, C.POINTER(sz)]
    L.flimo_voxel_grid.argtypes = [vp, pf, sz, C.c_float, pf, sz, C.POINTER(sz)]
    L.flimo_ekf_predict.argtypes = [vp, pd, pd, C.POINTER(FlimoImu), pd]
    L.flimo_propagated_frames.argtypes = [vp, dbl, dbl, vp, sz, C.POINTER(sz)]
    L.flimo_propagated_clear.argtypes = [vp]
    L.flimo_stream.argtypes = [vp]
    L.flimo_stream.restype = vp
    _lib = L
    return L
