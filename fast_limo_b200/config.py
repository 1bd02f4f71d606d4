"""The reference's YAML parameter files (config/*.yaml) -> the configuration objects of this package.

Same keys and the same defaults as `load_config` in the reference's ROS wrapper (src/main.cpp:101-168) and the same
post-processing as `Localizer::init` (fast_limo/Modules/Localizer.cpp:36-89): FoV degrees -> half angle in radians
(`main.cpp:145`), one LIMITS value for all 23 error-state entries (`:164-166`), leafSize[0] for all three axes
(`Localizer.cpp:61`), the extrinsic rotation as written in the file (Eigen::Map reads the row-major list column-major
and `init` transposes it back, `:80-82`).  Keys outside the registration path (topics, frames, debug, calibration
procedure) are parsed into `raw` and otherwise ignored.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from . import api
from .localizer import LocalizerConfig


def _get(d, path, default):
    for k in path.split("/"):
        if not isinstance(d, dict) or k not in d:
            return default
        d = d[k]
    return d


@dataclass
class Config:
    mapping: api.MappingConfig
    localizer: LocalizerConfig
    num_threads: int = 10
    sensor_type: int = 1
    calibration: dict = field(default_factory=dict)      # gravity_align / accel / gyro / time — IMU calibration stays with the caller
    intrinsics: dict = field(default_factory=dict)       # accel bias / sm, gyro bias
    imu2baselink: dict = field(default_factory=dict)     # R (3x3), t — IMU -> base-link transform stays with the caller
    raw: dict = field(default_factory=dict)


def load_config(src) -> Config:
    """src: path of a YAML file in the reference's format, or the already parsed dict."""
    if isinstance(src, dict):
        y = src
    else:
        import yaml
        with open(src) as f:
            y = yaml.safe_load(f) or {}

    mapping = api.MappingConfig(
        NUM_MATCH_POINTS=int(_get(y, "iKFoM/Mapping/NUM_MATCH_POINTS", 5)),
        MAX_NUM_MATCHES=int(float(_get(y, "iKFoM/MAX_NUM_MATCHES", 2000))),
        MAX_NUM_PC2MATCH=int(float(_get(y, "iKFoM/MAX_NUM_PC2MATCH", 1.e+4))),
        MAX_DIST_PLANE=float(_get(y, "iKFoM/Mapping/MAX_DIST_PLANE", 2.0)),
        PLANE_THRESHOLD=float(_get(y, "iKFoM/Mapping/PLANES_THRESHOLD", 5.e-2)),
        estimate_extrinsics=bool(_get(y, "estimate_extrinsics", True)),
        octree_bucket_size=int(_get(y, "iKFoM/Mapping/Octree/bucket_size", 2)),
        octree_min_extent=float(_get(y, "iKFoM/Mapping/Octree/min_extent", 0.2)),
        octree_downsampling=bool(_get(y, "iKFoM/Mapping/Octree/downsampling", True)),
    )

    sensor_type = int(_get(y, "sensor_type", 1))
    crop = bool(_get(y, "filters/cropBox/active", True))
    voxel = bool(_get(y, "filters/voxelGrid/active", True))
    filters = api.FilterConfig(
        cropBoxMin=tuple(float(v) for v in _get(y, "filters/cropBox/box/min", [-1.0, -1.0, -1.0])) if crop else None,
        cropBoxMax=tuple(float(v) for v in _get(y, "filters/cropBox/box/max", [1.0, 1.0, 1.0])) if crop else None,
        min_dist=float(_get(y, "filters/minDistance/value", 4.0)) if _get(y, "filters/minDistance/active", False) else None,
        rate_value=int(_get(y, "filters/rateSampling/value", 4)) if _get(y, "filters/rateSampling/active", False) else None,
        fov_angle=float(np.float32(_get(y, "filters/FoV/value", 360.0)) * math.pi / 360.0) if _get(y, "filters/FoV/active", False) else None,
        leafSize=float(_get(y, "filters/voxelGrid/leafSize", [0.25, 0.25, 0.25])[0]) if voxel else None,
        sensor_type=sensor_type,
        end_of_sweep=bool(_get(y, "end_of_sweep", False)),
    )

    R_l = np.asarray(_get(y, "extrinsics/lidar/R", [0.0] * 9), np.float32).reshape(3, 3)
    localizer = LocalizerConfig(
        filters=filters,
        MAX_NUM_ITERS=int(_get(y, "iKFoM/MAX_NUM_ITERS", 3)),
        LIMITS=float(_get(y, "iKFoM/LIMITS", 1.e-3)),
        cov_gyro=float(_get(y, "iKFoM/covariance/gyro", 6.e-4)),
        cov_acc=float(_get(y, "iKFoM/covariance/accel", 1.e-2)),
        cov_bias_gyro=float(_get(y, "iKFoM/covariance/bias_gyro", 1.e-5)),
        cov_bias_acc=float(_get(y, "iKFoM/covariance/bias_accel", 3.e-4)),
        time_offset=bool(_get(y, "time_offset", True)),
        lidar2baselink_R=tuple(tuple(float(v) for v in row) for row in R_l),
        lidar2baselink_t=tuple(float(v) for v in _get(y, "extrinsics/lidar/t", [0.0, 0.0, 0.0])),
    )
    return Config(
        mapping=mapping, localizer=localizer, num_threads=int(_get(y, "num_threads", 10)), sensor_type=sensor_type,
        calibration=dict(gravity_align=bool(_get(y, "calibration/gravity_align", True)), accel=bool(_get(y, "calibration/accel", True)),
                         gyro=bool(_get(y, "calibration/gyro", True)), time=float(_get(y, "calibration/time", 3.0))),
        intrinsics=dict(accel_bias=[float(v) for v in _get(y, "intrinsics/accel/bias", [0.0] * 3)],
                        gyro_bias=[float(v) for v in _get(y, "intrinsics/gyro/bias", [0.0] * 3)],
                        accel_sm=[float(v) for v in _get(y, "intrinsics/accel/sm", [0.0] * 9)]),
        imu2baselink=dict(R=np.asarray(_get(y, "extrinsics/imu/R", [0.0] * 9), np.float32).reshape(3, 3),
                          t=[float(v) for v in _get(y, "extrinsics/imu/t", [0.0] * 3)]),
        raw=y)
