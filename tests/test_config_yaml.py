"""CPU: the reference's YAML parameter format -> this package's configuration objects (fast_limo_b200/config.py).
Keys and defaults: src/main.cpp:101-168; post-processing: fast_limo/Modules/Localizer.cpp:36-89."""
import glob
import math
import os

import numpy as np
import pytest

from fast_limo_b200 import api
from fast_limo_b200.config import load_config

YAML = """
num_threads: 6
sensor_type: 3
estimate_extrinsics: false
time_offset: false
end_of_sweep: true
extrinsics:
  lidar:
    t: [ 0.5, -0.25, 1.0 ]
    R: [ 0., -1., 0.,
         1.,  0., 0.,
         0.,  0., 1. ]
filters:
  cropBox:
    active: true
    box:
      min: [ -2.0, -1.5, -1.0 ]
      max: [ 2.0, 1.5, 1.0 ]
  voxelGrid:
    active: true
    leafSize: [ 0.75, 0.5, 0.5 ]
  minDistance:
    active: true
    value: 3.5
  FoV:
    active: true
    value: 180
  rateSampling:
    active: false
    value: 7
iKFoM:
  MAX_NUM_ITERS: 2
  MAX_NUM_MATCHES: 5000
  MAX_NUM_PC2MATCH: 1.e+4
  LIMITS: 0.002
  Mapping:
    NUM_MATCH_POINTS: 5
    MAX_DIST_PLANE: 1.5
    PLANES_THRESHOLD: 4.0e-2
    Octree:
      bucket_size: 2
      min_extent: 0.3
      downsampling: false
  covariance:
    gyro: 6.01e-4
    accel: 1.53e-2
    bias_gyro: 1.54e-5
    bias_accel: 3.38e-4
"""


def test_yaml_keys_and_postprocessing(tmp_path):
    p = tmp_path / "robot.yaml"
    p.write_text(YAML)
    c = load_config(str(p))
    m, l, f = c.mapping, c.localizer, c.localizer.filters
    assert (m.MAX_NUM_MATCHES, m.MAX_NUM_PC2MATCH, m.MAX_DIST_PLANE, m.PLANE_THRESHOLD) == (5000, 10000, 1.5, 0.04)
    assert (m.estimate_extrinsics, m.octree_min_extent, m.octree_downsampling, m.octree_bucket_size) == (False, 0.3, False, 2)
    assert (l.MAX_NUM_ITERS, l.LIMITS, l.time_offset) == (2, 0.002, False)
    assert (l.cov_gyro, l.cov_acc, l.cov_bias_gyro, l.cov_bias_acc) == (6.01e-4, 1.53e-2, 1.54e-5, 3.38e-4)
    assert f.cropBoxMin == (-2.0, -1.5, -1.0) and f.cropBoxMax == (2.0, 1.5, 1.0)
    assert f.min_dist == 3.5 and f.rate_value is None                  # inactive filters are off whatever their value
    assert abs(f.fov_angle - math.pi / 2) < 1e-7                       # half of the field of view, radians (main.cpp:145)
    assert f.leafSize == 0.75                                          # leafSize[0] for all three axes (Localizer.cpp:61)
    assert (f.sensor_type, f.end_of_sweep, c.num_threads) == (3, True, 6)
    # the rotation is used as written in the file (Map reads it column-major, init transposes it back)
    assert np.array_equal(np.asarray(l.lidar2baselink_R), [[0, -1, 0], [1, 0, 0], [0, 0, 1]])
    assert l.lidar2baselink_t == (0.5, -0.25, 1.0)
    pc = f.to_c()                                                      # and it reaches the C ABI struct
    assert (pc.crop_active, pc.dist_active, pc.rate_active, pc.fov_active, pc.voxel_active) == (1, 1, 0, 1, 1)
    assert abs(pc.leafSize - 0.75) < 1e-7 and pc.sensor_type == 3 and pc.end_of_sweep == 1


def test_defaults_are_the_wrappers():
    c = load_config({})
    m, l, f = c.mapping, c.localizer, c.localizer.filters
    assert m == api.MappingConfig()                                    # NUM_MATCH_POINTS 5, 2000 / 10000 caps, 2.0, 0.05, ...
    assert (l.MAX_NUM_ITERS, l.LIMITS, l.time_offset) == (3, 1e-3, True)
    assert (l.cov_gyro, l.cov_acc, l.cov_bias_gyro, l.cov_bias_acc) == (6.e-4, 1.e-2, 1.e-5, 3.e-4)
    assert f.cropBoxMin == (-1.0, -1.0, -1.0) and f.leafSize == 0.25 and f.min_dist is None and f.fov_angle is None
    assert c.calibration == dict(gravity_align=True, accel=True, gyro=True, time=3.0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/config"), reason="reference checkout not present (GPU box)")
def test_reference_yaml_files_load():
    """Every parameter file the reference ships parses; pass counts as SURVEY section 8 (a9) lists them."""
    files = sorted(glob.glob("/root/reference/config/*.yaml"))
    assert len(files) >= 4
    iters = {}
    for p in files:
        c = load_config(p)
        assert c.mapping.NUM_MATCH_POINTS == 5 and c.localizer.LIMITS > 0
        R = np.asarray(c.localizer.lidar2baselink_R, np.float64)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-3)              # a rotation, as written
        iters[os.path.basename(p)] = c.localizer.MAX_NUM_ITERS
    assert iters["kitti.yaml"] == 3 and iters["cat.yaml"] == 2
