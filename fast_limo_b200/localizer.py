"""Host mirror of fast_limo::Localizer's two callbacks over libflimo_cuda (ctypes), with the reference's names.

    Localizer::updateIMU         fast_limo/Modules/Localizer.cpp:401-531  (post-calibration branch :512-528)
    Localizer::propagateImu      :583-608
    Localizer::updatePointCloud  :245-399
    Localizer::deskewPointCloud  :733-853 (host part: time offset :797-805, integrateImu :808, last_state :820)
    Localizer::init_iKFoM_state  :672-694

Only sequencing lives here; every stage is a call into the library (`api.Mapper`): filters / sort / deskew / voxel
grid and the registration on the GPU, prediction and the propagated-state ring in the library's host algebra.
What is NOT mirrored: the standstill IMU calibration (:409-510) — `updateIMU` takes the sample as `imu_buffer` stores
it, `updateIMU_raw` applies the IMU -> base-link transform and the intrinsic correction first (:697-728, :512-520) —
the debug clouds, the CPU statistics and the condition-variable wait of
`propagatedFromTimeRange` (a scan that arrives before its IMU data raises instead).
"""
from dataclasses import dataclass, field

import numpy as np

from . import api


@dataclass
class LocalizerConfig:
    """The Config fields the two callbacks read (fast_limo/Utils/Config.hpp)."""
    filters: api.FilterConfig = field(default_factory=api.FilterConfig)
    MAX_NUM_ITERS: int = 3                       # iKFoM/MAX_NUM_ITERS
    LIMITS: float = 0.001                        # iKFoM/LIMITS (one value for all 23, as the YAMLs have it)
    cov_gyro: float = 6.e-4                      # defaults of src/main.cpp:159-162
    cov_acc: float = 1.e-2
    cov_bias_gyro: float = 1.e-5
    cov_bias_acc: float = 3.e-4
    time_offset: bool = False
    gravity: float = 9.81
    lidar2baselink_R: tuple = ((1, 0, 0), (0, 1, 0), (0, 0, 1))
    lidar2baselink_t: tuple = (0.0, 0.0, 0.0)
    calibrate_gyro: bool = False                 # True: the stand-still gyro bias stays constant (Localizer.cpp:344-345)
    calibrate_accel: bool = False


def _quat_from_R(R):
    from scipy.spatial.transform import Rotation
    return Rotation.from_matrix(np.asarray(R, np.float64)).as_quat()


def _RT_f32(q, t):
    """Eigen::Quaternionf::toRotationMatrix + translation as a row-major 4x4 float32 (State::get_extr_RT)."""
    x, y, z, w = (np.float32(v) for v in q)
    two = np.float32(2)
    tx, ty, tz = two * x, two * y, two * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    one = np.float32(1)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = [[one - (tyy + tzz), txy - twz, txz + twy], [txy + twz, one - (txx + tzz), tyz - twx],
                 [txz - twy, tyz + twx, one - (txx + tyy)]]
    T[:3, 3] = np.asarray(t, np.float32)
    return T


class Localizer:
    def __init__(self, mapper: api.Mapper, config: LocalizerConfig = None, pos=(0, 0, 0), quat=(0, 0, 0, 1), vel=(0, 0, 0),
                 bias_gyro=(0, 0, 0), bias_accel=(0, 0, 0)):
        self.map = mapper
        self.config = config or LocalizerConfig()
        c = self.config
        # init_iKFoM_state (:672-694)
        from .synth import default_P0, make_state
        self.x = make_state(pos, quat, _quat_from_R(c.lidar2baselink_R), c.lidar2baselink_t, vel, bias_gyro, bias_accel,
                            (0.0, 0.0, -c.gravity))
        # S2(Eigen::Vector3d) rescales to length 9.809 whatever `gravity` says (S2.hpp:123-126)
        self.x[23:26] *= 9.809 / np.linalg.norm(self.x[23:26])
        self.P = default_P0()
        self.lidar2baselink_T = _RT_f32(self.x[7:11], self.x[11:14])
        self.scan_stamp = 0.0
        self.prev_scan_stamp = 0.0
        self.imu_stamp = 0.0
        self.n_imu = 0
        self.last_imu = None
        self.last = {}                           # what the last updatePointCloud did (sizes, passes, "null" reason)
        # this->state.b: the biases subtracted from raw IMU samples (:514-515).  They follow the filter's estimate after
        # every LiDAR update unless the calibrate_* flags pin them to the calibrated values (:344-351).
        self.state_b_gyro = np.asarray(bias_gyro, np.float32).copy()
        self.state_b_accel = np.asarray(bias_accel, np.float32).copy()
        self.map.propagated_clear()

    def _cov4(self):
        c = self.config
        return (c.cov_gyro, c.cov_acc, c.cov_bias_gyro, c.cov_bias_acc)

    # -- IMU callback ---------------------------------------------------------------------------------------------
    def updateIMU(self, stamp, dt, lin_accel, ang_vel):
        """:512-528 for a calibrated base-link sample: remember it, propagate the filter, push the propagated state."""
        self.imu_stamp = float(stamp)
        self.last_imu = (np.asarray(lin_accel, np.float32), np.asarray(ang_vel, np.float32))
        self.n_imu += 1
        self.x, self.P = self.map.ekf_predict(self.x, self.P, stamp, dt, lin_accel, ang_vel, self._cov4())

    def updateIMU_raw(self, stamp, lin_accel, ang_vel, imu2baselink_R=np.eye(3), imu2baselink_t=(0, 0, 0), accel_sm=np.eye(3)):
        """:403-404 + :512-520 for a sample in the IMU frame: imu2baselink (:697-728; float32 like the reference: dt from
        the stamps with the 1/200 s fallback, rotation into the base-link frame, lever-arm terms) and the intrinsic
        correction, then `updateIMU`.  The standstill calibration (:409-510) stays with the caller."""
        f32 = np.float32
        R, t = np.asarray(imu2baselink_R, f32), np.asarray(imu2baselink_t, f32)
        dt = float(stamp) - getattr(self, "_prev_imu_stamp", 0.0)
        if dt == 0.0 or dt > 0.1:
            dt = 1.0 / 200.0
        w = R @ np.asarray(ang_vel, f32)
        w_prev = getattr(self, "_ang_vel_prev", w)            # function-local static, initialised with the first sample
        a = R @ np.asarray(lin_accel, f32)
        a = a + np.cross(((w - w_prev) / f32(dt)).astype(f32), -t) + np.cross(w, np.cross(w, -t))
        self._ang_vel_prev, self._prev_imu_stamp = w, float(stamp)
        a = (np.asarray(accel_sm, f32) @ a.astype(f32)) - self.state_b_accel
        w = w - self.state_b_gyro
        self.updateIMU(stamp, dt, a.astype(f32), w.astype(f32))

    # -- LiDAR callback -------------------------------------------------------------------------------------------
    def updatePointCloud(self, raw, time_stamp):
        """:245-399.  raw: api.RAW_POINT records of one message.  Returns True when the scan was registered and mapped."""
        self.last = {}
        if len(raw) < 1:
            return self._null("raw pointcloud is empty", stamp=False)
        if self.n_imu == 0:
            return self._null("IMU buffer is empty", stamp=False)
        c = self.config
        n_kept, t_last = self.map.prep_filter_sort(raw, time_stamp, c.filters)             # :262-302, :744-789
        self.last["n_filtered"] = n_kept
        if n_kept < 1:
            return self._null("no points left after the filters", stamp=False)
        offset = 0.0
        if c.time_offset:                                                                   # :797-801
            offset = min(self.imu_stamp - t_last - 1.e-4, 0.0)
        self.scan_stamp = t_last + offset                                                   # :805
        frames = self.map.propagated_frames(self.prev_scan_stamp, self.scan_stamp)          # :808
        self.last["n_frames"] = len(frames)
        if len(frames) < 1:
            return self._null("no frames obtained from IMU propagation")                    # the first scan (prev stamp 0)
        lq, lp = self.x[3:7].astype(np.float32), self.x[0:3].astype(np.float32)             # last_state (:820)
        n_pc2match = self.map.prep_deskew(frames, lq, lp, self.lidar2baselink_T, offset)    # :822-843 (+ :313-321)
        self.last["n_pc2match"] = n_pc2match
        if n_pc2match <= 1:
            return self._null("NULL ITERATION")
        # with an empty map Mapper::match returns no matches (Mapper.cpp:61): every pass has zero rows, the state
        # stays at the prediction and the map is then initialised with this scan
        self.x, self.P, passes = self.map.update(self.x, self.P, c.MAX_NUM_ITERS, c.LIMITS)       # :333
        self.last["passes"] = passes
        if not c.calibrate_gyro:                                                            # state = corrected_state (:344-351)
            self.state_b_gyro = self.x[17:20].astype(np.float32)
        if not c.calibrate_accel:
            self.state_b_accel = self.x[20:23].astype(np.float32)
        self.lidar2baselink_T = _RT_f32(self.x[7:11], self.x[11:14])                        # :356
        self.map.add_scan(self.x, self.scan_stamp)                                          # :361 + :377, all of pc2match
        self.prev_scan_stamp = self.scan_stamp                                              # :398
        return True

    def _null(self, why, stamp=True):
        self.last["null"] = why
        if stamp:
            self.prev_scan_stamp = self.scan_stamp                                          # :398 runs on every path past :805
        return False

    # -- getters ---------------------------------------------------------------------------------------------------
    def _state_f32(self):
        x = self.x
        return dict(p=x[0:3].astype(np.float32), q=x[3:7].astype(np.float32), qLI=x[7:11].astype(np.float32),
                    pLI=x[11:14].astype(np.float32), v=x[14:17].astype(np.float32), bg=x[17:20].astype(np.float32),
                    ba=x[20:23].astype(np.float32), g=x[23:26].astype(np.float32),
                    w=None if self.last_imu is None else self.last_imu[1], a=None if self.last_imu is None else self.last_imu[0],
                    time=self.imu_stamp)

    def getWorldState(self):
        """:174-188: State(_iKFoM.get_x()) with the last IMU sample and stamp; v is rotated into the body frame."""
        out = self._state_f32()
        out["v"] = _RT_f32(out["q"], (0, 0, 0))[:3, :3].T @ out["v"]
        return out

    def getBodyState(self):
        """:157-172: the same in the LiDAR frame — p += pLI, q *= qLI (as the reference composes them), local velocity."""
        out = self._state_f32()
        out["p"] = out["p"] + out["pLI"]
        x1, y1, z1, w1 = out["q"]
        x2, y2, z2, w2 = out["qLI"]
        out["q"] = np.float32([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2,
                               w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2])
        out["v"] = _RT_f32(out["q"], (0, 0, 0))[:3, :3].T @ out["v"]
        return out

    def getPoseCovariance(self):
        """:214-229: 6x6 [orientation, position] blocks of P, flattened column-major like Eigen::Map."""
        P = self.P
        out = np.zeros((6, 6))
        out[0:3, 0:3], out[0:3, 3:6], out[3:6, 0:3], out[3:6, 3:6] = P[3:6, 3:6], P[3:6, 0:3], P[0:3, 3:6], P[0:3, 0:3]
        return out.flatten(order="F")

    def getTwistCovariance(self):
        """:231-244, as written: the block at (6, 6) — the extrinsic rotation, not the velocity at 12 — and cov_gyro."""
        out = np.zeros((6, 6))
        out[0:3, 0:3] = self.P[6:9, 6:9]
        out[3:6, 3:6] = self.config.cov_gyro * np.eye(3)
        return out.flatten(order="F")

    def get_pc2match_pointcloud(self):
        return self.map.prep_get(3)
