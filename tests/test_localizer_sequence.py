"""Closed loop of the two Localizer callbacks on the CPU (sequencing test).

fast_limo_b200.localizer.Localizer only sequences calls into an `api.Mapper`.  Here the GPU stages of that object
(filters / sort, deskew, voxel grid, iterated update, world transform, map insert) are answered by the CPU ORACLE —
the checker standing in for the device, which this test may do and the product may not — while the IMU side
(esekf::predict, propagated-state ring) is the library's own host algebra on a host-only handle.  What is checked is
the call sequence of Localizer.cpp:245-399 / :401-531: the dropped first scan, the map initialised by the second
one with zero matches, the carried covariance, the stamps, and that the estimate follows the trajectory without
any ground truth on the way in.  The same loop on the device: tests/test_gpu_experimental.py::test_localizer_closed_loop.
"""
import numpy as np

from fast_limo_b200 import api, synth
from fast_limo_b200.localizer import Localizer, LocalizerConfig

BIG = 1 << 18
ACC_BIAS = np.float32([0.3, 0.0, 0.0])


class OracleStages:
    """Duck-typed api.Mapper: IMU side from libflimo_cuda (host-only handle), scan side from the oracle."""

    def __init__(self, oracle, leaf):
        self.O = oracle
        self.host = api.Mapper(device=-1)
        self.om = oracle.OracleMap()
        self.ocfg = oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4)
        self.leaf = leaf
        self.pc2match = None
        self.calls = []

    # IMU side: the product's host algebra
    def ekf_predict(self, *a, **k):
        return self.host.ekf_predict(*a, **k)

    def propagated_frames(self, t0, t1):
        return self.host.propagated_frames(t0, t1)

    def propagated_clear(self):
        self.host.propagated_clear()

    # scan side: oracle
    def prep_filter_sort(self, raw, sweep_ref_time, filters):
        self.calls.append("filter_sort")
        f = filters
        crop = (list(f.cropBoxMin), list(f.cropBoxMax)) if f.cropBoxMin is not None else None
        self.pcfg = self.O.make_prep_cfg(crop=crop, min_dist=f.min_dist, rate=f.rate_value, fov=f.fov_angle,
                                         sensor_type=f.sensor_type, end_of_sweep=f.end_of_sweep, leaf=f.leafSize)
        self.raw, self.ref_time = raw, sweep_ref_time
        self.order = self.O.prep_filter_sort(raw, self.pcfg, sort=True)
        if len(self.order) == 0:
            return 0, 0.0
        return len(self.order), float(self.O.prep_times(raw, self.order[-1:], self.pcfg, sweep_ref_time)[0])

    def prep_deskew(self, frames, last_q, last_p, T, offset=0.0):
        self.calls.append("deskew")
        _, body = self.O.prep_deskew(self.raw, self.order, self.pcfg, self.ref_time, offset, frames, last_q, last_p, T)
        pc = self.O.prep_voxel(body, self.leaf) if self.leaf else body
        self.pc2match = np.ascontiguousarray(pc[:, :3])
        return len(self.pc2match)

    def update(self, x, P, max_iter, limits):
        self.calls.append("update")
        xo, Po, tr = self.om.update(self.ocfg, x, P, max_iter, limits, self.pc2match)
        self.rows = [t["rows"] for t in tr]
        return xo, Po, len(tr)

    def scan_to_world(self, x):
        self.calls.append("to_world")
        R = synth.quat_to_R(x[3:7].astype(np.float32)).astype(np.float32)
        return (self.pc2match @ R.T + x[0:3].astype(np.float32)).astype(np.float32)

    def add(self, pts, stamp):
        self.calls.append("add")
        self.om.add(np.ascontiguousarray(pts))
        self.stamp = stamp

    def add_scan(self, x, stamp):
        self.add(self.scan_to_world(x), stamp)

    def exists(self):
        return self.om.size() > 0

    def size(self):
        return self.om.size()


def test_closed_loop_sequence(oracle, flimo_lib):
    S = synth.Stream(rings=64, azimuths=512, imu_hz=200.0)   # sparser rings register to their own pattern (cm-level lag)
    stages = OracleStages(oracle, leaf=0.5)
    filt = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=0.5, sensor_type=1)
    x0 = S.state(0.0)
    loc = Localizer(stages, LocalizerConfig(filters=filt, MAX_NUM_ITERS=3), pos=x0[0:3], quat=x0[3:7], vel=x0[14:17])
    t_imu, errs, dead = 0.0, [], []
    x_dead, P_dead = loc.x.copy(), loc.P.copy()          # the same IMU stream without any LiDAR correction
    dead_host = api.Mapper(device=-1)
    for k in range(8):
        raw, stamp = S.scan(k)
        t_need = stamp + S.dt + 1.0 / S.imu_hz
        for smp in zip(*S.imu(t_imu, t_need, sigma_acc=0.05, sigma_gyro=0.002)):
            smp = (smp[0], smp[1], smp[2] + ACC_BIAS, smp[3])        # an uncalibrated accelerometer bias
            loc.updateIMU(*smp)
            x_dead, P_dead = dead_host.ekf_predict(x_dead, P_dead, *smp)
        t_imu = t_need
        stages.calls.clear()
        P_before = loc.P.copy()
        ok = loc.updatePointCloud(raw, stamp)
        truth = S.state(loc.imu_stamp)
        errs.append(float(np.linalg.norm(loc.x[0:3] - truth[0:3])))
        dead.append(float(np.linalg.norm(x_dead[0:3] - truth[0:3])))
        if k == 0:
            # prev_scan_stamp = 0: no propagated state is older than the window -> integrateImu is empty -> NULL iteration
            assert ok is False and loc.last["null"].startswith("no frames") and stages.calls == ["filter_sort"]
            assert stages.size() == 0 and np.array_equal(loc.P, P_before)
            assert loc.prev_scan_stamp == loc.scan_stamp > 0.09
        elif k == 1:
            # Mapper::match on the empty map: every pass has zero rows, the state stays at the prediction, then Mapper::add
            assert ok and stages.calls == ["filter_sort", "deskew", "update", "to_world", "add"]
            assert all(r == 0 for r in stages.rows) and stages.size() > 0
        else:
            assert ok and stages.rows[0] > 3000 and loc.last["passes"] >= 2
            assert np.trace(loc.P[:6, :6]) < np.trace(P_before[:6, :6])       # the measurement tightened the pose
        assert abs(loc.scan_stamp - (stamp + S.dt)) < 2e-3 and stages.calls.count("add") == (1 if ok else 0)
        if ok:
            assert stages.stamp == loc.scan_stamp
    # the registered estimate stays at the centimetre level while dead reckoning with the same biased IMU drifts
    # (0.5 * 0.3 m/s^2 * (0.8 s)^2 = 9.6 cm)
    assert max(errs[2:]) < 0.03, errs
    assert dead[-1] > 0.08 and errs[-1] < 0.25 * dead[-1], (errs, dead)
    assert 15 < loc.last["n_frames"] < 30 and loc.last["n_pc2match"] > 4000
