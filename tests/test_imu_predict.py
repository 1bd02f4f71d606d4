"""IMU rate: esekf::predict + the propagated-state ring (CPU tests, host algebra of libflimo_cuda).

Reference: IKFoM_toolkit/esekfom/esekfom.hpp:279-384 (predict), IKFoM/use-ikfom.cpp:46-91 (process model),
fast_limo/Modules/Localizer.cpp:583-608 (propagateImu), :855-913 (integrateImu / propagatedFromTimeRange).

Three independent evaluations are compared: the product (fast_limo_b200/csrc/imu_host.hpp, block-wise closed
form), the oracle (oracle/predict.hpp, the reference's generic dense shape) and numpy / scipy written here
(state transition with scipy rotations, Jacobians by finite differences).
"""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from fast_limo_b200 import api, synth

G = 9.809
COV = (6.e-4, 1.e-2, 1.e-5, 3.e-4)      # src/main.cpp:159-162


def _state(seed):
    rng = np.random.default_rng(seed)
    g = rng.normal(size=3)
    g[0] += 0.3                           # keep away from the chart's singular direction (-L, 0, 0)
    g *= G / np.linalg.norm(g)
    return synth.make_state(rng.normal(0, 5, 3), Rot.random(random_state=seed).as_quat(),
                            Rot.from_rotvec(rng.normal(0, 0.05, 3)).as_quat(), rng.normal(0, 0.1, 3), rng.normal(0, 3, 3),
                            rng.normal(0, 1e-2, 3), rng.normal(0, 5e-2, 3), g)


def _spd(seed, scale=1e-2):
    rng = np.random.default_rng(seed + 100)
    A = rng.normal(size=(23, 23))
    return scale * (A @ A.T / 23 + np.eye(23))


def _imu(seed, n, rate=200.0, t0=10.0):
    rng = np.random.default_rng(seed + 7)
    stamps = t0 + (1 + np.arange(n)) / rate
    acc = (rng.normal(0, 1.5, (n, 3)) + [0, 0, G]).astype(np.float32)
    gyr = rng.normal(0, 0.8, (n, 3)).astype(np.float32)
    return stamps, np.full(n, 1.0 / rate), acc, gyr


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_predict_matches_oracle(oracle, flimo_lib, seed):
    x0, P0 = _state(seed), _spd(seed)
    stamps, dts, acc, gyr = _imu(seed, 60)
    m = api.Mapper(device=-1)
    orc = oracle.Propagator(x0, P0)
    x, P = x0, P0
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        x, P = m.ekf_predict(x, P, t, dt, a, w, COV)
        orc.propagate(t, dt, a, w, COV)
    xo, Po = orc.get()
    assert np.allclose(x, xo, rtol=0, atol=1e-12 * max(1.0, np.abs(xo).max()))
    assert np.allclose(P, Po, rtol=1e-11, atol=1e-15)
    assert np.allclose(P, P.T, rtol=1e-12, atol=1e-15)
    assert np.linalg.eigvalsh(0.5 * (P + P.T)).min() > 0
    # quaternion stays (numerically) unit: no normalisation in either code, 60 products
    assert abs(np.linalg.norm(x[3:7]) - 1) < 1e-12
    # propagated states: the float casts of the SAME doubles => identical records unless the doubles differ in
    # the last bits; compare field-wise at float32 resolution
    fo = orc.frames(stamps[9] + 1e-4, stamps[40] + 1e-4)
    fp = m.propagated_frames(stamps[9] + 1e-4, stamps[40] + 1e-4)
    assert len(fo) == len(fp) == 33                      # samples 9 .. 41: one before the start, first one past the end
    assert np.array_equal(fo["time"], fp["time"]) and fo["time"][0] == stamps[9] and fo["time"][-1] == stamps[41]
    for k in ("w", "a"):
        assert np.array_equal(fo[k], fp[k])
    assert np.array_equal(fp["a"], acc[9:42]) and np.array_equal(fp["w"], gyr[9:42])
    for k in ("q", "p", "v", "bg", "ba", "g"):
        assert np.allclose(fo[k], fp[k], rtol=3e-7, atol=1e-9)


def _transition(x, acc, gyro, dt):
    """x (+) f(x, u) dt with scipy rotations: pos, rot (xyzw), offR, offT, vel, bg, ba, grav."""
    y = x.copy()
    R = Rot.from_quat(x[3:7])
    y[0:3] = x[0:3] + x[14:17] * dt
    y[3:7] = (R * Rot.from_rotvec((gyro - x[17:20]) * dt)).as_quat()
    y[14:17] = x[14:17] + (R.apply(acc - x[20:23]) + x[23:26]) * dt
    return y


def _extract_F(m, x0, acc, gyr, dt):
    """F of P' = F P F^T + G Q G^T through the C ABI: with Q = 0 and P = e_j e_j^T, P' = F[:, j] F[:, j]^T."""
    F = np.zeros((23, 23))
    for j in range(23):
        P = np.zeros((23, 23))
        P[j, j] = 1.0
        _, Pn = m.ekf_predict(x0, P, 0.0, dt, acc, gyr, (0, 0, 0, 0))
        col = Pn[:, j] / np.sqrt(Pn[j, j])                # F[j, j] > 0 (F = I + O(dt))
        assert np.allclose(Pn, np.outer(col, col), atol=1e-14)
        F[:, j] = col
    return F


@pytest.mark.parametrize("seed", [4, 5])
def test_transition_and_jacobian_against_numpy(oracle, flimo_lib, seed):
    x0 = _state(seed)
    rng = np.random.default_rng(seed)
    acc = (rng.normal(0, 2, 3) + [0, 0, G]).astype(np.float32)
    gyr = rng.normal(0, 1.0, 3).astype(np.float32)
    dt = 0.01
    m = api.Mapper(device=-1)
    x1, _ = m.ekf_predict(x0, _spd(seed), 0.0, dt, acc, gyr, COV)
    ref = _transition(x0, acc.astype(np.float64), gyr.astype(np.float64), dt)
    sgn = np.sign(np.dot(ref[3:7], x1[3:7]))
    ref[3:7] *= sgn
    assert np.allclose(x1, ref, atol=1e-12)

    # Jacobian of the transition in (+)/(-) coordinates by central differences, (+)/(-) from the oracle
    F = _extract_F(m, x0, acc, gyr, dt)
    eps = 1e-5
    J = np.zeros((23, 23))
    a64, w64 = acc.astype(np.float64), gyr.astype(np.float64)
    mid = _transition(x0, a64, w64, dt)                  # both differences in the chart of the unperturbed result
    for j in range(23):
        d = np.zeros(23)
        d[j] = eps
        hi = _transition(oracle.boxplus(x0, d), a64, w64, dt)
        lo = _transition(oracle.boxplus(x0, -d), a64, w64, dt)
        J[:, j] = (oracle.boxminus(hi, mid) - oracle.boxminus(lo, mid)) / (2 * eps)
    mask = np.ones((23, 23), bool)
    mask[3:6, 3:6] = False
    assert np.allclose(F[mask], J[mask], atol=2e-8)
    # the rot-rot block: the exact Jacobian is Exp(-(gyro - bg) dt); the reference (and therefore this library)
    # leaves the identity there because of scalar(1/2) == 0 at esekfom.hpp:312
    assert np.array_equal(F[3:6, 3:6], np.eye(3))
    assert np.allclose(J[3:6, 3:6], Rot.from_rotvec(-(w64 - x0[17:20]) * dt).as_matrix(), atol=2e-8)


def test_process_model_finite_differences(oracle):
    x0 = _state(9)
    acc, gyr = np.array([0.3, -1.0, 9.6]), np.array([0.2, -0.4, 0.7])
    f, fx, fw = oracle.process_model(x0, acc, gyr)
    assert np.allclose(f[0:3], x0[14:17]) and np.allclose(f[3:6], gyr - x0[17:20])
    assert np.allclose(f[12:15], Rot.from_quat(x0[3:7]).apply(acc - x0[20:23]) + x0[23:26])
    assert not f[6:12].any() and not f[15:].any()
    eps = 1e-6
    J = np.zeros((24, 23))
    for j in range(23):
        d = np.zeros(23)
        d[j] = eps
        J[:, j] = (oracle.process_model(oracle.boxplus(x0, d), acc, gyr)[0] -
                   oracle.process_model(oracle.boxplus(x0, -d), acc, gyr)[0]) / (2 * eps)
    assert np.allclose(fx, J, atol=1e-7)
    # noise enters as gyro - ng, acc - na (use-ikfom.cpp:82-90): df/dw by differences on the inputs
    Jw = np.zeros((24, 12))
    for j in range(3):
        d = np.zeros(3)
        d[j] = eps
        Jw[:, j] = (oracle.process_model(x0, acc, gyr - d)[0] - oracle.process_model(x0, acc, gyr + d)[0]) / (2 * eps)
        Jw[:, 3 + j] = (oracle.process_model(x0, acc - d, gyr)[0] - oracle.process_model(x0, acc + d, gyr)[0]) / (2 * eps)
    assert np.allclose(fw[:, :6], Jw[:, :6], atol=1e-7)
    assert np.array_equal(fw[15:18, 6:9], np.eye(3)) and np.array_equal(fw[18:21, 9:12], np.eye(3))


def test_noise_term(flimo_lib):
    """P = 0 isolates G Q G^T: rot <- gyro noise through A, vel <- accel noise through R, biases directly."""
    x0 = _state(11)
    acc, gyr = np.float32([0.1, 0.2, 9.7]), np.float32([0.5, -0.3, 0.2])
    dt = 0.005
    m = api.Mapper(device=-1)
    _, Pn = m.ekf_predict(x0, np.zeros((23, 23)), 0.0, dt, acc, gyr, COV)
    exp = np.zeros((23, 23))
    w = (gyr.astype(np.float64) - x0[17:20]) * dt
    th = np.linalg.norm(w)
    K = np.array([[0, w[2], -w[1]], [-w[2], 0, w[0]], [w[1], -w[0], 0]])      # hat(-w)
    A = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (1 - np.sin(th) / th) / th ** 2 * K @ K
    exp[3:6, 3:6] = COV[0] * dt * dt * A @ A.T
    exp[12:15, 12:15] = COV[1] * dt * dt * np.eye(3)
    exp[15:18, 15:18] = COV[2] * dt * dt * np.eye(3)
    exp[18:21, 18:21] = COV[3] * dt * dt * np.eye(3)
    assert np.allclose(Pn, exp, atol=1e-18, rtol=1e-10)


def _select(times_newest_first, t0, t1):
    """propagatedFromTimeRange restated on a plain list (Localizer.cpp:878-913)."""
    b = times_newest_first
    if not b or b[0] < t1:
        return None
    it, last = 1, 0
    while it < len(b) and b[it] >= t1:
        last = it
        it += 1
    while it < len(b) and b[it] >= t0:
        it += 1
    if it == len(b):
        return []
    return list(reversed(b[last:it + 1]))


def test_propagated_ring_selection(oracle, flimo_lib):
    x0, P0 = _state(21), _spd(21)
    m = api.Mapper(device=-1)
    orc = oracle.Propagator(x0, P0)
    n = 2300                                             # > capacity 2000: the oldest 300 fall off
    stamps, dts, acc, gyr = _imu(21, n, rate=400.0, t0=100.0)
    gyr *= 0.05
    acc = (acc - np.float32([0, 0, G])) * np.float32(0.05) + np.float32([0, 0, G])
    x, P = x0, P0
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        x, P = m.ekf_predict(x, P, t, dt, a, w, COV)
        orc.propagate(t, dt, a, w, COV)
    kept = list(stamps[::-1][:2000])
    windows = [(stamps[500] + 1e-5, stamps[540] + 1e-5),    # ordinary scan
               (stamps[500], stamps[540]),                  # boundaries exactly on samples (>= in both loops)
               (stamps[2250], stamps[2299]),                # ends on the newest sample
               (stamps[2298] + 1e-5, stamps[2299]),         # shortest window
               (0.0, stamps[700]),                          # first scan: prev_scan_stamp = 0 -> not enough states
               (stamps[100], stamps[400]),                  # starts before the oldest retained state (index 300)
               (stamps[300], stamps[400]),                  # starts ON the oldest retained state -> nothing older
               (stamps[301], stamps[400])]                  # one older state exists
    for t0, t1 in windows:
        exp = _select(kept, t0, t1)
        got = m.propagated_frames(t0, t1)
        go = orc.frames(t0, t1)
        assert go is not None and list(go["time"]) == exp
        assert list(got["time"]) == exp
    assert len(m.propagated_frames(0.0, stamps[700])) == 0
    assert len(m.propagated_frames(stamps[301], stamps[400])) == 101
    # the IMU thread is behind the scan: the reference blocks, the library reports it
    assert orc.frames(stamps[10], stamps[-1] + 0.01) is None
    with pytest.raises(api.FlimoError, match="IMU behind"):
        m.propagated_frames(stamps[10], stamps[-1] + 0.01)
    m.propagated_clear()
    with pytest.raises(api.FlimoError, match="IMU behind"):
        m.propagated_frames(stamps[500], stamps[540])


def test_frames_feed_the_deskew_oracle(oracle, flimo_lib):
    """The records of the ring are what deskewPointCloud consumes: oracle deskew on product frames == on oracle frames."""
    x0, P0 = _state(31), synth.default_P0()
    x0[14:17] = [8.0, 0.5, 0.0]
    m = api.Mapper(device=-1)
    orc = oracle.Propagator(x0, P0)
    stamps, dts, acc, gyr = _imu(31, 80, rate=400.0, t0=50.0)
    gyr *= 0.3
    x, P = x0, P0
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        x, P = m.ekf_predict(x, P, t, dt, a, w, COV)
        orc.propagate(t, dt, a, w, COV)
    t_begin, t_end = stamps[20] + 1e-4, stamps[60] + 1e-4
    fp, fo = m.propagated_frames(t_begin, t_end), orc.frames(t_begin, t_end)
    rng = np.random.default_rng(5)
    n = 4000
    raw = np.zeros(n, api.RAW_POINT)
    pts = rng.normal(0, 15, (n, 3)).astype(np.float32)
    raw["x"], raw["y"], raw["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    raw["time"] = np.sort(rng.uniform(0, t_end - t_begin, n)).astype(np.float32)
    cfg = oracle.make_prep_cfg(sensor_type=1)
    order = oracle.prep_filter_sort(raw, cfg, sort=True)
    T = np.eye(4, dtype=np.float32)
    lq, lp = x[3:7].astype(np.float32), x[0:3].astype(np.float32)
    wp, bp = oracle.prep_deskew(raw, order, cfg, t_begin, 0.0, fp, lq, lp, T)
    wo, bo = oracle.prep_deskew(raw, order, cfg, t_begin, 0.0, fo, lq, lp, T)
    assert np.allclose(wp, wo, atol=2e-5) and np.allclose(bp, bo, atol=2e-5)
    assert np.isfinite(wp).all()


def test_localizer_mirror_imu_side(flimo_lib):
    """Localizer::updateIMU over the synthetic stream's IMU: dead reckoning follows the trajectory, the ring hands out
    the frames of a scan interval, and the LiDAR callback refuses a host-only handle (no CPU path)."""
    from fast_limo_b200.localizer import Localizer, LocalizerConfig
    S = synth.Stream(rings=8, azimuths=64, imu_hz=200.0)
    x_true = S.state(0.0)
    m = api.Mapper(device=-1)
    loc = Localizer(m, LocalizerConfig(), pos=x_true[0:3], quat=x_true[3:7], vel=x_true[14:17])
    assert abs(np.linalg.norm(loc.x[23:26]) - 9.809) < 1e-12
    assert loc.updatePointCloud(np.zeros(4, api.RAW_POINT), 0.0) is False and loc.last["null"] == "IMU buffer is empty"
    stamps, dts, acc, gyr = S.imu(0.0, 1.0)
    assert len(stamps) == 200 and stamps[0] == 0.005 and stamps[-1] == 1.0
    for t, dt, a, w in zip(stamps, dts, acc, gyr):
        loc.updateIMU(t, dt, a, w)
    p, q, v = loc.x[0:3], loc.x[3:7], loc.x[14:17]
    truth = S.state(1.0)
    ws, bs = loc.getWorldState(), loc.getBodyState()
    assert abs(ws["v"][0] - S.v) < 2e-2 and abs(ws["v"][1]) < 2e-2 and ws["time"] == 1.0    # body-frame velocity: forward
    assert np.allclose(bs["p"], ws["p"]) and np.allclose(bs["v"], ws["v"], atol=1e-6)      # identity extrinsics here
    assert np.array_equal(ws["w"], gyr[-1]) and np.array_equal(ws["a"], acc[-1])
    pc = loc.getPoseCovariance().reshape(6, 6, order="F")
    assert np.array_equal(pc[3:6, 3:6], loc.P[0:3, 0:3]) and np.array_equal(pc[0:3, 3:6], loc.P[3:6, 0:3])
    tc = loc.getTwistCovariance().reshape(6, 6, order="F")
    assert np.array_equal(tc[0:3, 0:3], loc.P[6:9, 6:9]) and tc[4, 4] == loc.config.cov_gyro
    assert np.linalg.norm(p - truth[0:3]) < 2e-2          # explicit Euler over 200 steps: a few mm
    assert np.linalg.norm(v - truth[14:17]) < 2e-2
    assert abs(np.dot(q, truth[3:7])) > 1 - 1e-10         # constant rate: the orientation is exact
    assert np.all(np.diag(loc.P)[:3] > 1.0)               # position uncertainty grew from the initial 1.0
    fr = m.propagated_frames(0.5, 0.6)
    ideal = S.frames(0.5, 0.6)
    assert len(fr) == 22 and abs(fr["time"][0] - 0.495) < 1e-12 and abs(fr["time"][-1] - 0.6) < 1e-12
    k = {round(t, 6): i for i, t in enumerate(ideal["time"])}
    sel = [k[round(t, 6)] for t in fr["time"]]
    assert np.abs(fr["p"] - ideal["p"][sel]).max() < 1e-2 and np.abs(fr["q"] - ideal["q"][sel]).max() < 1e-6
    assert np.allclose(fr["w"], ideal["w"][sel]) and np.allclose(fr["a"], ideal["a"][sel], atol=1e-6)
    raw, stamp = S.scan(3)
    with pytest.raises(api.FlimoError, match="host-only"):
        loc.updatePointCloud(raw, stamp)


def test_ring_selection_property(flimo_lib):
    """hypothesis: for arbitrary (irregular, possibly repeated) stamps and windows the ring returns exactly what the
    reference's iterator walk returns."""
    from hypothesis import given, settings, strategies as st
    m = api.Mapper(device=-1)
    x0, P0 = _state(41), synth.default_P0()
    a, w = np.float32([0, 0, G]), np.float32([0, 0, 0.1])

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.integers(0, 3), min_size=0, max_size=40), st.integers(-5, 130), st.integers(-5, 130))
    def check(gaps, i0, i1):
        m.propagated_clear()
        t, stamps = 0.0, []
        for gp in gaps:                                   # gaps of 0 repeat a stamp (two IMU messages with one time)
            t += gp * 0.25
            stamps.append(t)
            m.ekf_predict(x0, P0, t, 0.005, a, w, COV)
        t0, t1 = i0 * 0.125, i1 * 0.125                   # on and between the sample times
        exp = _select(stamps[::-1], t0, t1)
        if exp is None:
            with pytest.raises(api.FlimoError, match="IMU behind"):
                m.propagated_frames(t0, t1)
        else:
            assert list(m.propagated_frames(t0, t1)["time"]) == exp

    check()


def test_predict_keeps_covariance_symmetric_positive(flimo_lib):
    """2000 predictions (10 s at 200 Hz) from the reference's initial covariance: P stays symmetric positive definite
    and grows monotonically in the unobserved directions."""
    m = api.Mapper(device=-1)
    x, P = _state(43), synth.default_P0()
    rng = np.random.default_rng(43)
    tr = [np.trace(P)]
    for i in range(2000):
        a = (rng.normal(0, 0.5, 3) + [0, 0, G]).astype(np.float32)
        w = rng.normal(0, 0.3, 3).astype(np.float32)
        x, P = m.ekf_predict(x, P, 0.005 * (i + 1), 0.005, a, w, COV)
        if i % 500 == 499:
            assert np.abs(P - P.T).max() <= 1e-12 * np.abs(P).max()
            assert np.linalg.eigvalsh(0.5 * (P + P.T)).min() > 0
            tr.append(np.trace(P))
    assert all(b > a_ for a_, b in zip(tr, tr[1:]))
    assert abs(np.linalg.norm(x[3:7]) - 1) < 1e-10 and abs(np.linalg.norm(x[23:26]) - G) < 1e-9


def test_raw_imu_to_baselink(flimo_lib):
    """imu2baselink (Localizer.cpp:697-728): rotation into the base-link frame, lever arm, dt fallback, intrinsic correction."""
    from fast_limo_b200.localizer import Localizer
    m = api.Mapper(device=-1)
    loc = Localizer(m, bias_accel=(0.1, 0, 0), bias_gyro=(0, 0, 0.5))    # this->state.b (Localizer.cpp:67-68)
    Rz = Rot.from_euler("z", 90, degrees=True).as_matrix()
    lever = np.array([0.5, 0.0, 0.0])
    # first sample: dt = stamp - 0 > 0.1 -> 1/200; no angular acceleration term (previous rate := current)
    loc.updateIMU_raw(10.0, [1.0, 0.0, 9.809], [0.0, 0.0, 2.0], Rz, lever)
    a, w = loc.last_imu
    # R a = (0, 1, 9.809); centripetal w x (w x -t) with w = (0,0,2), t = (0.5,0,0): +4 * 0.5 along x
    assert np.allclose(a, [0.0 + 2.0 - 0.1, 1.0, 9.809], atol=1e-5) and np.allclose(w, [0, 0, 1.5], atol=1e-6)
    fr = m.propagated_frames(0.0, 10.0)
    assert len(fr) == 0                                       # only one propagated state, nothing older than the window
    # second sample 10 ms later with a changed rate: Euler term (dw/dt) x (-t)
    loc.updateIMU_raw(10.01, [1.0, 0.0, 9.809], [0.0, 0.0, 3.0], Rz, lever)
    a2, w2 = loc.last_imu
    alpha = (3.0 - 2.0) / np.float32(0.01)
    assert np.allclose(w2, [0, 0, 3.0 - 0.5], atol=1e-6)
    assert np.allclose(a2, [0.0 + 9.0 * 0.5 - 0.1, 1.0 - alpha * 0.5, 9.809], rtol=2e-4, atol=1e-3)
    assert abs(loc.imu_stamp - 10.01) < 1e-12 and loc.n_imu == 2
