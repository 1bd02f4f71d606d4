"""CPU: the oracle's manifold ops and iterated update against independent numpy/scipy code."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from fast_limo_b200 import synth

G = 9.809


def _rand_state(rng):
    q = Rot.random(random_state=int(rng.integers(1 << 30))).as_quat()
    q2 = (Rot.from_rotvec(rng.normal(0, 0.05, 3))).as_quat()
    g = rng.normal(size=3)
    g = g / np.linalg.norm(g) * G
    if g[0] < -0.9 * G:
        g[0] = -g[0]
    return synth.make_state(rng.normal(0, 5, 3), q, q2, rng.normal(0, 0.2, 3), rng.normal(0, 2, 3),
                            rng.normal(0, 0.01, 3), rng.normal(0, 0.1, 3), g)


def test_boxplus_matches_scipy(oracle):
    rng = np.random.default_rng(0)
    for _ in range(50):
        x = _rand_state(rng)
        d = rng.normal(0, 0.3, 23)
        y = oracle.boxplus(x, d)
        assert np.allclose(y[0:3], x[0:3] + d[0:3])
        assert np.allclose(y[11:14], x[11:14] + d[9:12])
        for qo, do in ((3, 3), (7, 6)):
            ref = (Rot.from_quat(x[qo:qo + 4]) * Rot.from_rotvec(d[do:do + 3])).as_quat()
            got = y[qo:qo + 4]
            assert min(np.abs(got - ref).max(), np.abs(got + ref).max()) < 1e-12
        assert abs(np.linalg.norm(y[23:26]) - G) < 1e-9          # S2 stays on the sphere
        ang = np.arccos(np.clip(y[23:26] @ x[23:26] / G / G, -1, 1))
        assert abs(ang - np.linalg.norm(d[21:23])) < 1e-9        # rotated by |delta| (Bx orthonormal)


def test_boxminus_inverts_boxplus(oracle):
    rng = np.random.default_rng(1)
    for _ in range(50):
        x = _rand_state(rng)
        d = rng.normal(0, 0.2, 23)
        y = oracle.boxplus(x, d)
        back = oracle.boxminus(y, x)
        assert np.allclose(back, d, atol=1e-9)
    assert np.allclose(oracle.boxminus(x, x), 0, atol=1e-15)


def test_invert(oracle):
    rng = np.random.default_rng(2)
    A = rng.normal(size=(23, 23))
    A = A @ A.T + np.eye(23)
    assert np.allclose(oracle.invert(A), np.linalg.inv(A), rtol=1e-10, atol=1e-12)


# ---- an independent numpy transcription of SURVEY Appendix A (n >= 23 branch) --------------------
def _hat(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _A(v):
    n = np.linalg.norm(v)
    if n < 1e-11:
        return np.eye(3)
    K = _hat(v)
    return np.eye(3) + (1 - np.cos(n)) / n ** 2 * K + (1 - np.sin(n) / n) / n ** 2 * K @ K


def _Bx(v):
    d = G + v[0]
    return np.array([[-v[1], -v[2]], [G - v[1] ** 2 / d, -v[2] * v[1] / d], [-v[2] * v[1] / d, G - v[2] ** 2 / d]]) / G


def _np_update(oracle, x, P, max_iter, limit, H, h, R=0.001):
    x = x.copy()
    xp, Pp = x.copy(), P.copy()
    t = 0
    HTH, HTh = H.T @ H, H.T @ h
    for i in range(-1, max_iter):
        dx = oracle.boxminus(x, xp)          # manifold ops are tested separately above
        dxn = dx.copy()
        P = Pp.copy()
        for idx in (3, 6):
            J = _A(dx[idx:idx + 3]).T
            dxn[idx:idx + 3] = J @ dxn[idx:idx + 3]
            P[idx:idx + 3, :] = J @ P[idx:idx + 3, :]
            P[:, idx:idx + 3] = P[:, idx:idx + 3] @ J.T
        Nx = _Bx(x[23:26]).T @ _hat(x[23:26]) / G ** 2
        d2 = dx[21:23]
        Bp = _Bx(xp[23:26])
        Mx = -_hat(xp[23:26]) @ Bp if np.linalg.norm(d2) < 1e-11 else -_hat(xp[23:26]) @ _A(Bp @ d2).T @ Bp
        J2 = Nx @ Mx
        dxn[21:23] = J2 @ dxn[21:23]
        P[21:23, :] = J2 @ P[21:23, :]
        P[:, 21:23] = P[:, 21:23] @ J2.T
        Pt = np.linalg.inv(P / R)
        Pt[:12, :12] += HTH
        Pi = np.linalg.inv(Pt)
        Kh = Pi[:, :12] @ HTh
        Kx = np.zeros((23, 23))
        Kx[:, :12] = Pi[:, :12] @ HTH
        dx_ = Kh + (Kx - np.eye(23)) @ dxn
        x = oracle.boxplus(x, dx_)           # non-degenerate: filter is the identity
        if np.all(np.abs(dx_) <= limit):
            t += 1
        if t > 1 or i == max_iter - 1:
            L = P.copy()
            for idx in (3, 6):
                J = _A(dx_[idx:idx + 3]).T
                L[idx:idx + 3, :] = J @ P[idx:idx + 3, :]
                Kx[idx:idx + 3, :12] = J @ Kx[idx:idx + 3, :12]
                L[:, idx:idx + 3] = L[:, idx:idx + 3] @ J.T
                P[:, idx:idx + 3] = P[:, idx:idx + 3] @ J.T
            Nx = _Bx(x[23:26]).T @ _hat(x[23:26]) / G ** 2
            d2 = dx_[21:23]
            Mx = -_hat(xp[23:26]) @ Bp if np.linalg.norm(d2) < 1e-11 else -_hat(xp[23:26]) @ _A(Bp @ d2).T @ Bp
            J2 = Nx @ Mx
            L[21:23, :] = J2 @ P[21:23, :]
            Kx[21:23, :12] = J2 @ Kx[21:23, :12]
            L[:, 21:23] = L[:, 21:23] @ J2.T
            P[:, 21:23] = P[:, 21:23] @ J2.T
            return x, L - Kx[:, :12] @ P[:12, :]
    return x, P


@pytest.mark.parametrize("max_iter,limit", [(0, 0.001), (2, 0.0), (3, 0.001), (4, 1e9)])
def test_update_matches_numpy_transcription(oracle, max_iter, limit):
    rng = np.random.default_rng(3)
    x0 = _rand_state(rng)
    x0[23:26] = [0.3, -0.2, -np.sqrt(G * G - 0.13)]
    P0 = synth.default_P0()
    H = rng.normal(size=(600, 12))
    H[:, 6:] *= 0.2
    h = rng.normal(0, 0.03, 600)
    xo, Po, tr = oracle.update_fixed(x0, P0, max_iter, limit, H, h)
    xn, Pn = _np_update(oracle, x0, P0, max_iter, limit, H, h)
    assert np.allclose(xo, xn, rtol=0, atol=1e-10)
    assert np.allclose(Po, Pn, rtol=1e-7, atol=1e-12)
    assert len(tr) <= max_iter + 1
    if limit >= 1e9:
        assert len(tr) == min(2, max_iter + 1)      # converged twice -> early exit (t > 1)


def test_update_few_rows_and_empty(oracle):
    rng = np.random.default_rng(4)
    x0 = _rand_state(rng)
    P0 = synth.default_P0()
    # no measurement at all: state unchanged, covariance unchanged (first mapped scan)
    xo, Po, tr = oracle.update_fixed(x0, P0, 2, 0.001, np.zeros((0, 12)), np.zeros(0))
    assert np.allclose(xo, x0, atol=1e-15) and np.allclose(Po, P0, atol=1e-12) and len(tr) == 2
    # fewer rows than states: pose block is frozen by the degeneracy filter (HTH restated as 0)
    H = rng.normal(size=(10, 12))
    h = rng.normal(0, 0.03, 10)
    xo, Po, tr = oracle.update_fixed(x0, P0, 2, 0.001, H, h)
    assert np.allclose(xo[:7], x0[:7], atol=1e-15)
    assert not np.allclose(xo[7:14], x0[7:14], atol=1e-12)
