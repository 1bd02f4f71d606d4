"""Seeded synthetic worlds, maps and LiDAR scans for the parity tests and bench.py.

SURVEY.md section 8(d): the world is made of planes (ground, an enclosure of four walls,
axis-aligned boxes) so that the reference's plane gate (``Plane.cpp:107-114``, 0.05 m)
passes; map samples carry sigma = 0.01 m Gaussian noise along the surface normal and a
small uniform jitter on every coordinate so that no two kNN candidates tie
(``Octree.hpp:73,80`` keeps the earlier-visited point on ties — order dependent).

Everything is generated with ``numpy.random.default_rng(seed)`` (PCG64): identical on the
dev container and on the GPU box (same image).  Nothing here reads /root/reference.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class World:
    half: float                    # ground plane spans [-half, half]^2 at z = 0
    wall: float                    # enclosure walls at x = +-wall, y = +-wall
    wall_h: float
    boxes: np.ndarray              # (B, 6) xmin ymin zmin xmax ymax zmax
    rects: list = field(default_factory=list)   # (origin3, u3, v3, normal3, area)


def make_world(seed, half=120.0, wall=70.0, wall_h=25.0, n_boxes=60):
    rng = np.random.default_rng(seed)
    boxes = []
    tries = 0
    while len(boxes) < n_boxes and tries < 100000:
        tries += 1
        cx, cy = rng.uniform(-half + 12, half - 12, 2)
        sx, sy = rng.uniform(3.0, 18.0, 2)
        h = rng.uniform(2.0, 20.0)
        if abs(cx) < 6 + sx / 2 and abs(cy) < 6 + sy / 2:
            continue                      # keep the sensor neighbourhood free
        if abs(abs(cx) - wall) < sx / 2 + 1 or abs(abs(cy) - wall) < sy / 2 + 1:
            continue                      # do not straddle the enclosure
        boxes.append([cx - sx / 2, cy - sy / 2, 0.0, cx + sx / 2, cy + sy / 2, h])
    w = World(half, wall, wall_h, np.asarray(boxes, np.float64).reshape(-1, 6))
    R = w.rects

    def rect(o, u, v):
        o, u, v = (np.asarray(a, np.float64) for a in (o, u, v))
        n = np.cross(u, v)
        area = np.linalg.norm(n)
        R.append((o, u, v, n / area, area))

    rect([-half, -half, 0], [2 * half, 0, 0], [0, 2 * half, 0])                       # ground
    for s in (-1, 1):
        rect([s * wall, -wall, 0], [0, 2 * wall, 0], [0, 0, wall_h])                  # x = +-wall
        rect([-wall, s * wall, 0], [2 * wall, 0, 0], [0, 0, wall_h])                  # y = +-wall
    for b in w.boxes:
        x0, y0, z0, x1, y1, z1 = b
        rect([x0, y0, z1], [x1 - x0, 0, 0], [0, y1 - y0, 0])                          # roof
        rect([x0, y0, z0], [x1 - x0, 0, 0], [0, 0, z1 - z0])
        rect([x0, y1, z0], [x1 - x0, 0, 0], [0, 0, z1 - z0])
        rect([x0, y0, z0], [0, y1 - y0, 0], [0, 0, z1 - z0])
        rect([x1, y0, z0], [0, y1 - y0, 0], [0, 0, z1 - z0])
    return w


def sample_map(world, n_points, seed, sigma=0.01, jitter=1e-3):
    """Area-proportional surface samples (float32, shape (n, 3))."""
    rng = np.random.default_rng(seed)
    areas = np.array([r[4] for r in world.rects])
    counts = rng.multinomial(n_points, areas / areas.sum())
    out = np.empty((n_points, 3), np.float64)
    k = 0
    for (o, u, v, n, _), c in zip(world.rects, counts):
        if c == 0:
            continue
        a = rng.random((c, 1))
        b = rng.random((c, 1))
        out[k:k + c] = o + a * u + b * v + rng.normal(0.0, sigma, (c, 1)) * n
        k += c
    out += rng.uniform(-jitter, jitter, out.shape)
    rng.shuffle(out, axis=0)
    return out.astype(np.float32)


def _raycast(world, origin, dirs):
    """Nearest positive hit distance of rays (origin + t*dirs) with the world surfaces."""
    t_best = np.full(dirs.shape[0], np.inf)
    ox, oy, oz = origin
    dx, dy, dz = dirs[:, 0], dirs[:, 1], dirs[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -oz / dz                                                     # ground
        hx, hy = ox + t * dx, oy + t * dy
        ok = (t > 1e-6) & (np.abs(hx) <= world.half) & (np.abs(hy) <= world.half)
        t_best = np.where(ok & (t < t_best), t, t_best)
        for s in (-1.0, 1.0):                                            # enclosure, seen from inside
            t = (s * world.wall - ox) / dx
            hy, hz = oy + t * dy, oz + t * dz
            ok = (t > 1e-6) & (np.abs(hy) <= world.wall) & (hz >= 0) & (hz <= world.wall_h)
            t_best = np.where(ok & (t < t_best), t, t_best)
            t = (s * world.wall - oy) / dy
            hx, hz = ox + t * dx, oz + t * dz
            ok = (t > 1e-6) & (np.abs(hx) <= world.wall) & (hz >= 0) & (hz <= world.wall_h)
            t_best = np.where(ok & (t < t_best), t, t_best)
        for b in world.boxes:                                            # slabs
            t0x, t1x = (b[0] - ox) / dx, (b[3] - ox) / dx
            t0y, t1y = (b[1] - oy) / dy, (b[4] - oy) / dy
            t0z, t1z = (b[2] - oz) / dz, (b[5] - oz) / dz
            tn = np.maximum(np.maximum(np.minimum(t0x, t1x), np.minimum(t0y, t1y)), np.minimum(t0z, t1z))
            tf = np.minimum(np.minimum(np.maximum(t0x, t1x), np.maximum(t0y, t1y)), np.maximum(t0z, t1z))
            ok = (tn <= tf) & (tn > 1e-6)
            t_best = np.where(ok & (tn < t_best), tn, t_best)
    return t_best


def quat_from_rpy(roll, pitch, yaw):
    cr, sr = np.cos(roll / 2), np.sin(roll / 2)
    cp, sp = np.cos(pitch / 2), np.sin(pitch / 2)
    cy, sy = np.cos(yaw / 2), np.sin(yaw / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])            # x, y, z, w


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def spinning_scan(world, pos, quat, rings, azimuths, seed, fov_up_deg=2.0, fov_down_deg=-24.8, sigma=0.01,
                  max_range=150.0):
    """Ring-major spinning-LiDAR scan in the BODY frame (float32, (rings*azimuths, 3)).

    Rays that hit nothing (sky) are replaced by the previous valid return so the scan always
    has exactly rings*azimuths points (a real driver would drop them; the benchmark wants a
    fixed N).  Range noise sigma along the ray.
    """
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(fov_up_deg, fov_down_deg, rings))
    az = np.linspace(0.0, 2 * np.pi, azimuths, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_body = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (rings, azimuths))], -1)
    d_body = d_body.reshape(-1, 3)
    Rwb = quat_to_R(np.asarray(quat, np.float64))
    d_world = d_body @ Rwb.T
    t = _raycast(world, np.asarray(pos, np.float64), d_world)
    bad = ~np.isfinite(t) | (t > max_range)
    if bad.all():
        raise ValueError("no ray hits the world")
    idx = np.where(~bad, np.arange(t.size), 0)
    np.maximum.accumulate(idx, out=idx)
    first_ok = int(np.argmax(~bad))
    idx[:first_ok] = first_ok
    t = t[idx] + rng.normal(0.0, sigma, t.size)
    pts = d_body[idx] * t[:, None]
    pts += rng.uniform(-1e-4, 1e-4, pts.shape)
    return pts.astype(np.float32)


def rosette_scan(world, pos, quat, n_points, seed, cone_deg=70.0, sigma=0.01):
    """Livox-style non-repetitive pattern: golden-angle spiral inside a forward cone."""
    rng = np.random.default_rng(seed)
    i = np.arange(n_points) + 0.5
    half = np.deg2rad(cone_deg / 2)
    r = half * np.sqrt(i / n_points)
    th = i * np.pi * (3 - np.sqrt(5))
    d_body = np.stack([np.cos(r), np.sin(r) * np.cos(th), np.sin(r) * np.sin(th) - 0.15], -1)
    d_body /= np.linalg.norm(d_body, axis=1, keepdims=True)
    Rwb = quat_to_R(np.asarray(quat, np.float64))
    t = _raycast(world, np.asarray(pos, np.float64), d_body @ Rwb.T)
    bad = ~np.isfinite(t)
    idx = np.where(~bad, np.arange(t.size), 0)
    np.maximum.accumulate(idx, out=idx)
    idx[:int(np.argmax(~bad))] = int(np.argmax(~bad))
    t = t[idx] + rng.normal(0.0, sigma, t.size)
    return (d_body[idx] * t[:, None]).astype(np.float32)


def make_state(pos, quat, offR=(0, 0, 0, 1), offT=(0, 0, 0), vel=(0, 0, 0), bg=(0, 0, 0), ba=(0, 0, 0),
               grav=(0, 0, -9.809)):
    """Flat 26-double state_ikfom: pos rot(xyzw) offset_R_L_I(xyzw) offset_T_L_I vel bg ba grav."""
    return np.concatenate([np.asarray(a, np.float64) for a in (pos, quat, offR, offT, vel, bg, ba, grav)])


def default_P0():
    """init_iKFoM_state covariance (Localizer.cpp:685-693)."""
    d = np.ones(23)
    d[6:12] = 1e-6
    d[15:18] = 1e-5
    d[18:21] = 1e-4
    d[21:23] = 1e-6
    return np.diag(d)


@dataclass
class Case:
    name: str
    map_pts: np.ndarray
    scan: np.ndarray
    truth: np.ndarray       # state26 at the true pose
    init: np.ndarray        # state26 at the perturbed (predicted) pose
    passes: int


_CFG = {
    # name: (seed, half, wall, boxes, map points, rings, azimuths, passes)
    "tiny": (1000, 30.0, 20.0, 6, 20_000, 8, 256, 3),
    "c1": (1001, 40.0, 30.0, 10, 100_000, 16, 1024, 1),
    "c2": (1002, 120.0, 70.0, 60, 5_000_000, 64, 2048, 3),
    "c5": (1005, 100.0, 60.0, 40, 2_000_000, 32, 1024, 3),
}


def make_case(name, map_points=None):
    """BASELINE.json configs as seeded synthetic cases: truth + (0.05 m, 0.5 deg)."""
    if name == "c4":
        seed, half, wall, nb, m = 1004, 160.0, 90.0, 90, 20_000_000
        world = make_world(seed, half, wall, 25.0, nb)
        pos = np.array([0.3, -0.2, 1.8])
        quat = quat_from_rpy(0.01, -0.02, 0.4)
        scan = rosette_scan(world, pos, quat, 300_000, seed + 2)
        passes = 3
    else:
        seed, half, wall, nb, m, rings, az, passes = _CFG[name]
        world = make_world(seed, half, wall, 25.0, nb)
        pos = np.array([0.3, -0.2, 1.8])
        quat = quat_from_rpy(0.01, -0.02, 0.4)
        scan = spinning_scan(world, pos, quat, rings, az, seed + 2)
    if map_points is not None:
        m = int(map_points)
    mp = sample_map(world, m, seed + 1)
    truth = make_state(pos, quat)
    dq = quat_from_rpy(np.deg2rad(0.3), np.deg2rad(-0.25), np.deg2rad(0.3))      # ~0.5 deg total
    x, y, z, w = quat
    a, b, c, d = dq
    q2 = np.array([w * a + x * d + y * c - z * b, w * b - x * c + y * d + z * a, w * c + x * b - y * a + z * d,
                   w * d - x * a - y * b - z * c])
    init = make_state(pos + np.array([0.03, -0.03, 0.025]), q2 / np.linalg.norm(q2))
    return Case(name, mp, scan, truth, init, passes)


# ---- raw LiDAR messages + IMU-propagated frames for the scan preparation (deskew) tests ---------------
def make_raw_message(scan_xyz, sensor_type=1, sweep=0.1, stamp=100.0, rings=None, seed=7, n_nan=0):
    """fast_limo::Point records (32 bytes) for a scan: per-point time grows with the azimuth index
    (ring-major spinning scans: point i of a ring fires at sweep * i / azimuths).  Returns a numpy
    record array with the dtype of api.RAW_POINT."""
    from .api import RAW_POINT
    n = scan_xyz.shape[0]
    raw = np.zeros(n, RAW_POINT)
    raw["x"], raw["y"], raw["z"] = scan_xyz[:, 0], scan_xyz[:, 1], scan_xyz[:, 2]
    rng = np.random.default_rng(seed)
    raw["intensity"] = rng.uniform(0, 255, n).astype(np.float32)
    if rings:
        az = n // rings
        frac = (np.arange(n) % az) / az
    else:
        frac = np.arange(n) / n
    frac = frac + rng.uniform(0, 0.2 / n, n)                      # no exact ties in time
    if sensor_type == 0:
        raw["t"] = (frac * sweep * 1e9).astype(np.uint32)
    elif sensor_type == 1:
        raw["time"] = (frac * sweep).astype(np.float32)
    elif sensor_type == 2:
        raw["timestamp"] = stamp + frac * sweep
    else:
        raw["timestamp"] = (stamp + frac * sweep) * 1e9
    if n_nan:
        bad = rng.choice(n, n_nan, replace=False)
        raw["x"][bad[: n_nan // 2]] = np.nan
        raw["z"][bad[n_nan // 2:]] = np.inf
    return raw


def make_frames(t0, t1, rate_hz=200.0, speed=10.0, yaw_rate=0.6, seed=3):
    """IMU-propagated states (what Localizer::integrateImu returns) covering [t0, t1]: constant
    forward speed + yaw rate with small accelerations / biases, float32 like fast_limo::State."""
    from .api import FRAME
    rng = np.random.default_rng(seed)
    n = int(np.ceil((t1 - t0) * rate_hz)) + 3
    fr = np.zeros(n, FRAME)
    ts = t0 - 1.0 / rate_hz + np.arange(n) / rate_hz
    fr["time"] = ts
    yaw = yaw_rate * (ts - t0)
    fr["q"][:, 2], fr["q"][:, 3] = np.sin(yaw / 2), np.cos(yaw / 2)
    fr["p"][:, 0] = speed * (ts - t0)
    fr["p"][:, 1] = 0.5 * speed * yaw_rate * (ts - t0) ** 2
    fr["p"][:, 2] = 1.8
    fr["v"][:, 0], fr["v"][:, 1] = speed * np.cos(yaw), speed * np.sin(yaw)
    fr["w"] = np.array([0.02, -0.03, yaw_rate], np.float32) + rng.normal(0, 0.01, (n, 3)).astype(np.float32)
    fr["a"] = np.array([0.3, speed * yaw_rate, 9.809], np.float32) + rng.normal(0, 0.05, (n, 3)).astype(np.float32)
    fr["bg"] = np.array([0.001, -0.002, 0.0015], np.float32)
    fr["ba"] = np.array([0.01, 0.02, -0.01], np.float32)
    fr["g"] = np.array([0, 0, -9.809], np.float32)
    return fr


# ---- a motion-distorted scan stream (BASELINE config c3: replay with map growth) ------------------------
class Stream:
    """Vehicle on a circle (radius r, speed v) inside the canyon world; spinning scans whose points are
    taken from the pose AT THEIR OWN FIRING TIME (so deskewing is really needed), ideal IMU frames."""

    def __init__(self, seed=1003, rings=64, azimuths=2048, radius=40.0, speed=10.0, scan_dt=0.1, imu_hz=200.0):
        self.world = make_world(seed, 120.0, 70.0, 25.0, 60)
        self.rings, self.az, self.r, self.v, self.dt, self.imu_hz, self.seed = rings, azimuths, radius, speed, scan_dt, imu_hz, seed
        self.omega = speed / radius

    def pose(self, t):
        """position (…,3), yaw of the trajectory at time(s) t."""
        t = np.asarray(t, np.float64)
        th = self.omega * t
        p = np.stack([self.r * np.sin(th), self.r * (1.0 - np.cos(th)) - self.r, np.full_like(th, 1.8)], -1)
        return p, th

    def state(self, t):
        p, yaw = self.pose(t)
        return make_state(p, quat_from_rpy(0.0, 0.0, float(yaw)), vel=(self.v * np.cos(yaw), self.v * np.sin(yaw), 0.0))

    def scan(self, k):
        """Raw message k (api.RAW_POINT records, VELODYNE timing: `time` = seconds since the sweep start
        stamp k*dt) and its start stamp."""
        from .api import RAW_POINT
        rng = np.random.default_rng(self.seed + 17 * k)
        n = self.rings * self.az
        el = np.deg2rad(np.linspace(2.0, -24.8, self.rings))
        az = np.linspace(0.0, 2 * np.pi, self.az, endpoint=False)
        ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
        d_body = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (self.rings, self.az))], -1).reshape(-1, 3)
        frac = np.tile(np.arange(self.az) / self.az, self.rings) + rng.uniform(0, 0.2 / n, n)
        t_rel = (frac * self.dt).astype(np.float32)
        stamp = k * self.dt
        p, yaw = self.pose(stamp + t_rel.astype(np.float64))
        c, s = np.cos(yaw), np.sin(yaw)
        d_world = np.stack([c * d_body[:, 0] - s * d_body[:, 1], s * d_body[:, 0] + c * d_body[:, 1], d_body[:, 2]], -1)
        rng_t = _raycast(self.world, (p[:, 0], p[:, 1], p[:, 2]), d_world)
        bad = ~np.isfinite(rng_t) | (rng_t > 150.0)
        rng_t = rng_t + rng.normal(0.0, 0.01, n)
        pts = (d_body * rng_t[:, None]).astype(np.float32)
        pts[bad] = np.nan                                           # no return: the NaN filter drops them
        raw = np.zeros(n, RAW_POINT)
        raw["x"], raw["y"], raw["z"], raw["time"] = pts[:, 0], pts[:, 1], pts[:, 2], t_rel
        return raw, stamp

    def imu(self, t0, t1, sigma_acc=0.0, sigma_gyro=0.0, seed=0):
        """Calibrated base-link IMU samples with t0 < stamp <= t1 on the IMU clock (stamp = i / imu_hz):
        (stamps, dt, lin_accel (n,3) float32, ang_vel (n,3) float32) — what Localizer::updateIMU pushes on
        imu_buffer.  Constant speed on the circle: specific force = centripetal (body +y) - gravity."""
        i0, i1 = int(np.floor(t0 * self.imu_hz + 1e-9)) + 1, int(np.floor(t1 * self.imu_hz + 1e-9))
        idx = np.arange(i0, i1 + 1)
        n = len(idx)
        rng = np.random.default_rng(self.seed + 7919 * i0 + seed)
        acc = np.tile(np.float64([0.0, self.v * self.omega, 9.809]), (n, 1)) + rng.normal(0, 1, (n, 3)) * sigma_acc
        gyr = np.tile(np.float64([0.0, 0.0, self.omega]), (n, 1)) + rng.normal(0, 1, (n, 3)) * sigma_gyro
        return idx / self.imu_hz, np.full(n, 1.0 / self.imu_hz), acc.astype(np.float32), gyr.astype(np.float32)

    def frames(self, t0, t1):
        """fast_limo::State records at the IMU rate covering [t0, t1] (ideal IMU, zero biases)."""
        from .api import FRAME
        n = int(np.ceil((t1 - t0) * self.imu_hz)) + 3
        ts = t0 - 1.0 / self.imu_hz + np.arange(n) / self.imu_hz
        p, yaw = self.pose(ts)
        fr = np.zeros(n, FRAME)
        fr["time"] = ts
        fr["q"][:, 2], fr["q"][:, 3] = np.sin(yaw / 2), np.cos(yaw / 2)
        fr["p"] = p
        fr["v"][:, 0], fr["v"][:, 1] = self.v * np.cos(yaw), self.v * np.sin(yaw)
        fr["w"][:, 2] = self.omega
        fr["a"][:, 1] = self.v * self.omega                        # centripetal, body frame (y = left)
        fr["a"][:, 2] = 9.809
        fr["g"][:, 2] = -9.809
        return fr
