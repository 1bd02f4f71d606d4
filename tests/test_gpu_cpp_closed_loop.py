"""GPU (-m gpu): the reference's class surface in C++ (include/fast_limo/**, libfast_limo.so) in closed loop.

tests/cpp/closed_loop.cpp drives fast_limo::Localizer::getInstance() the way the ROS wrapper does (init(config), updateIMU per IMU
message, updatePointCloud per LiDAR message) on a synthetic stream written to disk here; the host mirror in Python
(fast_limo_b200/localizer.py) is fed the same messages.  Both sit on the same C ABI, so their poses must agree to round-off;
and the estimate must track the ground truth."""
import os
import struct
import subprocess

import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIG = 1 << 18
N_SCANS = 60


@pytest.mark.timeout(900)
def test_cpp_localizer_closed_loop_matches_python_mirror(flimo_lib, tmp_path):
    from fast_limo_b200.localizer import Localizer, LocalizerConfig
    S = synth.Stream(rings=64, azimuths=512, imu_hz=200.0)
    x0 = S.state(0.0)
    m = api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    filt = api.FilterConfig(cropBoxMin=(-1, -1, -1), cropBoxMax=(1, 1, 1), min_dist=3.0, leafSize=0.5, sensor_type=1)
    loc = Localizer(m, LocalizerConfig(filters=filt, MAX_NUM_ITERS=3), pos=x0[0:3], quat=x0[3:7], vel=x0[14:17])
    stream = tmp_path / "stream.bin"
    py_states, errs = [], []
    with open(stream, "wb") as f:
        f.write(struct.pack("<Qddii", N_SCANS, 0.5, 3.0, 3, 1))
        f.write(np.float64(x0[0:3]).tobytes() + np.float64(x0[3:7]).tobytes() + np.float64(x0[14:17]).tobytes())
        t_imu = 0.0
        for k in range(N_SCANS):
            raw, stamp = S.scan(k)
            t_need = stamp + S.dt + 1.0 / S.imu_hz
            stamps, dts, acc, gyr = S.imu(t_imu, t_need, sigma_acc=0.05, sigma_gyro=0.002)
            acc = acc + np.float32([0.3, 0.0, 0.0])                      # an uncalibrated accelerometer bias
            t_imu = t_need
            f.write(struct.pack("<dQ", stamp, len(stamps)))
            for t, a, w in zip(stamps, acc, gyr):
                f.write(struct.pack("<d3f3f", t, *a, *w))
                loc.updateIMU_raw(t, a, w)
            f.write(struct.pack("<Q", len(raw)))
            f.write(raw.tobytes())
            loc.updatePointCloud(raw, stamp)
            py_states.append(np.concatenate([loc.x, np.diag(loc.P), [m.size(), loc.last.get("passes", 0)]]))
            errs.append(float(np.linalg.norm(loc.x[0:3] - S.state(loc.imu_stamp)[0:3])))
    m.close()
    libdir = os.path.join(ROOT, "fast_limo_b200")
    exe = tmp_path / "closed_loop"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "closed_loop.cpp"),
                    "-o", str(exe), "-L", libdir, "-lfast_limo", "-lflimo_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    poses = tmp_path / "poses.bin"
    r = subprocess.run([str(exe), str(stream), str(poses)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    cpp = np.fromfile(poses, np.float64).reshape(N_SCANS, 51)
    py = np.array(py_states)
    assert np.array_equal(cpp[:, 49], py[:, 49])                         # map sizes after every scan
    assert np.array_equal(cpp[2:, 50], py[2:, 50])                       # passes of every registered scan
    assert np.abs(cpp[:, :26] - py[:, :26]).max() <= 1e-9                # same C ABI underneath: states agree to round-off
    assert np.allclose(cpp[:, 26:49], py[:, 26:49], rtol=1e-6, atol=1e-12)
    assert cpp[-1, 49] > 100000 and max(errs[2:]) < 0.06, (cpp[-1, 49], max(errs[2:]))
