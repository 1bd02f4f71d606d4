"""CPU: the C++ façade (include/fast_limo_gpu/Mapper.hpp) compiles with plain g++ against
include/flimo.h, links libflimo_cuda.so and runs the host-side entry points."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_compiles_and_runs(flimo_lib, tmp_path):
    exe = tmp_path / "facade_smoke"
    libdir = os.path.join(ROOT, "fast_limo_b200")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "facade_smoke.cpp"), "-o", str(exe),
           "-L", libdir, "-l:libflimo_cuda.so", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "facade ok" in r.stdout
