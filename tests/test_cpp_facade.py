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


def _gxx(args, **kw):
    return subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include")] + args, capture_output=True, text=True, **kw)


def test_reference_call_surface_type_checks(flimo_lib):
    """fast_limo::Localizer / fast_limo::Mapper with the reference's signatures (include/fast_limo/**): the ROS wrapper's
    call sites (src/main.cpp:16-93,178-206), the marker code's member accesses and the IKFoM hook (use-ikfom.cpp:10-31)
    compile against the headers without ROS / PCL / Eigen."""
    r = _gxx(["-fsyntax-only", os.path.join(ROOT, "tests", "cpp", "wrapper_callsites.cpp")])
    assert r.returncode == 0, r.stderr


def test_closed_loop_harness_links(flimo_lib, tmp_path):
    """The C++ closed-loop harness builds against libfast_limo.so (the GPU run is tests/test_gpu_cpp_closed_loop.py); without
    a GPU Localizer::init must fail loudly (no CPU path behind the reference's class surface either)."""
    libdir = os.path.join(ROOT, "fast_limo_b200")
    assert os.path.exists(os.path.join(libdir, "libfast_limo.so"))
    exe = tmp_path / "closed_loop"
    r = _gxx(["-O1", os.path.join(ROOT, "tests", "cpp", "closed_loop.cpp"), "-o", str(exe), "-L", libdir, "-lfast_limo", "-lflimo_cuda",
              f"-Wl,-rpath,{libdir}"])
    assert r.returncode == 0, r.stderr
    import struct
    import torch
    if torch.cuda.is_available():
        return
    stream = tmp_path / "s.bin"
    stream.write_bytes(struct.pack("<Qddii3d4d3d", 0, 0.5, 3.0, 3, 1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0))
    run = subprocess.run([str(exe), str(stream), str(tmp_path / "p.bin")], capture_output=True, text=True, timeout=60)
    assert run.returncode != 0 and "no CUDA device" in (run.stderr + run.stdout)
