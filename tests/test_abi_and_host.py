"""CPU: the C-ABI library loads and exports what include/flimo.h declares; host-side EKF state
machine (product code) against the oracle; pose constants; no compute call touches a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from fast_limo_b200 import _lib, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(flimo_lib):
    hdr = open(os.path.join(ROOT, "include", "flimo.h")).read()
    declared = sorted(set(re.findall(r"\b(flimo_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(flimo_lib, name), f"{name} declared in flimo.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_version_and_defaults(flimo_lib):
    assert b"sm_100a" in flimo_lib.flimo_version()
    c = _lib.FlimoCfg()
    flimo_lib.flimo_cfg_default(C.byref(c))
    assert (c.NUM_MATCH_POINTS, c.MAX_NUM_PC2MATCH, c.MAX_DIST_PLANE, c.PLANE_THRESHOLD) == (5, 10000, 2.0, 0.05)


def test_create_without_gpu_fails_loudly(flimo_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.FlimoError, match="no CUDA device|no CPU path"):
        api.Mapper(device=0)
    m = api.Mapper(device=-1)                       # host-only handle: EKF algebra only
    with pytest.raises(api.FlimoError, match="host-only"):
        m.add(np.zeros((10, 3), np.float32))
    with pytest.raises(api.FlimoError, match="host-only"):
        m.match(synth.make_state([0, 0, 0], [0, 0, 0, 1]), np.zeros((10, 3), np.float32))


def test_unsupported_k_rejected(flimo_lib):
    with pytest.raises(api.FlimoError, match="NUM_MATCH_POINTS"):
        api.Mapper(api.MappingConfig(NUM_MATCH_POINTS=7), device=-1)


@pytest.mark.parametrize("form", ["12x12", "reference", "cooperative"])
@pytest.mark.parametrize("n_rows,max_iter,limit", [(800, 3, 0.001), (800, 0, 0.001), (800, 4, 1e9), (10, 2, 0.001), (0, 2, 0.001),
                                                   (800, 6, 1e-3), (40, 3, 1e-7)])
def test_host_ekf_matches_oracle(oracle, flimo_lib, n_rows, max_iter, limit, form, monkeypatch):
    # the gain is formed with one 12x12 inverse by default; FLIMO_EKF_REFERENCE_FORM=1 (read when the handle is
    # created) evaluates it with the reference's two 23x23 inversions; FLIMO_EKF_COOP=1 runs the DEVICE code path's
    # phase-structured step (csrc/ekf_step.hpp, one 12x25 elimination per pass) item by item on the host — all three
    # must agree with the oracle
    monkeypatch.delenv("FLIMO_EKF_REFERENCE_FORM", raising=False)
    monkeypatch.delenv("FLIMO_EKF_COOP", raising=False)
    if form == "reference":
        monkeypatch.setenv("FLIMO_EKF_REFERENCE_FORM", "1")
    if form == "cooperative":
        monkeypatch.setenv("FLIMO_EKF_COOP", "1")
    rng = np.random.default_rng(n_rows + max_iter)
    from scipy.spatial.transform import Rotation as Rot
    x0 = synth.make_state(rng.normal(0, 3, 3), Rot.random(random_state=1).as_quat(),
                          Rot.from_rotvec([0.01, -0.02, 0.03]).as_quat(), [0.1, 0, 0.05], [1, 0.5, 0], [1e-3] * 3, [0.02] * 3,
                          [0.2, -0.1, -np.sqrt(9.809 ** 2 - 0.05)])
    P0 = synth.default_P0()
    H = rng.normal(size=(n_rows, 12))
    H[:, 6:] *= 0.2
    h = rng.normal(0, 0.03, n_rows)
    xo, Po, tr = oracle.update_fixed(x0, P0, max_iter, limit, H, h)
    m = api.Mapper(device=-1)
    m.ekf_begin(x0, P0, max_iter, limit)
    HTH, HTh = H.T @ H, H.T @ h
    passes = 0
    done = False
    while not done:
        cur = m.ekf_state()
        if passes:
            assert np.allclose(cur, tr[passes - 1]["state"], atol=1e-12)
        done = m.ekf_step(HTH, HTh, n_rows)
        passes += 1
    x, P = m.ekf_end()
    assert passes == len(tr)
    assert np.allclose(x, xo, atol=1e-11)
    assert np.allclose(P, Po, rtol=1e-8, atol=1e-11)


def test_unpack96_layout(flimo_lib):
    p = np.arange(96, dtype=np.float64)
    r = api.unpack96(p)
    assert r.HTH[0, 0] == 0 and r.HTH[0, 11] == 11 and r.HTH[1, 1] == 12 and r.HTH[11, 11] == 77
    assert np.array_equal(r.HTH, r.HTH.T)
    assert np.array_equal(r.HTh, np.arange(78, 90))
    assert (r.n_rows, r.sum_sq_res, r.n_valid) == (90, 91.0, 92)


def _prototypes():
    """name -> number of parameters, parsed from include/flimo.h (comments stripped)."""
    hdr = open(os.path.join(ROOT, "include", "flimo.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(flimo_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_ctypes_declarations_match_header(flimo_lib):
    """Every prototype of flimo.h has ctypes argtypes of the same arity (an ABI drift between the header and
    fast_limo_b200/_lib.py would otherwise only show up as stack garbage on the GPU box)."""
    protos = _prototypes()
    assert set(protos) == set(_lib.SYMBOLS)
    for name, n_args in protos.items():
        fn = getattr(flimo_lib, name)
        if fn.argtypes is None:
            assert n_args == 0, f"{name}: {n_args} parameters in flimo.h, no argtypes in _lib.py"
        else:
            assert len(fn.argtypes) == n_args, f"{name}: {n_args} parameters in flimo.h, {len(fn.argtypes)} in _lib.py"


def test_struct_layouts_match_header(flimo_lib, tmp_path):
    """sizeof of every struct of flimo.h as gcc lays it out == the ctypes / numpy mirrors."""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "flimo.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(flimo_cfg), '
                   'sizeof(flimo_prep_cfg), sizeof(flimo_frame), sizeof(flimo_msg_layout), sizeof(flimo_imu), sizeof(flimo_stats)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    mirrors = [C.sizeof(_lib.FlimoCfg), C.sizeof(_lib.FlimoPrepCfg), api.FRAME.itemsize, C.sizeof(_lib.FlimoMsgLayout),
               C.sizeof(_lib.FlimoImu), C.sizeof(_lib.FlimoStats)]
    assert sizes == mirrors
