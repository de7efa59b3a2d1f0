"""The C++ host side: the PDELab-concept shim (include/ax1gpu_pdelab_shim.hh) compiles against the C ABI
without DUNE (CPU check) and drives one time step on the GPU to the same result as the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import oracle_problem, small_setup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "shim_step.cpp")


def _build(tmp_path):
    exe = str(tmp_path / "shim_step")
    libdir = os.path.join(ROOT, "dune_ax1_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir,
           "-l:libax1gpu.so", "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64",
           "-lcudart"]
    subprocess.check_call(cmd)
    return exe


def test_shim_compiles_without_dune(axlib, tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [1, 2, 3])   # 1: IGO / LS / PDESOLVER seam, 2: TimeLoopShim, 3: MoriStepShim
def test_shim_time_step_matches_oracle(oracle, axlib, tmp_path, mode):
    exe = _build(tmp_path)
    s = small_setup(nx=6)
    slab_w = 4
    blob = str(tmp_path / "problem.bin")
    with open(blob, "wb") as f:
        f.write(struct.pack("4i", s["nx"], s["ny"], slab_w, mode))
        for a, dt in ((s["x"], "f8"), (s["y"], "f8"), (s["cell_subdomain"], "u1"), (s["eps_memb"], "f8"),
                      (s["node_volume"], "f8"), (s["dirichlet"], "u1"), (s["g_leak"], "f8"), (s["g_nav"], "f8"),
                      (s["g_kv"], "f8"), (s["u0"], "f8")):
            f.write(np.ascontiguousarray(a, dtype=dt).tobytes())
    out = str(tmp_path / "u.bin")
    r = subprocess.run([exe, blob, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ug = np.fromfile(out)
    P = oracle_problem(oracle, s)
    # the adaptive time loop's first step is already 1.1 dt0 (acme2_cyl_simulation.hh:334-338 with
    # diagInfo.iterations = 0 at the start); the IGO / PDESOLVER seam is driven with dt0 itself
    dt = 1.0 * 1.1 if mode == 2 else 1.0
    P.set_time(0.0, dt); P.update_state(s["u0"]); P.set_uold(s["u0"])
    if mode == 3:     # the electroneutral (Mori) build: one operator-split step
        uc, mr = P.mori_step(s["u0"], oracle.default_newton_opts(reduction=1e-10, abs_limit=1e-12, lin_maxit=500,
                                                                 ilu_fill=0, slab_w=slab_w))
        assert mr.status == 0
        scale = np.max(np.abs(uc.reshape(-1, 4)), axis=0)
        err = np.max(np.abs(ug - uc).reshape(-1, 4) / scale, axis=0)
        assert np.all(err < 1e-9), (err, r.stdout)
        return
    uc, rc = P.newton(s["u0"], oracle.default_newton_opts(reduction=1e-9, abs_limit=1e-14, min_lin_reduction=1e-10,
                                                        ilu_fill=0, slab_w=slab_w))
    assert rc.status == 0
    scale = np.max(np.abs(uc.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(ug - uc).reshape(-1, 4) / scale, axis=0)
    assert np.all(err[:3] < 1e-8) and err[3] < 1e-7, (err, r.stdout)
