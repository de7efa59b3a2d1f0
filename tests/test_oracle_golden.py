"""Pin the CPU oracle (and the product's host model) against the golden vectors the reference ships:
the converged equilibrium states src/simulation_states/equilibrium/cyl/yasp_square{10,1}mm_p0.dat
(extracted to tests/golden/ by tests/golden/make_golden.py). The reference's own tests hold no
assertion on the hot path, so these are the known-answer tests available.
CPU only."""
import json
import os
import re

import numpy as np
import pytest

from helpers import oracle_problem, small_setup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G10 = json.load(open(os.path.join(ROOT, "tests", "golden", "yasp_square10mm.json")))
G1 = json.load(open(os.path.join(ROOT, "tests", "golden", "yasp_square1mm.json")))


def _equilibrium_problem(oracle, g, ymax):
    cfg = oracle.default_config(xmin_global=0.0, xmax_global=g["x"][-1], ymin=0.0, ymax=ymax, do_equilibration=1,
                                t_equilibrium=1.0e9, dummy_value=10.0)
    P = oracle.Problem(cfg, np.array(g["x"]), np.array(g["y"]), g_leak=0.5, g_nav=120.0, g_kv=36.0)
    u = np.array(g["unew"])
    P.set_time(g["time"], g["dt"])
    P.set_uold(u)
    P.update_state(u)
    return P, u


# ---- KAT-1: radial grid vector, bit exact (oracle and product host model) ----------------------
@pytest.mark.parametrize("g,ymax", [(G10, 10.0e6), (G1, 1.0e6)])
def test_grid_vector_bit_exact(oracle, axlib, g, ymax):
    want = np.array(g["y"])
    got_o = oracle.generate_y(ymax=ymax)
    got_h = axlib.generate_y(ymax=ymax)
    assert len(got_o) == len(want) and np.array_equal(got_o, want)
    assert len(got_h) == len(want) and np.array_equal(got_h, want)


def test_dof_layout_and_constants(oracle):
    # 2 x 181 vertices x [Na,K,Cl,pot], x fastest: consecutive 4-blocks are the two x-vertices of one y
    u = np.array(G10["unew"]).reshape(181, 2, 4)
    assert np.allclose(u[:, 0, :], u[:, 1, :], rtol=1e-8, atol=1e-7)
    # bulk values: cytosol Na 12 / K 125 / Cl 137, ES 100 / 4 / 104 (square10mm.config)
    assert np.allclose(u[0, 0, :3], [12.0, 125.0, 137.0], rtol=5e-3)
    assert np.allclose(u[-1, 0, :3], [100.0, 4.0, 104.0], rtol=1e-6)
    assert abs(u[0, 0, 3] - (-2.71830)) < 1e-4          # phi_in = -65.46 mV / (kT/e)
    P, _ = _equilibrium_problem(oracle, G10, 10.0e6)
    c = P.constants()
    assert abs(c["PC"] - 0.4525173207) < 1e-9           # electrolyte.hh:66
    assert np.allclose(c["D"], [1330.0, 1960.0, 2030.0], rtol=1e-14)
    assert abs(c["kT_e"] * 1e3 - 24.08115745) < 1e-7    # mV


# ---- channel states: leak ratios + HH steady state m,h,n at v - v_rest = 0 ----------------------
@pytest.mark.parametrize("g", [G10, G1])
def test_channel_steady_state(oracle, g):
    s = small_setup(nx=1)
    P = oracle_problem(oracle, s)
    got = P.channel_state()
    want = np.array(g["channelStates"])
    assert np.max(np.abs(got - want) / want) < 5e-16


# ---- KAT-2/3: the shipped state is a steady state of the oracle's residual ----------------------
def test_equilibrium_poisson_rows_cancel(oracle):
    P, u = _equilibrium_problem(oracle, G10, 10.0e6)
    r, a = P.residual(u, with_abs=True)
    rel = (np.abs(r) / (a + 1e-300)).reshape(181, 2, 4)
    # potential rows through both Debye layers and across the membrane rows (j = 30, 31)
    assert rel[18:45, :, 3].max() < 2e-9
    assert rel[30:32, :, 3].max() < 1e-10
    vm = P.membrane_potential(u)
    assert abs(vm[0] - (-64.9076)) < 1e-3


def test_equilibrium_is_steady_under_one_step(oracle):
    """One implicit-Euler step (dt = 10 us, equilibration mode: leak only, dummy_value = 10) from the
    shipped converged state leaves it unchanged to ~1e-6 relative: the strongest end-to-end check
    of residual + Jacobian + flux + solver available without the DUNE binary."""
    for g, ymax in ((G10, 10.0e6), (G1, 1.0e6)):
        for analytic in (7, 0):
            P, u = _equilibrium_problem(oracle, g, ymax)
            P.cfg.analytic_volume_jacobian = analytic
            un, res = P.newton(u, oracle.default_newton_opts())   # reference tolerances 1e-5 / 1e-5
            assert res.status == 0 and res.iterations <= 3
            scale = np.max(np.abs(u.reshape(-1, 4)), axis=0)
            change = np.max(np.abs(un - u).reshape(-1, 4) / scale, axis=0)
            assert np.all(change < 1e-5), change


def test_fd_and_analytic_jacobian_agree(oracle):
    s = small_setup(nx=3)
    Pa = oracle_problem(oracle, s, analytic=1)
    Pf = oracle_problem(oracle, s, analytic=0)
    for P in (Pa, Pf):
        P.set_time(0.0, 1.0); P.set_uold(s["u0"]); P.update_state(s["u0"])
    Ja = Pa.jacobian(s["u0"]).reshape(-1, 4, 4)
    Jf = Pf.jacobian(s["u0"]).reshape(-1, 4, 4)
    rowptr, _ = Pa.pattern()
    rows = np.repeat(np.arange(Pa.nv), np.diff(rowptr))
    scale = np.zeros((Pa.nv, 4))
    np.maximum.at(scale, rows, np.max(np.abs(Ja), axis=2))
    assert np.max(np.abs(Ja - Jf) / (scale[rows][:, :, None] + 1e-300)) < 1e-5


def test_jacobian_is_derivative_of_residual(oracle):
    """Directional derivative check: (R(u + h d) - R(u - h d)) / 2h = J d, away from the membrane rows
    (whose opposite-side coupling is deliberately not in the Jacobian: quasi-Newton)."""
    s = small_setup(nx=3)
    P = oracle_problem(oracle, s)
    u = s["u0"]
    P.set_time(0.0, 1.0); P.set_uold(u); P.update_state(u)
    J = P.jacobian(u)
    rng = np.random.default_rng(1)
    d = rng.standard_normal(P.n) * np.abs(u).clip(1e-3)
    h = 1e-6
    fd = (P.residual(u + h * d) - P.residual(u - h * d)) / (2 * h)
    jd = P.spmv(J, d)
    nvx = s["nx"] + 1
    keep = np.ones(P.nv, dtype=bool)
    keep[(P.jm) * nvx:(P.jm + 2) * nvx] = False
    _, _, dm = P.maps()
    keep &= dm == 0
    sc = P.spmv(np.abs(J), np.abs(d)).reshape(-1, 4)
    err = np.abs(fd - jd).reshape(-1, 4)[keep] / (sc[keep] + 1e-300)
    assert err.max() < 1e-6, err.max()


def test_bicgstab_ilu_true_residual(oracle):
    s = small_setup(nx=12)
    P = oracle_problem(oracle, s)
    u = s["u0"]
    P.set_time(0.0, 1.0); P.set_uold(u); P.update_state(u)
    J = P.jacobian(u)
    r = P.residual(u)
    for fill, sw in ((1, 0), (0, 0), (0, 4)):
        z, rr, res = P.solve(J, r, 1e-8, maxit=500, ilu_fill=fill, slab_w=sw)
        assert res.converged == 1
        assert np.linalg.norm(r - P.spmv(J, z)) <= 2e-8 * np.linalg.norm(r)
    # ILU(1) on the x-fastest 9-point stencil: 13 blocks per interior row
    import ctypes as C
    nb1 = oracle.lib().ax1o_ilu_num_blocks(P.h, 1, 0)
    nb0 = oracle.lib().ax1o_ilu_num_blocks(P.h, 0, 0)
    assert nb0 == (3 * 12 + 1) * (3 * s["ny"] + 1)
    nvx, nvy = 13, s["ny"] + 1
    interior_extra = 4 * (nvx - 4) * (nvy - 2)   # +-2 and +-(nx-2..) fill blocks of fully interior rows
    assert nb1 >= nb0 + interior_extra


# ---- host model of the product vs oracle (bit exact maps), multi-rank local grids ----------------
@pytest.mark.parametrize("nranks,rank", [(1, 0), (3, 0), (3, 1), (3, 2)])
def test_host_model_matches_oracle(oracle, axlib, nranks, rank):
    from dune_ax1_b200 import setups
    bc = axlib.BoundaryConfig(left_cytosol=(True, False), right_extra=(False, True), bottom=(False, False))
    s = setups.make_setup(9, dx=100.0e3, rank=rank, nranks=nranks, bc=bc)
    P = oracle_problem(oracle, s)
    cs, nvol, dm = P.maps()
    assert np.array_equal(cs, s["cell_subdomain"])
    assert np.array_equal(nvol, s["node_volume"])
    assert np.array_equal(dm, s["dirichlet"])
    assert nvol.max() > 1e6 and nvol.min() == 1.0  # volume scaling active far from the membrane


# ---- the C ABI library loads and exports every declared symbol; no GPU -> loud failure ---------
def test_abi_symbols_exported(axlib):
    hdr = open(os.path.join(ROOT, "include", "ax1gpu.h")).read()
    names = set(re.findall(r"\b(ax1(?:gpu|host)_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) > 35
    L = axlib.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback(axlib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dune_ax1_b200 import setups
    s = small_setup(nx=2)
    with pytest.raises(axlib.Ax1GpuError) as e:
        setups.make_device(s)
    assert "no CUDA device" in str(e.value)


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "dune_ax1_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "ax1o_" not in txt and "libax1oracle" not in txt, f
