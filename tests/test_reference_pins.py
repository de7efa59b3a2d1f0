"""Pins of the oracle against the REFERENCE'S OWN COMPILED CODE (oracle/_ref/libax1ref.so, built by
`make -C oracle _ref` from /root/reference/dune/ax1/{common/ax1_gridvector.hh, channels/channel.hh,
channels/channelset.hh} against the stand-in headers of oracle/ref_stubs): the GridVector primitives behind the
grid generator, and the Hodgkin-Huxley channel set (steady-state initialisation, leak re-balancing, backward-Euler
gating step, effective conductances) -- SURVEY row a10. The product's host model is held to the same numbers.
CPU only; skipped where the reference tree is absent and no prebuilt library travelled along."""
import ctypes as C

import numpy as np
import pytest

from helpers import oracle_problem, small_setup


@pytest.fixture(scope="module")
def ref(oracle):
    oracle.build_ref()
    R = oracle.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return R


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


GRID_CASES = [
    (0, 0.0, 17, 0.5, 0.0, 0.0),                 # equidistant_n_h
    (0, 494.5, 10, 0.5, 0.0, 0.0),
    (1, 505.0, 7, 1.0e4, 0.0, 0.0),              # equidistant_n_xend
    (1, 0.0, 100, 1.0e7, 0.0, 0.0),              # the paper's x vector
    (6, 0.0, 12, 100.0, 495.0, 0.0),             # geometric_n_h0_xend (Newton iteration for q)
    (7, 510.0, 15, 1.0e5, 3.0e5, 0.0),           # geometric_n_hend_xend
    (8, 0.0, 0, 100.0, 0.5, 495.0),              # geometric_h0_hend_xend, h0 guaranteed (towards the membrane)
    (9, 510.0, 0, 0.5, 1.0e5, 4.0e5),            # hend guaranteed (away from the membrane)
    (8, 300.0, 0, 10.0, 0.5, 495.0),
]


@pytest.mark.parametrize("case", GRID_CASES)
def test_gridvector_primitives_bit_exact(oracle, ref, case):
    op, start, n, a, b, c = case
    out_r, out_o = np.zeros(4096), np.zeros(4096)
    nr = ref.ax1ref_gridvector(op, start, n, a, b, c, _dp(out_r), 4096)
    no = oracle.lib().ax1o_gridvector(op, start, n, a, b, c, _dp(out_o), 4096)
    assert nr > 1 and nr == no
    assert np.array_equal(out_r[:nr], out_o[:no])


@pytest.mark.parametrize("gbar,v_init", [((0.5, 120.0, 36.0), -65.0), ((0.5, 1200.0, 360.0), -64.9076),
                                         ((0.3, 40.0, 90.0), -70.0)])
def test_channel_set_against_reference_classes(oracle, ref, gbar, v_init):
    """DynamicChannelSet with [Na_leak, K_leak, Nav, Kv]: initChannels at v_init, then backward-Euler gating steps
    through a depolarisation. The oracle is driven through its public API on a one-column problem whose membrane
    potential is set to the same values."""
    g_leak, g_nav, g_kv = gbar
    dt_us = 10.0
    v_steps = np.array([-60.0, -40.0, -10.0, 20.0, 35.0, 10.0, -30.0, -70.0, -75.0, v_init])
    out = np.zeros(9 * (len(v_steps) + 1))
    rc = ref.ax1ref_channelset(g_leak, g_nav, g_kv, 0.13, 0.87, v_init, dt_us * 1e-6, _dp(v_steps), len(v_steps), _dp(out))
    assert rc == 0
    want = out.reshape(-1, 9)

    s = small_setup(nx=1, perturbation=0.0)
    s["g_leak"][:], s["g_nav"][:], s["g_kv"][:] = g_leak, g_nav, g_kv
    P = oracle_problem(oracle, s)
    kT_e_mV = P.constants()["kT_e"] * 1e3
    nvx, jm = s["nx"] + 1, P.jm

    def state_with_vm(v_mV):
        u = s["u0"].reshape(-1, nvx, 4).copy()
        u[jm, :, 3] = v_mV / kT_e_mV          # cytosol-side membrane vertices
        u[jm + 1, :, 3] = 0.0                 # extracellular side
        assert abs(P.membrane_potential(u.reshape(-1))[0] - v_mV) < 1e-12 * max(1.0, abs(v_mV))
        return u.reshape(-1)

    def check(stage, got5):
        w = want[stage]
        assert np.max(np.abs(got5 - w[:5]) / np.abs(w[:5])) < 1e-14, (stage, got5, w[:5])
        # effective conductances as the flux kernel forms them from the same state
        geff = np.array([g_leak * got5[0], g_leak * got5[1], g_nav * got5[2] ** 3 * got5[3], g_kv * got5[4] ** 4])
        assert np.max(np.abs(geff - w[5:]) / (np.abs(w[5:]) + 1e-300)) < 1e-13

    t = 0.0
    P.set_time(t, dt_us)
    P.update_state(state_with_vm(v_init))      # initChannels (steady state, v_rest := v_init, leak re-balancing)
    assert abs(P.v_rest() - v_init) < 1e-12
    check(0, P.channel_state(new=True))
    P.accept_state()
    for k, v in enumerate(v_steps):
        P.set_time(t, dt_us)
        P.update_state(state_with_vm(v))       # backward-Euler step from the accepted state
        check(k + 1, P.channel_state(new=True))
        P.accept_state()
        t += dt_us
    # the shipped channelStates vector (m, h, n at rest) is the reference's own steady state too
    if v_init == -65.0:
        assert abs(want[0][2] - 0.052932485257249577) < 1e-15 and abs(want[0][4] - 0.31767691406069742) < 1e-15


def test_poisson_constant_against_reference_electrolyte(oracle, ref):
    """Electrolyte(eps = 80, T = 279.45 K, stdCon = N_A, lengthScale = 1e-9) as default_config.hh builds it."""
    out = np.zeros(2)
    ref.ax1ref_electrolyte(80.0, 279.45, 6.02214129e23, 1e-9, _dp(out))
    P = oracle_problem(oracle, small_setup(nx=1))
    assert P.constants()["PC"] == out[0]
    assert abs(out[0] - 0.4525173207) < 1e-9


def test_node_volumes_against_reference_cylinder_geometry(oracle, axlib, ref):
    """Acme2CylGeometry::volume() = pi (r_last + r_first) * vol2D of the reference (acme2_cyl_geometrywrapper.hh:64-72)
    is what Physics::calcNodeVolumes divides: node volume = min over the adjacent cells of V_cell / V_reference, with
    V_reference the first cell above the scaling threshold. Checked for the oracle's and the product host model's maps."""
    s = small_setup(nx=4)
    P = oracle_problem(oracle, s)
    _, nvol_o, _ = P.maps()
    x, y = s["x"], s["y"]
    nx, ny = s["nx"], s["ny"]

    def vol(i, j):
        out = np.zeros(2)
        ref.ax1ref_cyl_geometry(2, x[i], y[j], x[i + 1], y[j + 1], 0.5, 0.5, _dp(out))
        assert out[0] == 2 * np.pi * (y[j] + 0.5 * (y[j + 1] - y[j])) * ((x[i + 1] - x[i]) * (y[j + 1] - y[j])) or \
            abs(out[0] / (2 * np.pi * 0.5 * (y[j] + y[j + 1]) * (x[i + 1] - x[i]) * (y[j + 1] - y[j])) - 1) < 1e-15
        return out[1]

    y_start = 500.0 + 5.0 + 500.0                     # y_memb + d_memb + cellWidth (physics.hh:2286-2397)
    j0 = int(np.nonzero(0.5 * (y[:-1] + y[1:]) > y_start)[0][0])
    v_ref = vol(0, j0)
    nvx = nx + 1
    for (i, j) in ((2, ny), (0, ny), (nx, ny - 3), (1, j0 + 5), (3, j0 + 1)):
        cells = [(ci, cj) for ci in (i - 1, i) for cj in (j - 1, j) if 0 <= ci < nx and 0 <= cj < ny
                 and 0.5 * (y[cj] + y[cj + 1]) > y_start]
        want = min(vol(ci, cj) / v_ref for (ci, cj) in cells)
        assert nvol_o[i + nvx * j] == want, (i, j)
        assert s["node_volume"][i + nvx * j] == want, (i, j)     # product host model (ax1host_physics_maps)
    assert np.all(nvol_o[: nvx * j0] == 1.0)
    # a membrane-interface face: the wrapped edge volume is the cylinder surface 2 pi r L
    out = np.zeros(2)
    ref.ax1ref_cyl_geometry(1, x[1], 505.0, x[2], 505.0, 0.5, 0.0, _dp(out))
    assert out[1] == (np.pi * (505.0 + 505.0)) * (x[2] - x[1])


# ---- the reference's own local operators and parameter classes on single cells -------------------------------------
@pytest.fixture(scope="module")
def refop(oracle):
    oracle.build_ref()
    R = oracle.ref_operator_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return R


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class _CellRunner:
    """Feeds one cell of an oracle problem to the reference operators (oracle/ref_operator_driver.cpp)."""

    def __init__(self, oracle, refop, s, t, dt, stim):
        self.P = oracle_problem(oracle, s)
        self.P.set_time(t, dt); self.P.set_uold(s["u0"]); self.P.update_state(s["u0"])
        self.R, self.s = refop, s
        self.cst = self.P.constants()
        _, self.nvol, _ = self.P.maps()
        self.nvx = s["nx"] + 1
        self.stim = np.array([1.0 if stim else 0.0, t + dt, s["params"]["tinj_start"], s["params"]["tinj_end"],
                              s["params"]["stim_x"], s["params"]["stim_y"], s["params"]["I_inj"]])

    def verts(self, i, j):
        n = self.nvx
        return [i + n * j, i + 1 + n * j, i + n * (j + 1), i + 1 + n * (j + 1)]

    def local(self, i, j, rng):
        u0 = self.s["u0"]
        xl = np.array([[u0[4 * v + c] for v in self.verts(i, j)] for c in range(4)]).reshape(-1)
        return xl * (1.0 + 1e-2 * rng.standard_normal(16))

    def ref(self, which, jac, i, j, xl, eps):
        x, y = self.s["x"], self.s["y"]
        cell = np.array([x[i], x[i + 1], y[j], y[j + 1]])
        D = self.cst["D"]
        phys = np.array([eps, D[0], D[1], D[2], self.cst["PC"], 1.0, 1e-6])
        val = np.array([1, 1, -1], dtype=np.int32)
        nv4 = np.array([self.nvol[v] for v in self.verts(i, j)])
        out = np.zeros(256 if jac else 16)
        rc = self.R.ax1ref_local_operator(which, jac, _dp(cell), _dp(np.ascontiguousarray(xl)), _dp(phys), _ip(val),
                                          _dp(nv4), 1, _dp(self.stim), _dp(out))
        assert rc == 0
        return out


def test_volume_operators_against_reference_local_operators(oracle, refop):
    """Acme2CylOperatorFullyCoupled::alpha_volume / jacobian_volume with the reference's DefaultPoissonParameters and
    DefaultNernstPlanckParameters (A, b, c, f incl. the Na injection source), NernstPlanckTimeLocalOperator and
    ConvectionDiffusionFEM (membrane Poisson), all compiled from /root/reference, evaluated cell by cell on perturbed
    local coefficients of the shipped equilibrium state: rows a1, a3, a5, a6, a7, a8, a12 and the node-volume scaling
    of a13. The oracle's kernels agree to rounding (1e-14 of the largest entry of the same scalar row)."""
    s = small_setup(nx=4, stimulation=True)
    run = _CellRunner(oracle, refop, s, t=29.0, dt=1.0, stim=True)     # LOP time 30 us: injection is on
    P, jm = run.P, run.P.jm
    rng = np.random.default_rng(1)
    cells = [(1, 0), (0, 3), (2, jm - 1), (3, jm + 1), (1, jm + 9), (0, 100), (3, s["ny"] - 1), (2, jm - 12)]
    assert s["x"][1] <= s["params"]["stim_x"] <= s["x"][2]              # cell (1, 0) holds the injection point
    for (i, j) in cells:
        xl = run.local(i, j, rng)
        for which in (0, 1):
            a, b = run.ref(which, 0, i, j, xl, 80.0), P.cell_kernel(i, j, which, xl, weight=1.0)
            scale = np.abs(b).reshape(4, 4).max(axis=1, keepdims=True) + 1e-300
            assert np.max(np.abs(a - b).reshape(4, 4) / scale) < 1e-14, (i, j, which)
            A, B = run.ref(which, 1, i, j, xl, 80.0).reshape(16, 16), P.cell_jacobian(i, j, which, xl, weight=1.0)
            scale = np.abs(B).max(axis=1, keepdims=True) + 1e-300
            assert np.max(np.abs(A - B) / scale) < 1e-14, (i, j, which)
            assert np.abs(B).max() > 0
    # the injection source really is in the compared numbers
    xl = run.local(1, 0, rng)
    run_off = _CellRunner(oracle, refop, small_setup(nx=4, stimulation=False), t=29.0, dt=1.0, stim=False)
    assert np.max(np.abs(run.ref(0, 0, 1, 0, xl, 80.0) - run_off.ref(0, 0, 1, 0, xl, 80.0))[:4]) > 0
    # membrane cell: Poisson with the membrane permittivity
    for i in range(s["nx"]):
        xl = run.local(i, jm, rng)
        a, b = run.ref(2, 0, i, jm, xl, 2.0), P.cell_kernel(i, jm, 2, xl, weight=1.0)
        assert np.max(np.abs(a - b)) < 1e-14 * np.max(np.abs(b)) and np.all(a[:12] == 0.0)
        A, B = run.ref(2, 1, i, jm, xl, 2.0).reshape(16, 16), P.cell_jacobian(i, jm, 2, xl, weight=1.0)
        assert np.max(np.abs(A - B)) < 1e-14 * np.max(np.abs(B))


@pytest.mark.parametrize("variant", ["hh", "equilibration", "flux_stimulation"])
def test_membrane_boundary_term_against_reference_operator_and_flux_class(oracle, refop, variant):
    """Acme2CylOperatorFullyCoupled::alpha_boundary on membrane-interface faces, with the reference's
    DefaultNernstPlanckParameters::bctype / j (surface scaling A_ext / A_this, cosine-ramped flux stimulation) and
    MembraneFluxGridFunctionImplicit::evaluate over its DynamicChannelSet (reversal potentials, drift and diffusion
    terms per channel, equilibration: leak only x dummy_value) -- rows a2, a7, a9. The Physics look-ups that need the
    grid's interface maps (potential jump, concentration ratio, membrane surface) are answered from the opposite-side
    face-centre values the test computes from the same iterate. The oracle's boundary kernel agrees to rounding."""
    kw = dict(equilibration=(variant == "equilibration"))
    s = small_setup(nx=4, **kw)
    if variant == "flux_stimulation":
        s["params"].update(do_memb_flux_stimulation=1, memb_flux_ion=0, memb_flux_inj=3.0e5, tinj_start=20.0,
                           tinj_end=2000.0, stim_x=250.0e3)
    P = oracle_problem(oracle, s)
    u0, t, dt = s["u0"], 399.0, 1.0
    P.set_time(t, dt); P.set_uold(u0); P.update_state(u0); P.accept_state()
    P.update_state(u0 * (1.0 + 1e-3))                 # a gating step away from the steady state
    cst = P.constants()
    _, nvol, _ = P.maps()
    nvx, x, y, jm = s["nx"] + 1, s["x"], s["y"], P.jm
    rng = np.random.default_rng(3)
    u = u0 * (1.0 + 1e-2 * rng.standard_normal(u0.shape))
    st = P.channel_state(new=True).reshape(5, -1)
    p = s["params"]
    changed = 0
    for (i, from_cyt) in ((0, 1), (2, 1), (2, 0), (3, 0), (1, 1), (1, 0)):
        j = jm - 1 if from_cyt else jm + 1
        vs = [i + nvx * j, i + 1 + nvx * j, i + nvx * (j + 1), i + 1 + nvx * (j + 1)]
        xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
        jopp = jm + 1 if from_cyt else jm
        ua, ub = u[4 * (i + nvx * jopp):][:4], u[4 * (i + 1 + nvx * jopp):][:4]
        opp = 0.5 * ua + 0.5 * ub                      # Q1 value at the centre of the opposite interface face
        memb = np.array([y[jopp], opp[3], opp[0], opp[1], opp[2], (np.pi * (y[jm + 1] + y[jm + 1])) * (x[i + 1] - x[i])])
        gl, gn, gk = s["g_leak"][i], s["g_nav"][i], s["g_kv"][i]
        geff = np.array([gl * st[0, i], gl * st[1, i], gn * st[2, i] ** 3 * st[3, i], gk * st[4, i] ** 4])
        equil = np.array([p["do_equilibration"], t, p["t_equilibrium"], p["dummy_value"]], dtype=np.float64)
        fst = np.array([p.get("do_memb_flux_stimulation", 0), p.get("memb_flux_ion", 2), p.get("memb_flux_inj", 0.0)],
                       dtype=np.float64)
        stim = np.array([0, t + dt, p["tinj_start"], p["tinj_end"], p["stim_x"], p["stim_y"], p["I_inj"]], dtype=np.float64)
        cell = np.array([x[i], x[i + 1], y[j], y[j + 1]])
        phys = np.array([80.0, cst["D"][0], cst["D"][1], cst["D"][2], cst["PC"], 1.0, 1e-6])
        val = np.array([1, 1, -1], dtype=np.int32)
        nv4 = np.array([nvol[v] for v in vs])
        out = np.zeros(16)
        rc = refop.ax1ref_elec_alpha_boundary(_dp(cell), from_cyt, _dp(np.ascontiguousarray(xl)), _dp(phys), _ip(val),
                                              _dp(nv4), 1, _dp(stim), _dp(memb), _dp(geff), _dp(equil), _dp(fst),
                                              279.45, _dp(out))
        assert rc == 0
        mine = P.cell_kernel(i, j, 3, xl, u=u, weight=1.0)
        assert np.max(np.abs(mine)) > 0 and np.all(mine[8:] == 0.0)        # Cl has no channel, potential: Neumann 0
        assert np.max(np.abs(out - mine)) <= 1e-14 * np.max(np.abs(mine)), (i, from_cyt)
        if variant == "flux_stimulation":
            s0 = dict(s); s0["params"] = dict(p, do_memb_flux_stimulation=0)
            P0 = oracle_problem(oracle, s0)
            P0.set_time(t, dt); P0.set_uold(u0); P0.update_state(u0); P0.accept_state(); P0.update_state(u0 * (1.0 + 1e-3))
            changed += int(np.any(P0.cell_kernel(i, j, 3, xl, u=u, weight=1.0) != mine))
    if variant == "flux_stimulation":
        assert changed == 2          # exactly the two faces above position_x = 250 um (column 2, both sides)


def test_rownorm_against_reference_preconditioner(oracle, refop):
    """RowNormPreconditioner<GV, 4> (rownorm_preconditioner.hh:65-155, row a15) on the assembled Jacobian of a small
    problem incl. Dirichlet unit rows: bit-identical to the oracle's rownorm_scale (the device kernel is held
    bit-identical to the oracle in test_spmv_and_rownorm_parity)."""
    s = small_setup(nx=5, stimulation=True)
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    P.set_time(29.0, 1.0); P.set_uold(u0); P.update_state(u0)
    J = P.jacobian(u0)
    rowptr, col = P.pattern()
    rng = np.random.default_rng(9)
    r = rng.standard_normal(P.n)
    Jo, ro = P.rownorm(J, r)
    Jr, rr = J.copy(), r.copy()
    rc = refop.ax1ref_rownorm(P.nv, _ip(rowptr), _ip(col), _dp(Jr), _dp(rr))
    assert rc == 0
    assert np.array_equal(Jr, Jo) and np.array_equal(rr, ro)
    assert not np.array_equal(Jo, J)


# ------------------------------------------------------------------------------------------------------------------
# Ax1Newton: the reference's own line_search / prepare_step, compiled against stand-in PDELab Newton
# bases (oracle/ref_stubs/newton_standins.hh), on synthetic defect functions
@pytest.fixture(scope="module")
def refnewton(oracle):
    oracle.build_ref()
    R = oracle.ref_newton_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return R


def _line_search_case(name):
    """(u0, z, prev_defect, defect function of the trial u). lam(u) recovers the step length of a trial."""
    rng = np.random.default_rng(4)
    n = 8
    u0 = rng.standard_normal(n)
    e = rng.standard_normal(n)
    lam = lambda u, z: float(np.dot(u0 - u, z) / np.dot(z, z))
    if name == "full_step":            # u0 - z is the root: lambda = 1 passes the decrease test at once
        z = e.copy()
        return u0, z, np.linalg.norm(e), lambda u: np.linalg.norm(u - (u0 - e))
    if name == "overshoot":            # the root sits at lambda = 1/8: three halvings
        z = 8.0 * e
        return u0, z, np.linalg.norm(e), lambda u: np.linalg.norm(u - (u0 - e))
    if name == "never_improves":       # every trial is worse than where we stand: full step accepted anyway
        z = e.copy()
        return u0, z, 1.0, lambda u: 1.0 + np.linalg.norm(u - u0)
    if name == "best_in_middle":       # shallow minimum at lambda = 1/4, never a sufficient decrease
        z = e.copy()
        return u0, z, 1.0, lambda u: 0.99 + (lam(u, z) - 0.25) ** 2
    if name == "best_is_last":         # decreasing towards lambda -> 0 but never enough: best == last trial
        z = e.copy()
        return u0, z, 1.0, lambda u: 0.9999 + 1e-5 * lam(u, z)
    if name == "nan_then_ok":          # NaN defects for long steps are retried with a smaller lambda
        z = e.copy()
        return u0, z, 1.0, lambda u: float("nan") if lam(u, z) > 0.3 else 0.5 + lam(u, z)
    if name == "nan_everywhere_but_best":   # NaN must never become "best"
        z = e.copy()
        return u0, z, 1.0, lambda u: 0.97 if abs(lam(u, z) - 0.125) < 1e-12 else float("nan")
    raise KeyError(name)


@pytest.mark.parametrize("maxit", [10, 3])
@pytest.mark.parametrize("name", ["full_step", "overshoot", "never_improves", "best_in_middle", "best_is_last",
                                  "nan_then_ok", "nan_everywhere_but_best"])
def test_line_search_against_reference_newton(oracle, refnewton, name, maxit):
    """Ax1Newton::line_search (ax1_newton.hh:308-503), strategy hackbuschReuskenAcceptBestNoThrow as
    acme2_cyl_setup.hh:837-841 configures it: the same trial vectors in the same order, the same accepted u and
    defect as the oracle's line_search, bit for bit."""
    u0, z, prev, f = _line_search_case(name)
    n = len(u0)

    def run(which):
        seen = []

        def cb(ctx, up, nn):
            u = np.ctypeslib.as_array(up, shape=(nn,)).copy()
            seen.append(u)
            return float(f(u))
        cbf = oracle.DEFECT_CB(cb)
        u = u0.copy()
        if which == "ref":
            out = np.zeros(4)
            rc = refnewton.ax1ref_line_search(n, _dp(u), _dp(z.copy()), prev, prev, maxit, 0.5, 3, 0, cbf, None, None, 0,
                                              _dp(out))
            # a NaN defect at the finally accepted u escapes as NewtonDefectError (the oracle reports NaN and its
            # Newton loop turns that into NEWTON_DEFECT_NAN)
            assert rc == (3 if np.isnan(out[0]) else 0)
            assert out[1] == len(seen) and out[2] == len(seen)      # preAssembly before every defect (:346-349)
            assert out[3] == 1.0                                      # the solution container follows u
            return u, out[0], seen
        tr = np.zeros(64); ntr = C.c_int(0)
        d = oracle.lib().ax1o_line_search(n, _dp(u), _dp(z.copy()), prev, prev, maxit, cbf, None, _dp(tr), 64,
                                          C.byref(ntr))
        assert ntr.value == len(seen)
        return u, d, seen

    ur, dr, sr = run("ref")
    uo, do, so = run("oracle")
    assert len(sr) == len(so) and all(np.array_equal(a, b) for a, b in zip(sr, so))
    assert np.array_equal(ur, uo)
    assert dr == do or (np.isnan(dr) and np.isnan(do))
    # the scenario really exercises the branch it is named after
    if maxit == 10:
        expect = {"full_step": 1, "overshoot": 4, "never_improves": 12, "best_in_middle": 11, "best_is_last": 10,
                  "nan_then_ok": 3, "nan_everywhere_but_best": 11}[name]
        assert len(sr) == expect, (name, len(sr))


def test_line_search_other_strategies_of_reference(oracle, refnewton):
    """The strategies the product does not offer behave as documented in the reference (so that the one it does
    offer is known to be the right one): noLineSearch takes the full step with one defect evaluation; plain
    Hackbusch-Reusken and AcceptBest throw NewtonLineSearchError and restore u when nothing improved."""
    u0, z, prev, f = _line_search_case("never_improves")
    n = len(u0)
    for strategy, rc_expect, u_expect, evals in ((0, 0, u0 - z, 1), (1, 1, u0, 11), (2, 1, u0, 11), (3, 0, u0 - z, 12)):
        cbf = oracle.DEFECT_CB(lambda ctx, up, nn: float(f(np.ctypeslib.as_array(up, shape=(nn,)))))
        u = u0.copy(); out = np.zeros(4)
        rc = refnewton.ax1ref_line_search(n, _dp(u), _dp(z.copy()), prev, prev, 10, 0.5, strategy, 0, cbf, None, None, 0,
                                          _dp(out))
        assert rc == rc_expect and out[1] == evals
        assert np.array_equal(u, u_expect)


def test_prepare_step_applies_rownorm_of_reference(oracle, refnewton):
    """Ax1Newton::prepare_step (ax1_newton.hh:216-293) with setRowPreconditioner(true): (A, r) leave it equilibrated
    exactly as the oracle's rownorm_scale does; with it off they are untouched."""
    s = small_setup(nx=5, stimulation=True)
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    P.set_time(29.0, 1.0); P.set_uold(u0); P.update_state(u0)
    J = P.jacobian(u0)
    rowptr, col = P.pattern()
    r = np.random.default_rng(9).standard_normal(P.n)
    Jo, ro = P.rownorm(J, r)
    Jr, rr = J.copy(), r.copy()
    assert refnewton.ax1ref_prepare_step(P.nv, _ip(rowptr), _ip(col), _dp(Jr), _dp(rr), 1) == 0
    assert np.array_equal(Jr, Jo) and np.array_equal(rr, ro)
    Jr, rr = J.copy(), r.copy()
    assert refnewton.ax1ref_prepare_step(P.nv, _ip(rowptr), _ip(col), _dp(Jr), _dp(rr), 0) == 0
    assert np.array_equal(Jr, J) and np.array_equal(rr, r)


def _ref_boundary(refop, P, s, u, st, i, from_cyt, t, dt):
    """alpha_boundary of the reference's fully-coupled operator on the membrane-interface face of cell column i."""
    cst = P.constants()
    _, nvol, _ = P.maps()
    nvx, x, y, jm, p = s["nx"] + 1, s["x"], s["y"], P.jm, s["params"]
    j = jm - 1 if from_cyt else jm + 1
    vs = [i + nvx * j, i + 1 + nvx * j, i + nvx * (j + 1), i + 1 + nvx * (j + 1)]
    xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
    jopp = jm + 1 if from_cyt else jm
    ua, ub = u[4 * (i + nvx * jopp):][:4], u[4 * (i + 1 + nvx * jopp):][:4]
    opp = 0.5 * ua + 0.5 * ub
    memb = np.array([y[jopp], opp[3], opp[0], opp[1], opp[2], (np.pi * (y[jm + 1] + y[jm + 1])) * (x[i + 1] - x[i])])
    gl, gn, gk = s["g_leak"][i], s["g_nav"][i], s["g_kv"][i]
    geff = np.array([gl * st[0, i], gl * st[1, i], gn * st[2, i] ** 3 * st[3, i], gk * st[4, i] ** 4])
    equil = np.array([p["do_equilibration"], t, p["t_equilibrium"], p["dummy_value"]], dtype=np.float64)
    fst = np.array([p.get("do_memb_flux_stimulation", 0), p.get("memb_flux_ion", 2), p.get("memb_flux_inj", 0.0)],
                   dtype=np.float64)
    stim = np.array([0, t + dt, p["tinj_start"], p["tinj_end"], p["stim_x"], p["stim_y"], p["I_inj"]], dtype=np.float64)
    cell = np.array([x[i], x[i + 1], y[j], y[j + 1]])
    phys = np.array([80.0, cst["D"][0], cst["D"][1], cst["D"][2], cst["PC"], 1.0, 1e-6])
    val = np.array([1, 1, -1], dtype=np.int32)
    nv4 = np.array([nvol[v] for v in vs])
    out = np.zeros(16)
    rc = refop.ax1ref_elec_alpha_boundary(_dp(cell), from_cyt, _dp(np.ascontiguousarray(xl)), _dp(phys), _ip(val),
                                          _dp(nv4), 1, _dp(stim), _dp(memb), _dp(geff), _dp(equil), _dp(fst),
                                          279.45, _dp(out))
    assert rc == 0
    return vs, out


def test_global_residual_is_scatter_of_reference_element_operators(oracle, refop):
    """Row a13 as far as it can be pinned without PDELab: the oracle's GLOBAL residual of one implicit-Euler step equals
    the reference's own compiled element operators (alpha_volume and alpha_boundary of the fully-coupled operator with
    the implicit membrane flux, the time operator and the membrane Poisson operator) evaluated on every cell / every
    membrane-interface face of the grid and scattered here in numpy as r = dt * A(u) + M(u) - M(u_old), constrained
    rows zeroed. Checks the gather / scatter indexing, the DOF layout, the implicit-Euler combination and the
    constraint handling of the oracle's assembler on all 4 x 180 cells, every row."""
    s = small_setup(nx=4, stimulation=True)
    t, dt = 29.0, 1.0
    run = _CellRunner(oracle, refop, s, t=t, dt=dt, stim=True)
    P, jm, nvx = run.P, run.P.jm, s["nx"] + 1
    rng = np.random.default_rng(12)
    uold = s["u0"]
    u = uold * (1.0 + 1e-3 * rng.standard_normal(uold.shape))
    _, _, dmask = P.maps()
    st = P.channel_state(new=True).reshape(5, -1)
    cs = np.asarray(s["cell_subdomain"]).reshape(s["ny"], s["nx"])
    r_py = np.zeros(P.n)
    r_bnd = np.zeros(P.n)
    for j in range(s["ny"]):
        for i in range(s["nx"]):
            vs = run.verts(i, j)
            xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
            xo = np.array([[uold[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
            if cs[j, i] == 2:                       # membrane cell: Poisson with the membrane permittivity
                rl = dt * run.ref(2, 0, i, j, xl, 2.0)
            else:
                rl = dt * run.ref(0, 0, i, j, xl, 80.0) + run.ref(1, 0, i, j, xl, 80.0) - run.ref(1, 0, i, j, xo, 80.0)
            for c in range(4):
                for k, v in enumerate(vs):
                    r_py[4 * v + c] += rl[4 * c + k]
    for i in range(s["nx"]):
        for from_cyt in (1, 0):
            vs, out = _ref_boundary(refop, P, s, u, st, i, from_cyt, t, dt)
            for c in range(4):
                for k, v in enumerate(vs):
                    r_bnd[4 * v + c] += dt * out[4 * c + k]
    r_py += r_bnd
    for v in range(P.nv):
        for c in range(4):
            if (int(dmask[v]) >> c) & 1:
                r_py[4 * v + c] = 0.0
    r_or, S = P.residual(u, with_abs=True)
    err = (np.abs(r_py - r_or) / (S + 1e-300)).reshape(-1, 4)
    assert err.max() < 1e-13, err.max()
    # the membrane flux really is in the compared numbers, on the two interface vertex lines only
    touched = np.nonzero(np.abs(r_bnd).reshape(-1, 4).max(axis=1) > 0)[0]
    assert len(touched) == 2 * nvx and set((touched // nvx).tolist()) == {jm, jm + 1}


def test_global_jacobian_is_scatter_of_reference_element_jacobians(oracle, refop):
    """The same for the assembled matrix: the oracle's global analytic Jacobian (block CSR) equals the reference's
    compiled jacobian_volume methods of the three local operators scattered over all cells as dt * dA + dM, on every
    unconstrained row outside the two membrane-interface vertex lines (whose concentration rows also carry the
    finite-difference boundary term)."""
    s = small_setup(nx=4, stimulation=True)
    t, dt = 29.0, 1.0
    run = _CellRunner(oracle, refop, s, t=t, dt=dt, stim=True)
    P, jm, nvx = run.P, run.P.jm, s["nx"] + 1
    rng = np.random.default_rng(13)
    u = s["u0"] * (1.0 + 1e-3 * rng.standard_normal(s["u0"].shape))
    _, _, dmask = P.maps()
    cs = np.asarray(s["cell_subdomain"]).reshape(s["ny"], s["nx"])
    rowptr, col = P.pattern()
    pos = {(int(r), int(col[k])): k for r in range(P.nv) for k in range(rowptr[r], rowptr[r + 1])}
    J_py = np.zeros((len(col), 4, 4))
    for j in range(s["ny"]):
        for i in range(s["nx"]):
            vs = run.verts(i, j)
            xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
            if cs[j, i] == 2:
                m = dt * run.ref(2, 1, i, j, xl, 2.0).reshape(16, 16)
            else:
                m = dt * run.ref(0, 1, i, j, xl, 80.0).reshape(16, 16) + run.ref(1, 1, i, j, xl, 80.0).reshape(16, 16)
            for ci in range(4):
                for vi in range(4):
                    for cj in range(4):
                        for vj in range(4):
                            val = m[ci * 4 + vi, cj * 4 + vj]
                            if val != 0.0:
                                J_py[pos[(vs[vi], vs[vj])], ci, cj] += val
    J_or = np.asarray(P.jacobian(u)).reshape(-1, 4, 4)
    rowof = np.repeat(np.arange(P.nv), np.diff(rowptr))
    scale = np.zeros((P.nv, 4)); np.maximum.at(scale, rowof, np.abs(J_or).max(axis=2))
    keep = np.ones((P.nv, 4), dtype=bool)
    for jj in (jm, jm + 1):
        keep[jj * nvx:(jj + 1) * nvx, :3] = False
    for v in range(P.nv):
        for c in range(4):
            if (int(dmask[v]) >> c) & 1:
                keep[v, c] = False
    rel = np.abs(J_py - J_or).max(axis=2) / (scale[rowof] + 1e-300)        # per (block, scalar row)
    assert rel[keep[rowof]].max() < 1e-13, rel[keep[rowof]].max()
    assert keep.sum() > 0.9 * keep.size


@pytest.mark.parametrize("variant", ["hh", "equilibration"])
def test_membrane_flux_output_terms_against_reference_flux_class(oracle, refop, variant):
    """The five per-ion output terms of MembraneFluxGridFunctionImplicit::evaluate (total, diffusion term, drift term,
    leak, voltage-gated; membranefluxgridfunction_implicit.hh:42, 83-361 -- the data behind memb_flux*.dat /
    the HDF5 datasets of acme2_cyl_output.hh:1215-1410), evaluated by the reference's own class on the face-centre
    values of both interface edges: the oracle's membrane_flux_terms agrees to rounding in every column."""
    s = small_setup(nx=6, equilibration=(variant == "equilibration"))
    P = oracle_problem(oracle, s)
    u0, t, dt = s["u0"], 399.0, 1.0
    P.set_time(t, dt); P.set_uold(u0); P.update_state(u0); P.accept_state()
    P.update_state(u0 * (1.0 + 1e-3))                 # a gating step away from the steady state
    rng = np.random.default_rng(5)
    u = u0 * (1.0 + 1e-2 * rng.standard_normal(u0.shape))
    mine = P.membrane_flux_terms(u)
    st = P.channel_state(new=True).reshape(5, -1)
    nvx, y, jm, p = s["nx"] + 1, s["y"], P.jm, s["params"]
    assert np.abs(mine[0, :2]).min() > 0 and np.all(mine[:, 2] == 0.0)       # Cl has no channel
    if variant == "equilibration":
        assert np.all(mine[4] == 0.0)                                          # voltage-gated channels are off
    else:
        assert np.abs(mine[4, :2]).min() > 0
    for i in range(s["nx"]):
        cyt = 0.5 * u[4 * (i + nvx * jm):][:4] + 0.5 * u[4 * (i + 1 + nvx * jm):][:4]
        es = 0.5 * u[4 * (i + nvx * (jm + 1)):][:4] + 0.5 * u[4 * (i + 1 + nvx * (jm + 1)):][:4]
        memb = np.array([y[jm + 1], es[3], es[0], es[1], es[2], 1.0])
        gl, gn, gk = s["g_leak"][i], s["g_nav"][i], s["g_kv"][i]
        geff = np.array([gl * st[0, i], gl * st[1, i], gn * st[2, i] ** 3 * st[3, i], gk * st[4, i] ** 4])
        equil = np.array([p["do_equilibration"], t, p["t_equilibrium"], p["dummy_value"]], dtype=np.float64)
        uin = np.array([cyt[3], cyt[0], cyt[1], cyt[2]])
        out = np.zeros(15)
        assert refop.ax1ref_membrane_flux_terms(1, _dp(uin), _dp(memb), _dp(geff), _dp(equil), 279.45, _dp(out)) == 0
        ref = out.reshape(5, 3)
        scale = np.abs(ref).max(axis=0, keepdims=True) + 1e-300
        assert np.max(np.abs(ref - mine[:, :, i]) / scale) < 1e-14, i
        assert np.allclose(ref[0], ref[2] - ref[1], rtol=1e-12, atol=0) and np.allclose(ref[0], ref[3] + ref[4], rtol=1e-12, atol=0)


# ---- electroneutral (Mori) model: the reference's split operators, parameter classes and charge layer --------------
@pytest.fixture(scope="module")
def refmori(oracle):
    oracle.build_ref()
    R = oracle.ref_mori_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return R


def test_mori_charge_layer_states_against_reference_class(oracle, refmori):
    """ChargeLayerMoriImplicit::init / timeStep / acceptTimeStep (charge_layer_mori_implicit.hh), compiled from the
    reference: gamma_j = z_j^2 c_j / sum z^2 c at initialisation, relaxation with tau = 1 ns afterwards. The oracle's
    mori_update_state / accept on a one-column problem driven with the same membrane concentrations: bit-identical."""
    s = small_setup(nx=1)
    P = oracle_problem(oracle, s)
    jm, nvx = P.jm, 2
    rng = np.random.default_rng(21)
    us = [s["u0"] * (1.0 + 5e-2 * rng.standard_normal(s["u0"].shape)) for _ in range(4)]
    con = []
    for u in us:
        cc = [0.5 * u[4 * (0 + nvx * jm) + c] + 0.5 * u[4 * (1 + nvx * jm) + c] for c in range(3)]
        ce = [0.5 * u[4 * (0 + nvx * (jm + 1)) + c] + 0.5 * u[4 * (1 + nvx * (jm + 1)) + c] for c in range(3)]
        con += cc + ce
    con = np.array(con)
    dt = 0.37
    out = np.zeros(6 * 4)
    val = np.array([1, 1, -1], dtype=np.int32)
    assert refmori.ax1ref_mori_gamma(_dp(con), 3, dt * 1e-6, _ip(val), _dp(out)) == 0
    ref = out.reshape(4, 2, 3)
    for k, u in enumerate(us):
        P.set_time(float(k), dt)
        P.mori_update_state(u)
        g = P.mori_gamma()                       # [side][old/new][ion][column]
        assert np.array_equal(g[:, 1, :, 0], ref[k]), k
        P.accept_state()
        assert np.array_equal(P.mori_gamma()[:, 0, :, 0], ref[k])
    assert np.allclose(ref.sum(axis=2), 1.0, atol=1e-15)


def _mori_scatter(oracle, refmori, which, s, P, u, ulast, uold, t, dt, gamma, active=1.0, neumann=1.0):
    """global residual of Mori sub-problem `which` as the numpy scatter of the reference's compiled alpha_volume /
    alpha_boundary on every electrolyte cell / membrane-interface face"""
    cst = P.constants()
    _, nvol, dmask = P.maps()
    nvx, x, y, jm, p = s["nx"] + 1, s["x"], s["y"], P.jm, s["params"]
    st = P.channel_state(new=True).reshape(5, -1)
    cs = np.asarray(s["cell_subdomain"]).reshape(s["ny"], s["nx"])
    val = np.array([1, 1, -1], dtype=np.int32)
    phys = np.array([80.0, cst["D"][0], cst["D"][1], cst["D"][2], cst["PC"], 1.0, 1e-6])
    stim = np.array([p["do_stimulation"], t + dt, p["tinj_start"], p["tinj_end"], p["stim_x"], p["stim_y"], p["I_inj"]],
                    dtype=np.float64)
    misc = np.array([279.45, dt, t, 5.0, neumann, active])
    r = np.zeros(P.n)
    loc = lambda vec, vs: np.ascontiguousarray(np.array([[vec[4 * v + c] for v in vs] for c in range(4)]).reshape(-1))
    xnew_vec = u if which == 0 else ulast          # what getSolutionPotNew holds in this phase
    for j in range(s["ny"]):
        for i in range(s["nx"]):
            if cs[j, i] == 2:
                continue                           # no membrane contributions
            vs = [i + nvx * j, i + 1 + nvx * j, i + nvx * (j + 1), i + 1 + nvx * (j + 1)]
            xl = loc(u, vs)
            # SolutionContainer on this cell: concentrations always from ulast ("ConNew"), the potential from the
            # phase's PotNew vector
            xnew = loc(ulast, vs)
            xnew[12:] = loc(xnew_vec, vs)[12:]
            xold = loc(uold, vs)
            nv4 = np.array([nvol[v] for v in vs])
            cell = np.array([x[i], x[i + 1], y[j], y[j + 1]])
            gl, gn, gk = s["g_leak"][i], s["g_nav"][i], s["g_kv"][i]
            geff = np.array([gl * st[0, i], gl * st[1, i], gn * st[2, i] ** 3 * st[3, i], gk * st[4, i] ** 4])
            for boundary in (0, 1):
                if boundary and j not in (jm - 1, jm + 1):
                    continue
                from_cyt = 1 if j == jm - 1 else 0
                jthis, jopp = (jm, jm + 1) if from_cyt else (jm + 1, jm)
                fc = lambda vec, row, c: 0.5 * vec[4 * (i + nvx * row) + c] + 0.5 * vec[4 * (i + 1 + nvx * row) + c]
                memb = np.array([y[jopp], fc(xnew_vec, jopp, 3), fc(ulast, jopp, 0), fc(ulast, jopp, 1), fc(ulast, jopp, 2),
                                 (np.pi * (y[jm + 1] + y[jm + 1])) * (x[i + 1] - x[i]), fc(uold, jthis, 3), fc(uold, jopp, 3),
                                 s["eps_memb"][i]])
                side = 0 if from_cyt else 1
                gam = np.ascontiguousarray(np.concatenate([gamma[side, 0, :, i], gamma[side, 1, :, i]]))
                out = np.zeros(16)
                rc = refmori.ax1ref_mori_operator(which, boundary, _dp(cell), from_cyt, _dp(xl), _dp(xnew), _dp(xold),
                                                  _dp(phys), _ip(val), _dp(nv4), 1, _dp(stim), _dp(memb), _dp(geff),
                                                  _dp(gam), _dp(misc), _dp(out))
                assert rc == 0
                w = dt if which == 1 else 1.0      # OneStepMethod weights the spatial concentration operator with dt
                for c in range(4):
                    for k, v in enumerate(vs):
                        r[4 * v + c] += w * out[4 * c + k]
    for v in range(P.nv):
        for c in range(4):
            if (int(dmask[v]) >> c) & 1:
                r[4 * v + c] = 0.0
    return r


@pytest.mark.parametrize("which", [0, 1])
def test_mori_global_residual_is_scatter_of_reference_split_operators(oracle, refop, refmori, which):
    """Row f4: the oracle's global residual of the Mori potential (which = 0) / concentration (1) sub-problem equals the
    reference's own Acme2CylMoriPotentialOperator / Acme2CylMoriConcentrationOperator (alpha_volume on every
    electrolyte cell, alpha_boundary on every membrane-interface face) with its MoriPoissonParameters,
    MoriNernstPlanckParameters (channel flux + capacitive flux, surface scaling, Na injection), ChargeLayerMoriImplicit
    and MembraneFluxGridFunctionImplicit, compiled from /root/reference and scattered here in numpy; for the
    concentration problem plus the (already pinned) time operator: r = dt A(c) + M(c) - M(c_old). Three different
    vectors play the three roles (unknowns, last iteration, previous time step), the charge layer has been
    initialised and advanced so that gamma_new != gamma_old."""
    s = small_setup(nx=3, stimulation=True)
    t, dt = 29.0, 0.8
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    rng = np.random.default_rng(31)
    uold = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    ulast = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    u = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    P.set_time(t - dt, dt); P.set_uold(u0); P.update_state(u0); P.accept_state()
    P.mori_update_state(u0); P.accept_state()
    P.set_time(t, dt); P.set_uold(uold); P.update_state(ulast)
    P.mori_update_state(ulast)
    gamma = P.mori_gamma()
    assert np.max(np.abs(gamma[:, 1] - gamma[:, 0])) > 1e-6
    r_py = _mori_scatter(oracle, refmori, which, s, P, u, ulast, uold, t, dt, gamma)
    if which == 1:                                 # + M(c) - M(c_old): the reference's time operator (refop)
        run = _CellRunner(oracle, refop, s, t=t, dt=dt, stim=True)
        run.P = P
        _, _, dmask = P.maps()
        cs = np.asarray(s["cell_subdomain"]).reshape(s["ny"], s["nx"])
        for j in range(s["ny"]):
            for i in range(s["nx"]):
                if cs[j, i] == 2:
                    continue
                vs = run.verts(i, j)
                xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
                xo = np.array([[uold[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
                rl = run.ref(1, 0, i, j, xl, 80.0) - run.ref(1, 0, i, j, xo, 80.0)
                for c in range(3):
                    for k, v in enumerate(vs):
                        if not (int(dmask[v]) >> c) & 1:
                            r_py[4 * v + c] += rl[4 * c + k]
    r_or, _ = P.mori_assemble(which, u, ulast, jacobian=False)
    # scale: sensitivity of each row to a relative perturbation of the inputs (what the row's terms add up to)
    scale = np.zeros_like(r_or)
    prng = np.random.default_rng(5)
    for _ in range(3):
        pu, pl, po = (v * (1.0 + 1e-9 * prng.standard_normal(v.size)) for v in (u, ulast, uold))
        P.set_uold(po)
        rp, _ = P.mori_assemble(which, pu, pl, jacobian=False)
        scale = np.maximum(scale, np.abs(rp - r_or) / 1e-9)
    P.set_uold(uold)
    err = np.abs(r_py - r_or) / (scale + np.abs(r_or) + 1e-300)
    assert err.max() < 1e-12, (which, err.max(), int(err.argmax()))
    rows = r_or.reshape(-1, 4)
    assert np.all(rows[:, :3] == 0.0) if which == 0 else np.all(rows[:, 3] == 0.0)
    assert np.abs(rows).max() > 0.0


def test_potential_row_is_one_number_times_the_valences(oracle, refop):
    """What the device's 64-byte-block SpMV operand rests on: in the analytic Jacobian the potential row
    couples to the three concentrations through the charge density only, so its entries are z_k times one number --
    bit for bit, because the valences are +-1. Checked (a) on the reference's own compiled element Jacobians
    (jacobian_volume of the electrolyte operator), (b) on the oracle's assembled matrix, with and without the row-norm
    scaling; (c) the finite-difference volume Jacobian (the reference's default mode) does NOT have that shape
    exactly, which is why the device checks every block before it uses the operand."""
    s = small_setup(nx=4, stimulation=True)
    t, dt = 29.0, 1.0
    run = _CellRunner(oracle, refop, s, t=t, dt=dt, stim=True)
    P = run.P
    z = np.array([1.0, 1.0, -1.0])          # Na, K, Cl (dune/ax1/common/constants.hh, default configuration)
    assert list(np.asarray(s["params"].get("valence", [1, 1, -1]), dtype=float)) == list(z)
    rng = np.random.default_rng(5)
    u = s["u0"] * (1.0 + 1e-3 * rng.standard_normal(s["u0"].shape))
    cs = np.asarray(s["cell_subdomain"]).reshape(s["ny"], s["nx"])
    checked = 0
    for j in range(s["ny"]):
        for i in range(s["nx"]):
            if cs[j, i] == 2:
                continue
            vs = run.verts(i, j)
            xl = np.array([[u[4 * v + c] for v in vs] for c in range(4)]).reshape(-1)
            m = run.ref(0, 1, i, j, xl, 80.0).reshape(4, 4, 4, 4)       # [row comp, row vertex, col comp, col vertex]
            q = m[3, :, 0, :] / z[0]
            for k in range(3):
                assert np.array_equal(z[k] * q, m[3, :, k, :])
            checked += np.count_nonzero(q)
    assert checked > 0
    J = np.asarray(P.jacobian(u)).reshape(-1, 4, 4)
    r = P.residual(u)
    Js, _ = P.rownorm(J.reshape(-1), r)
    for M in (J, np.asarray(Js).reshape(-1, 4, 4)):
        q = M[:, 3, 0] / z[0]
        assert np.count_nonzero(q) > 0.5 * len(q)
        for k in range(3):
            assert np.array_equal(z[k] * q, M[:, 3, k])
    # finite-difference volume Jacobians: proportional only to rounding
    s2 = small_setup(nx=4, stimulation=True)
    s2["params"].update(use_elec_jacobian_volume=0, use_memb_jacobian_volume=0, use_time_jacobian_volume=0)
    P2 = oracle_problem(oracle, s2, analytic=0)
    P2.set_time(t, dt); P2.set_uold(s2["u0"]); P2.update_state(s2["u0"])
    F = np.asarray(P2.jacobian(u)).reshape(-1, 4, 4)
    q = F[:, 3, 0] / z[0]
    exact = all(np.array_equal(z[k] * q, F[:, 3, k]) for k in range(3))
    close = all(np.allclose(z[k] * q, F[:, 3, k], rtol=1e-5, atol=1e-5 * np.abs(F[:, 3, :3]).max()) for k in range(3))
    assert close and not exact
