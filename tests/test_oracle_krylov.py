"""CPU checks of the oracle's second linear solver backend (RestartedGMRes + multigrid cycle,
dune/ax1/common/ax1_solverbackend.hh:73-302): the GMRES restatement against scipy's independent
implementation of left-preconditioned restarted GMRES, and the multigrid cycle as a linear operator."""
import numpy as np
import scipy.sparse.linalg as sla

from helpers import oracle_problem, small_setup


def _problem(oracle, nx, dx):
    s = small_setup(nx=nx, dx=dx)
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    P.set_time(0.0, 1.0)
    P.set_uold(u0)
    P.update_state(u0)
    return P, P.jacobian(u0), P.residual(u0)


def test_gmres_matches_scipy(oracle):
    P, J, r = _problem(oracle, 10, 100.0e3)
    n = P.n
    A = sla.LinearOperator((n, n), matvec=lambda x: P.spmv(J, x))
    M = sla.LinearOperator((n, n), matvec=lambda x: P.ilu_apply(J, x, ilu_fill=0, slab_w=4))
    MA = sla.LinearOperator((n, n), matvec=lambda x: M.matvec(A.matvec(x)))
    for restart, red in ((100, 1e-9), (9, 1e-6)):    # GMRES(9) stagnates near 1e-8 on this matrix
        z, rr, res = P.solve_ex(J, r, red, maxit=400, ilu_fill=0, slab_w=4, krylov=1, restart=restart)
        assert res.converged == 1
        hist = []
        zs, info = sla.gmres(MA, M.matvec(r), rtol=red, atol=0.0, restart=restart, maxiter=400,
                             callback=lambda rk: hist.append(rk), callback_type="pr_norm")
        assert info == 0
        assert abs(len(hist) - res.iterations) <= 1, (len(hist), res.iterations)
        assert np.max(np.abs(z - zs)) <= 1e3 * red * np.max(np.abs(zs))
        # b is left holding the true defect
        assert np.max(np.abs(rr - (r - P.spmv(J, z)))) <= 1e-9 * np.max(np.abs(r))


def test_amg_cycle_is_linear_and_helps(oracle):
    P, J, r = _problem(oracle, 48, 100.0)
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(P.n), rng.standard_normal(P.n)
    va, vb = P.amg_apply(J, a, 0, 8, 4), P.amg_apply(J, b, 0, 8, 4)
    vab = P.amg_apply(J, 2.0 * a - 3.0 * b, 0, 8, 4)
    assert np.max(np.abs(vab - (2.0 * va - 3.0 * vb))) <= 1e-9 * np.max(np.abs(vab))
    # a coarse correction must not touch constrained (Dirichlet) DOFs beyond what the smoother does
    _, _, dm = P.maps()
    con = np.repeat(dm[:, None], 4, axis=1) >> np.arange(4)[None, :] & 1
    d = np.zeros(P.n)
    v = P.amg_apply(J, d, 0, 8, 4)
    assert np.all(v == 0.0)
    assert con.any()
    z0, _, r0 = P.solve_ex(J, r, 1e-8, maxit=2000, ilu_fill=0, slab_w=8)
    z1, _, r1 = P.solve_ex(J, r, 1e-8, maxit=2000, ilu_fill=0, slab_w=8, amg=1, amg_cx=4)
    assert r0.converged == 1 and r1.converged == 1
    assert r1.iterations * 3 < r0.iterations, (r1.iterations, r0.iterations)
    nr = np.linalg.norm(r)
    assert np.linalg.norm(r - P.spmv(J, z0)) <= 2e-8 * nr and np.linalg.norm(r - P.spmv(J, z1)) <= 2e-8 * nr


def test_newton_with_gmres_amg_reaches_same_fields(oracle):
    s = small_setup(nx=48, dx=100.0)
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    P.set_time(0.0, 1.0)
    P.set_uold(u0)
    P.update_state(u0)
    o0 = oracle.default_newton_opts(reduction=1e-8, abs_limit=1e-14, min_lin_reduction=1e-6, ilu_fill=1, slab_w=0)
    o1 = oracle.default_newton_opts(reduction=1e-8, abs_limit=1e-14, min_lin_reduction=1e-6, ilu_fill=0, slab_w=8,
                                    krylov=1, restart=100, amg=1, amg_cx=4, lin_maxit=500)
    ua, ra = P.newton(u0, o0)
    P2 = oracle_problem(oracle, s)
    P2.set_time(0.0, 1.0)
    P2.set_uold(u0)
    P2.update_state(u0)
    ub, rb = P2.newton(u0, o1)
    assert ra.status == 0 and rb.status == 0
    scale = np.max(np.abs(ua.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(ua - ub).reshape(-1, 4) / scale, axis=0)
    # the potential is the worst-conditioned unknown; two different preconditioners at a 1e-6 linear
    # tolerance agree on it to ~1e-5, on the concentrations to ~1e-9
    assert np.all(err[:3] < 1e-8) and err[3] < 1e-4, err
