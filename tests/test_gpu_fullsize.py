"""Size-independent properties at the FULL size of the named workload (BASELINE.json metric config: Ny = 180,
Nx = 93,184 cell columns, 16.87 M vertices, 67.5 M DOF on one GPU), where the CPU oracle cannot be run in test
time. Everything goes through the C ABI with host buffers. Properties:
  * bitwise determinism of residual, defect norm and a whole fixed-work Newton step (atomic-free assembly,
    fixed-order reductions);
  * axial translation symmetry: for an x-independent state every interior column of the residual is the same;
  * the assembled Jacobian is the derivative of the assembled residual (directional difference quotient);
  * linearity of the SpMV and of both preconditioners (ILU0, ILU(1)) as operators;
  * the linear solvers deliver what they report: true residual |r - A z| <= reduction * |r| (BiCGStab), and the
    preconditioned residual for GMRES;
  * the row-norm equilibration leaves the solution of the linear system unchanged."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NX = 93184


@pytest.fixture(scope="module")
def full(axlib):
    import torch
    free, total = torch.cuda.mem_get_info()
    if total < 100 * 2 ** 30:
        pytest.skip("needs a GPU with > 100 GB (B200)")
    from dune_ax1_b200 import setups
    s = setups.make_setup(NX, dx=100.0, perturbation=1e-3)
    dev = setups.make_device(s)
    dev.set_time(0.0, 1.0)
    dev.set_uold(s["u0"])
    dev.update_state(s["u0"])
    yield dev, s
    dev.close()


def test_fullsize_determinism_and_jacobian_consistency(full, axlib):
    dev, s = full
    u0 = s["u0"]
    r1, n1 = dev.residual(u0)
    r2, n2 = dev.residual(u0)
    assert np.array_equal(r1, r2) and n1 == n2 and np.isfinite(n1)
    # J d  vs  (R(u + e d) - R(u - e d)) / 2e  for a smooth direction d (relative to the row scale |J| |d|)
    nvx, nvy = s["nx"] + 1, s["ny"] + 1
    i = np.arange(nvx)[None, :, None]
    j = np.arange(nvy)[:, None, None]
    dev.jacobian(u0, download=False)
    dm = s["dirichlet"]
    con = (np.repeat(dm[:, None], 4, axis=1) >> np.arange(4)[None, :] & 1).reshape(nvy, nvx, 4).astype(bool)
    jm = int(np.nonzero(s["cell_subdomain"].reshape(s["ny"], s["nx"])[:, 0] == 2)[0][0])
    smooth = np.sin(2 * np.pi * i / 53.0) * np.cos(2 * np.pi * j / 17.0)
    eps = 1e-6
    for comp in range(4):
        # one unknown at a time: a direction that moves all three concentrations by the same relative amount keeps
        # the bulk electroneutral and makes the potential rows of J d cancel to nothing
        d3 = np.zeros((nvy, nvx, 4))
        # concentrations: relative perturbation; potential (zero at the Dirichlet top): absolute perturbation
        d3[:, :, comp] = (u0.reshape(nvy, nvx, 4)[:, :, comp] if comp < 3 else 1.0) * smooth[:, :, 0]
        d = d3.reshape(-1)
        Jd = dev.spmv(d)
        rp, _ = dev.residual(u0 + eps * d)
        rm, _ = dev.residual(u0 - eps * d)
        fd = (rp - rm) / (2 * eps)
        # scale: the largest |J d| among the rows of the same grid line j and component (d is smooth and periodic
        # in i, so these rows carry comparable terms)
        Jd3, fd3 = Jd.reshape(nvy, nvx, 4), fd.reshape(nvy, nvx, 4)
        scale = np.max(np.abs(Jd3), axis=1, keepdims=True)
        err = np.abs(Jd3 - fd3) / (scale + 1e-300)
        err[con] = 0.0
        err[np.broadcast_to(scale < 1e-200, err.shape)] = 0.0      # rows this unknown does not couple to
        # the membrane-interface flux term is a quasi-Newton block in the reference (the opposite membrane side is
        # not differentiated): the concentration rows of the two interface grid lines are no exact derivatives
        err[jm:jm + 2, :, :3] = 0.0
        assert err.max() < 1e-5, (comp, err.max(), np.unravel_index(np.argmax(err), err.shape))
        assert np.array_equal(Jd[con.reshape(-1)], d[con.reshape(-1)])


def test_fullsize_translation_symmetry(axlib):
    import torch
    if torch.cuda.mem_get_info()[1] < 100 * 2 ** 30:
        pytest.skip("needs a GPU with > 100 GB (B200)")
    from dune_ax1_b200 import setups
    s = setups.make_setup(NX, dx=100.0, perturbation=0.0)
    dev = setups.make_device(s)
    u0 = s["u0"]
    dev.set_time(0.0, 1.0); dev.set_uold(u0); dev.update_state(u0)
    r, _ = dev.residual(u0)
    nvx, nvy = s["nx"] + 1, s["ny"] + 1
    r = r.reshape(nvy, nvx, 4)
    spread = np.max(np.abs(r[:, 1:-1, :] - r[:, 1:2, :]), axis=1)          # per (row j, component)
    # the x vector is built by repeated addition, so hx differs in the last bits from column to column:
    # columns agree to rounding of the summed term magnitudes (~1e16 on the membrane rows), not bitwise
    mag = np.max(np.abs(u0)) * 1e15
    assert spread.max() <= 1e-9 * mag, spread.max()
    # and the two membrane-interface rows see no axial current at all for an x-independent state
    vm = dev.membrane_potential(u0)
    assert np.max(np.abs(vm - vm[0])) < 1e-9
    dev.close()


@pytest.mark.parametrize("fill", [0, 1])
def test_fullsize_operators_are_linear(full, axlib, fill):
    dev, s = full
    dev.jacobian(s["u0"], download=False)
    rng = np.random.default_rng(5 + fill)
    x = rng.standard_normal(dev.n)
    y = rng.standard_normal(dev.n)
    lhs = dev.spmv(x + 2.0 * y)
    rhs = dev.spmv(x) + 2.0 * dev.spmv(y)
    assert np.max(np.abs(lhs - rhs)) <= 1e-12 * np.max(np.abs(rhs))
    dev.ilu_factor(128, fill=fill)
    lhs = dev.ilu_apply(x - 3.0 * y)
    rhs = dev.ilu_apply(x) - 3.0 * dev.ilu_apply(y)
    assert np.max(np.abs(lhs - rhs)) <= 1e-10 * np.max(np.abs(rhs))
    # the preconditioner approximates the inverse: |x - M^-1 A x| is well below |x| for a smooth x
    xs = np.tile(np.sin(np.arange(s["nx"] + 1) / 200.0)[None, :, None], (s["ny"] + 1, 1, 4)).reshape(-1)
    zs = dev.ilu_apply(dev.spmv(xs))
    assert np.linalg.norm(zs - xs) < 0.9 * np.linalg.norm(xs)
    dev.ilu_factor(128, fill=0)


def test_fullsize_linear_solvers_deliver_their_reduction(full, axlib):
    dev, s = full
    u0 = s["u0"]
    r, nrm = dev.residual(u0)
    dev.jacobian(u0, download=False)
    z, rr, res = dev.solve(r, 1e-6, maxit=2000, slab_w=128, fill=0)
    assert res.converged == 1
    true_r = r - dev.spmv(z)
    assert np.linalg.norm(true_r) <= 1.01e-6 * np.linalg.norm(r)
    assert np.max(np.abs(true_r - rr)) <= 1e-9 * np.max(np.abs(r))      # the recurrence residual is the true one
    # restarted GMRES (small restart: 8 basis vectors of 540 MB): preconditioned residual reduced as reported
    z2, rr2, res2 = dev.solve_gmres(r, 1e-3, restart=8, maxit=200, slab_w=128, fill=0)
    assert res2.converged == 1 and res2.reduction < 1e-3
    p0 = dev.ilu_apply(r)
    p1 = dev.ilu_apply(r - dev.spmv(z2))
    assert np.linalg.norm(p1) <= 1.05e-3 * np.linalg.norm(p0)


def test_fullsize_newton_step_is_deterministic(full, axlib):
    dev, s = full
    u0 = s["u0"]
    opts = axlib.default_newton_opts(maxit=2, lin_maxit=10, fixed_work=1, slab_w=128)
    outs = []
    for _ in range(2):
        dev.set_time(0.0, 1.0); dev.set_uold(u0); dev.update_state(u0)
        u, res = dev.newton(u0, opts)
        outs.append((u, res.defect))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert outs[0][1] < 1e-3 * res.first_defect          # two Newton iterations reduce the defect
