"""Electroneutral (Mori) operator-split model, SURVEY 8(f)4 / BASELINE configs[2]: device (csrc/mori.cu) against the
oracle's restatement (oracle/mori.cpp) of acme2_cyl_mori_{potential,concentration}_operator.hh, the Mori parameter
classes, ChargeLayerMoriImplicit and the step order of Acme2CylMoriSimulation::run."""
import numpy as np
import pytest

from helpers import oracle_problem, small_setup

pytestmark = pytest.mark.gpu


def _pair(oracle, axlib, s, t=0.0, dt=1.0):
    from dune_ax1_b200 import setups
    dev = setups.make_device(s)
    P = oracle_problem(oracle, s)
    return dev, P


def _bcrs_scale(P, vals, u):
    """row scale sum_k |J_ik| |u_k| of a BCRS matrix (what the rounding error of a residual row is measured against)"""
    rowptr, col = P.pattern()
    blocks = np.abs(vals.reshape(-1, 4, 4))
    ua = np.abs(u.reshape(-1, 4))
    out = np.zeros((P.nv, 4))
    for i in range(P.nv):
        for k in range(rowptr[i], rowptr[i + 1]):
            out[i] += blocks[k] @ ua[col[k]]
    return out.reshape(-1)


def _sensitivity_scale(P, which, u, ulast, uold, r):
    """|d r_i| per unit relative perturbation of the three input vectors: the magnitude of the terms a residual row
    adds up before they cancel (diffusion currents of the three species, channel vs capacitive flux, ...), i.e. what
    a rounding error has to be measured against"""
    rng = np.random.default_rng(3)
    eps, scale = 1e-9, np.zeros_like(r)
    for _ in range(3):
        pu, pl, po = (v * (1.0 + eps * rng.standard_normal(v.size)) for v in (u, ulast, uold))
        P.set_uold(po)
        rp, _ = P.mori_assemble(which, pu, pl, jacobian=False)
        scale = np.maximum(scale, np.abs(rp - r) / eps)
    P.set_uold(uold)
    return scale


@pytest.mark.parametrize("setup", [dict(nx=6), dict(nx=5, stimulation=True), dict(nx=7, equilibration=True)])
def test_mori_subproblem_assembly_parity(oracle, axlib, setup):
    """Residual and Jacobian of both linear sub-problems with three different vectors in the three roles (current
    iterate, last iteration, previous time step), charge layer initialised and advanced once: every entry to 1e-11 of
    the magnitude of the terms its row adds up, the constrained rows and the identity rows of the other sub-problem exactly."""
    s = small_setup(**setup)
    dev, P = _pair(oracle, axlib, s)
    u0 = s["u0"]
    rng = np.random.default_rng(11)
    uold = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    ulast = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    u = u0 * (1.0 + 1e-3 * rng.standard_normal(u0.size))
    t = 30.0 if setup.get("stimulation") else 0.0          # inside the injection window
    for X in (dev, P):
        X.set_time(t, 0.7); X.set_uold(u0); X.update_state(u0); X.accept_state()
        X.mori_update_state(u0); X.accept_state()             # charge layer: init + accept
        X.set_time(t + 0.7, 0.7); X.set_uold(uold); X.update_state(ulast)
        X.mori_update_state(ulast)                            # gamma_new != gamma_old
    assert np.max(np.abs(dev.mori_gamma() - P.mori_gamma())) < 1e-14
    assert np.allclose(P.mori_gamma().sum(axis=2), 1.0, atol=1e-12) or setup.get("equilibration")
    for which in (0, 1):
        rg, Jg = dev.mori_assemble(which, u, ulast)
        rc, Jc = P.mori_assemble(which, u, ulast)
        scale = _sensitivity_scale(P, which, u, ulast, uold, rc) + _bcrs_scale(P, Jc, u) + np.abs(rc) + 1e-300
        assert np.max(np.abs(rg - rc) / scale) < 1e-12, which
        rowscale = np.repeat(np.max(np.abs(Jc.reshape(-1, 4, 4)), axis=2).reshape(-1), 4).reshape(-1, 4, 4)
        # per scalar row: compare against the largest entry of that row of the block row
        rowptr, col = P.pattern()
        Jg3, Jc3 = Jg.reshape(-1, 4, 4), Jc.reshape(-1, 4, 4)
        for i in range(P.nv):
            blk = slice(rowptr[i], rowptr[i + 1])
            rs = np.max(np.abs(Jc3[blk]), axis=(0, 2))            # per scalar row of the block row
            assert np.all(np.abs(Jg3[blk] - Jc3[blk]) <= 1e-11 * rs[None, :, None] + 0.0), (which, i)
        # structure: the other sub-problem's rows are identity, and nothing couples the two
        rows = [0, 1, 2] if which == 0 else [3]
        diag = np.array([k for i in range(P.nv) for k in range(rowptr[i], rowptr[i + 1]) if col[k] == i])
        for a in rows:
            assert np.all(Jg3[diag, a, a] == 1.0)
            assert np.all(rg.reshape(-1, 4)[:, a] == 0.0)
    dev.close()


def test_mori_time_steps_match_oracle(oracle, axlib):
    """A run of operator-split steps through the C ABI the way Acme2CylMoriSimulation::run drives them (channel
    update, potential solve, charge-layer update, concentration solve, accept), with the Na injection switching on:
    sub-problem defects, Krylov iteration counts, fields, charge-layer and channel states against the oracle."""
    s = small_setup(nx=10, stimulation=True, perturbation=1e-3, tinj_start=3.0)
    dev, P = _pair(oracle, axlib, s)
    kw = dict(reduction=1e-10, abs_limit=1e-12, lin_maxit=500)
    og = axlib.default_newton_opts(slab_w=4, ilu_fill=0, **kw)
    oc = oracle.default_newton_opts(slab_w=4, ilu_fill=0, **kw)
    ug = uc = s["u0"]
    t, dt = 0.0, 1.0
    for step in range(8):
        for X, u in ((dev, ug), (P, uc)):
            X.set_time(t, dt); X.set_uold(u); X.update_state(u)
        ug, rg = dev.mori_step(ug, og)
        uc, rc = P.mori_step(uc, oc)
        dev.accept_state(); P.accept_state()
        assert rg.status == 0 and rc.status == 0
        assert abs(rg.pot_iterations - rc.pot_iterations) <= 1 and abs(rg.con_iterations - rc.con_iterations) <= 1
        assert abs(rg.pot_first_defect - rc.pot_first_defect) <= 1e-9 * rc.pot_first_defect
        assert abs(rg.con_first_defect - rc.con_first_defect) <= 1e-9 * rc.con_first_defect
        t += dt
    scale = np.max(np.abs(uc.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(ug - uc).reshape(-1, 4) / scale, axis=0)
    assert np.all(err < 1e-9), err
    assert np.max(np.abs(dev.mori_gamma() - P.mori_gamma())) < 1e-12
    sg, _ = dev.channel_state()
    assert np.max(np.abs(sg - P.channel_state())) < 1e-10
    # the injection changed the state: the step is not a no-op
    assert np.max(np.abs(uc - s["u0"]).reshape(-1, 4)[:, 0]) > 1e-6
    dev.close()
