"""BASELINE config 2 in miniature: the stimulated action-potential run of square10mm.config (Na injection
at x = 150 um, adaptive time step, HH channels), driven by the restated time loop of
Acme2CylSimulation::run on the device and on the oracle. Membrane-potential traces must agree."""
import numpy as np
import pytest

from helpers import oracle_problem

pytestmark = pytest.mark.gpu


def test_action_potential_upstroke_matches_oracle(oracle, axlib):
    from dune_ax1_b200 import setups
    from dune_ax1_b200.timeloop import TimeLoop
    nx, tend = 12, 330.0
    s = setups.make_setup(nx, dx=100.0e3, stimulation=True, perturbation=0.0)
    dev = setups.make_device(s)
    P = oracle_problem(oracle, s)
    # same preconditioner on both sides (one sub-slab = global ILU0), tolerances tightened to just above
    # the rounding floor of the residual so that both Newton solves stop for the same reason
    kw = dict(reduction=1e-8, abs_limit=20.0, min_lin_reduction=1e-8, maxit=50)
    og = axlib.default_newton_opts(slab_w=128, **kw)
    oc = oracle.default_newton_opts(ilu_fill=0, slab_w=0, **kw)
    ctl = dict(dtstart=1.0, do_stimulation=True, lower_limit=10, upper_limit=30)   # square10mm.config
    tg = TimeLoop(dev, s["u0"], newton_opts=og, **ctl)
    tc = TimeLoop(P, s["u0"], newton_opts=oc, **ctl)
    while tc.time < tend:
        tg.step()
        tc.step()
        assert abs(tg.time - tc.time) < 1e-12 and tg.trace[-1][2] == tc.trace[-1][2], (tg.trace[-1], tc.trace[-1])
    vg, vc = dev.membrane_potential(tg.u), P.membrane_potential(tc.u)
    assert vc.max() > 0.0 and vc.min() < -60.0          # upstroke at the injection site, rest far away
    assert np.max(np.abs(vg - vc)) < 1e-5                # mV, i.e. < 1e-7 of the 100 mV excursion
    scale = np.max(np.abs(tc.u.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(tg.u - tc.u).reshape(-1, 4) / scale, axis=0)
    assert np.all(err[:3] < 1e-8) and err[3] < 5e-7, err
    sg, _ = dev.channel_state()
    assert np.max(np.abs(sg - P.channel_state())) < 1e-8  # gating variables m, h, n and leak ratios
    dev.close()
