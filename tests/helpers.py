"""Shared helpers of the parity tests: build the SAME problem in the oracle and on the device."""
import numpy as np

from dune_ax1_b200 import setups


def oracle_problem(O, s, analytic=1, **cfg_over):
    """analytic: 1 = all three operators with analytic jacobian_volume, 0 = all by finite differences (reference
    default), otherwise a bit mask (1 elec, 2 memb, 4 time operator analytic)."""
    p = s["params"]
    analytic = 7 if analytic == 1 else int(analytic)
    bc = s["bc"].c
    cfg = O.default_config(
        xmin_global=0.0, xmax_global=s["xmax"], ymin=float(s["y"][0]), ymax=float(s["y"][-1]),
        do_stimulation=p["do_stimulation"], tinj_start=p["tinj_start"], tinj_end=p["tinj_end"], stim_x=p["stim_x"],
        stim_y=p["stim_y"], I_inj=p["I_inj"], do_equilibration=p["do_equilibration"],
        t_equilibrium=p["t_equilibrium"], dummy_value=p["dummy_value"], analytic_volume_jacobian=analytic,
        top_dirichlet=list(bc.top), bottom_dirichlet=list(bc.bottom),
        left_cytosol_dirichlet=list(bc.left_cytosol), right_cytosol_dirichlet=list(bc.right_cytosol),
        left_debye_dirichlet=list(bc.left_debye), right_debye_dirichlet=list(bc.right_debye),
        left_extra_dirichlet=list(bc.left_extra), right_extra_dirichlet=list(bc.right_extra),
        debye_layer_width=bc.debye_layer_width,
        use_rownorm_preconditioner=p.get("use_rownorm_preconditioner", 0), membrane_active=p.get("membrane_active", 1),
        automatic_channel_init=p.get("automatic_channel_init", 1),
        channel_resting_potential=p.get("channel_resting_potential", -65.0),
        do_memb_flux_stimulation=p.get("do_memb_flux_stimulation", 0), memb_flux_ion=p.get("memb_flux_ion", 2),
        memb_flux_inj=p.get("memb_flux_inj", 0.0), **cfg_over)
    return O.Problem(cfg, s["x"], s["y"], left_pb=s["left_pb"], right_pb=s["right_pb"], eps_memb=s["eps_memb"],
                     g_leak=s["g_leak"], g_nav=s["g_nav"], g_kv=s["g_kv"])


def small_setup(nx=6, **kw):
    kw.setdefault("dx", 100.0e3)
    kw.setdefault("perturbation", 1e-3)
    return setups.make_setup(nx, **kw)


def rel_err(a, b, scale):
    return float(np.max(np.abs(a - b) / (scale + 1e-300)))


def trace_restarts(fx):
    """Replay the time-step controller (dune_ax1_b200/timeloop.py) on a recorded trace (tests/golden/ap_trace_*.json):
    returns, per step, how often dt was halved after a failed Newton solve in the recorded run (the controller's dt
    divided by the recorded dt), asserting that everything else of the dt sequence follows from the controller."""
    from dune_ax1_b200.timeloop import TimeLoop

    class Replay:
        def __init__(self, rows):
            self.rows, self.k = rows, 0

        def membrane_potential(self, u):     # the controller looks at the state before the step it chooses dt for
            return np.array([self.rows[self.k - 1][8]]) if self.k > 0 else np.array([-65.0])

    rows = fx["steps"]
    dev = Replay(rows)
    tl = TimeLoop(dev, np.zeros(4), newton_opts=None, do_equilibration=False, t_equilibrium=0.0, **fx["ctl"])
    out = []
    for k, row in enumerate(rows):
        dev.k = k
        tl._choose_dt()
        n = 0
        while tl.dt != row[1] and tl.dt > row[1] and n < 10:
            tl.dt *= 0.5
            n += 1
        assert tl.dt == row[1], (k, tl.dt, row[1])
        out.append(n)
        tl.last_its, tl.its = tl.its, row[2]
        tl.time += tl.dt
        assert abs(tl.time - row[0]) <= 1e-9 * max(1.0, row[0])
    return out
