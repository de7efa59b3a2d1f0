"""Oracle trace of BASELINE configs[1] -- the square10mm.config action-potential run on the full paper grid
(100 x 180 cells, 73,124 DOF, Na injection at x = 150 um for t in (20, 2000) us, HH channels, adaptive implicit
Euler from the shipped equilibrium state) -- to t = 3 ms, as the fixture tests/test_gpu_action_potential.py holds
the device's run against. Runs the CPU oracle only (minutes of one core); run it in the build container:
    python tests/golden/make_ap_trace.py            # tolerances at the rounding floor (fields to 1e-8)
    python tests/golden/make_ap_trace.py --config   # the config's own tolerances (dt sequence / Newton history)
Newton tolerances are tightened to just above the rounding floor of the residual (as in the miniature test) so that
oracle and device stop for the same reason; the preconditioner is block ILU0 on 16-column sub-slabs, the library
default on this grid.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import cpu_setup as CS, pyoracle as O            # noqa: E402
from dune_ax1_b200.timeloop import TimeLoop                  # noqa: E402  (pure Python control flow)

NEWTON = dict(reduction=1e-8, abs_limit=20.0, min_lin_reduction=1e-8, maxit=50)
# second trace ("config"): the run as the reference's square10mm.config makes it -- its Newton tolerances (newtonReduction
# 1e-5, newtonAbsLimit 1e-5, newtonMinLinReduction 1e-5), the library-default preconditioner of the device (16-column
# sub-slabs + radial cut at row 91); the device free-runs against it: same dt sequence, same Newton count in every step
NEWTON_CONFIG = dict(reduction=1e-5, abs_limit=1e-5, min_lin_reduction=1e-5, maxit=50)
CTL = dict(dtstart=1.0, do_stimulation=True, lower_limit=10, upper_limit=30)
SLAB_W, TEND, SPLIT_CONFIG = 16, 3000.0, 91
SITES = (1, 50, 99)          # membrane columns: injection site (x = 150 um), middle, far end
STRIDE = 28                  # every 28th vertex of the final state is kept


def main():
    config = "--config" in sys.argv
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tend = float(args[0]) if args else TEND
    newton = NEWTON_CONFIG if config else NEWTON
    P, u0 = CS.make_problem(100, 100.0e3, stimulation=True)
    if config:
        P.set_ilu_split(SPLIT_CONFIG)
    tl = TimeLoop(P, u0, newton_opts=O.default_newton_opts(ilu_fill=0, slab_w=SLAB_W, **newton), **CTL)
    rows, profiles = [], []
    t0 = time.time()
    lin = 0
    while tl.time < tend - 1e-8:
        r = tl.step()
        lin += r.lin_iterations
        vm = P.membrane_potential(tl.u)
        rows.append([tl.time, tl.dt, r.iterations, r.lin_iterations] + [float(vm[i]) for i in SITES] +
                    [float(vm.min()), float(vm.max())])
        if len(rows) % 20 == 0:
            profiles.append([len(rows)] + [float(v) for v in vm])
            print("step %d t=%.2f dt=%.3f its=%d vm=%s wall %.0fs" % (len(rows), tl.time, tl.dt, r.iterations,
                                                                     rows[-1][4:7], time.time() - t0), flush=True)
    u = tl.u.reshape(-1, 4)
    out = {"newton": newton, "ctl": CTL, "slab_w": SLAB_W, "ilu_split": SPLIT_CONFIG if config else 0, "tend": tend, "sites": SITES, "stride": STRIDE,
           "columns": ["t_us", "dt_us", "newton_its", "krylov_its", "vm_site0", "vm_site1", "vm_site2", "vm_min", "vm_max"],
           "steps": rows, "u_final_sample": [] if config else u[::STRIDE].tolist(),
           "u_final_absmax": np.max(np.abs(u), axis=0).tolist(), "channel_state_final": P.channel_state().tolist(),
           "wall_s": time.time() - t0}
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "ap_trace_paper_grid_config.json" if config else "ap_trace_paper_grid.json")
    with open(dst, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(dst, len(rows), "steps", sum(r[2] for r in rows), "Newton its", lin, "Krylov its", "%.0f s" % out["wall_s"])


if __name__ == "__main__":
    main()
