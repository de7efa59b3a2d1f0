import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from oracle import pyoracle as O
from dune_ax1_b200 import api, setups
from helpers import oracle_problem, small_setup
s = small_setup(nx=5, stimulation=True)
for lin in (1e-8,):
    dev = setups.make_device(s); P = oracle_problem(O, s)
    og = api.default_newton_opts(reduction=1e-8, abs_limit=5.0, min_lin_reduction=lin, slab_w=3)
    oc = O.default_newton_opts(reduction=1e-8, abs_limit=5.0, min_lin_reduction=lin, ilu_fill=0, slab_w=3)
    ug = s["u0"].copy(); uc = s["u0"].copy(); t, dt = 19.0, 1.0
    for step in range(3):
        dev.set_time(t, dt); dev.update_state(ug); dev.set_uold(ug)
        P.set_time(t, dt); P.update_state(uc); P.set_uold(uc)
        ug2, rg = dev.newton(ug, og); uc2, rc = P.newton(uc, oc)
        print(lin, step, "gpu", rg.status, rg.iterations, rg.lin_iterations, "%.3e %.3e"%(rg.first_defect, rg.defect), "| cpu", rc.status, rc.iterations, rc.lin_iterations, "%.3e %.3e"%(rc.first_defect, rc.defect), api._err())
        ug, uc = ug2, uc2
        dev.accept_state(); P.accept_state(); t += dt
    scale = np.max(np.abs(uc.reshape(-1, 4)), axis=0)
    print("  field err", np.max(np.abs(ug - uc).reshape(-1, 4) / scale, axis=0))
    dev.close()
