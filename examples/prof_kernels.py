"""Profiling driver: full-size synthetic problem, one Jacobian + ILU0 factor + a few applies/SpMVs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 93184
s = setups.make_setup(nx, dx=100.0, perturbation=1e-3)
dev = setups.make_device(s)
dev.set_time(0.0, 1.0); dev.upload_u(s["u0"]); dev.set_uold_dev(); dev.update_state(); dev.accept_state()
dev.set_time(1.0, 1.0); dev.update_state(); dev.set_uold_dev()
dev.residual(None, want_r=False)
dev.jacobian(None, download=False)
sw = int(sys.argv[2]) if len(sys.argv) > 2 else 128
fill = int(os.environ.get("AX1_FILL", "0"))
dev.ilu_factor(sw, fill)
for name in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("ilu_apply", "spmv", "spmv_dot1", "spmv_dot2", "residual", "jacobian", "ilu_factor", "dot", "axpy", "pupdate", "compact")):
    print(name, dev.time_kernel(name, reps=2))
dev.close()
