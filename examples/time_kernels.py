"""Isolated device times of the named kernels at the headline size (or argv[1] columns): python examples/time_kernels.py [nx] [names...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 93184
names = sys.argv[2:] or ["spmv", "spmv_dot1", "spmv_dot2", "ilu_apply", "ilu_factor", "jacobian", "residual", "axpy", "pupdate", "dot"]
s = setups.make_setup(nx, dx=100.0, perturbation=1e-3)
dev = setups.make_device(s)
dev.set_time(0.0, 1.0); dev.upload_u(s["u0"]); dev.set_uold_dev(); dev.update_state(); dev.accept_state()
dev.newton_dev(api.default_newton_opts(maxit=1, lin_maxit=2, fixed_work=1, slab_w=128))
print(json.dumps({n: round(dev.time_kernel(n, 10), 4) for n in names}))
