"""A/B timing of the sweep kernels at the named full-size workload (or any nx): time_apply.py [nx] [slab_w] [fill] [split]
(split: radial cut row of the ILU0 sub-slabs, e.g. 91 on the paper grid; default none)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 93184
slab_w = int(sys.argv[2]) if len(sys.argv) > 2 else 128
fill = int(sys.argv[3]) if len(sys.argv) > 3 else 0
s = setups.make_setup(nx, dx=100.0, perturbation=1e-3)
dev = setups.make_device(s)
dev.set_time(0.0, 1.0); dev.upload_u(s["u0"]); dev.set_uold_dev(); dev.update_state(); dev.accept_state()
dev.jacobian(None, download=False)
if len(sys.argv) > 4:
    dev.set_ilu_split(int(sys.argv[4]))
dev.ilu_factor(slab_w, fill)
for rep in range(3):
    print("nx %d slab %d fill %d TMA=%s: ilu_apply %.4f ms  ilu_factor %.4f ms spmv %.4f ms" % (nx, slab_w, fill, os.environ.get("AX1_ILU_TMA", "1"),
          dev.time_kernel("ilu_apply", reps=20), dev.time_kernel("ilu_factor", reps=5), dev.time_kernel("spmv", reps=20)), flush=True)
dev.close()
