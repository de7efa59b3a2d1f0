import sys, os, numpy as np
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
from dune_ax1_b200 import api, setups
for nx, dx in ((100, 100e3), (7251, 100.0), (93184, 100.0)):
    for fd in (0, 1):
        s = setups.make_setup(nx, dx=dx, perturbation=1e-3)
        if fd:
            s["params"] = dict(s["params"], use_elec_jacobian_volume=0, use_memb_jacobian_volume=0, use_time_jacobian_volume=0)
        dev = setups.make_device(s)
        dev.set_time(0.0, 1.0); dev.upload_u(s["u0"]); dev.set_uold_dev(); dev.update_state(); dev.accept_state()
        dev.jacobian(None, download=False)
        import time, torch
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): dev.jacobian(None, download=False)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print("nx %d fd=%d drv_jacobian %.4f ms (wall, 5 calls)" % (nx, fd, (t1 - t0) / 5 * 1e3), flush=True)
        dev.close()
