"""Linear-solver comparison on the myelin configuration (first Newton step of the stimulated run):
BiCGStab / GMRES with ILU0, ILU(1) or the multigrid cycle. usage: myelin_lin.py [reduction]"""
import sys, time, json, numpy as np, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
red = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-2
s = setups.make_myelin_setup()
dev = setups.make_device(s)
u0 = s["u0"]
dev.set_time(0.0, 1.0); dev.set_uold(u0); dev.update_state(u0)
r, nrm = dev.residual(u0)
dev.jacobian(u0, download=False)
out = []
def run(name, amg, cx, fill, slab_w, krylov):
    dev.set_amg(bool(amg), cx if amg else 0)
    dev.ilu_factor(slab_w, fill)
    t0 = time.time()
    if krylov:
        z, rr, res = dev.solve_gmres(r, red, restart=100, maxit=2000)
    else:
        z, rr, res = dev.solve(r, red, maxit=2000)
    wall = time.time() - t0
    print("%-28s conv %d its %4d red %.2e wall %.3fs (%.2f ms/it)" % (name, res.converged, res.iterations, res.reduction,
          wall, 1e3 * wall / max(res.iterations, 1)), flush=True)
    out.append(dict(name=name, converged=res.converged, iterations=res.iterations, wall_s=wall))
for red_ in (red,):
    run("bicgstab ilu0 w32", 0, 0, 0, 32, 0)
    run("bicgstab ilu0 w128", 0, 0, 0, 128, 0)
    run("bicgstab ilu1 w32", 0, 0, 1, 32, 0)
    for cx in (4, 8, 16, 32):
        run("bicgstab amg cx%d ilu0 w32" % cx, 1, cx, 0, 32, 0)
    run("bicgstab amg cx8 ilu0 w128", 1, 8, 0, 128, 0)
    run("gmres amg cx8 ilu0 w32", 1, 8, 0, 32, 1)
    run("gmres ilu0 w32", 0, 0, 0, 32, 1)
json.dump(out, open("gpurun_out/myelin_lin.json", "w"))
dev.close()
