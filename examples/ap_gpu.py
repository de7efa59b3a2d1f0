"""Full paper grid (square10mm.config: 100 x 180 cells, 73,124 DOF) action-potential run on the device:
adaptive time loop until tend, records the V_m trace at the injection site and far away."""
import sys, time, json, numpy as np, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
from dune_ax1_b200.timeloop import TimeLoop, DeviceTimeLoop
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 100
tend = float(sys.argv[2]) if len(sys.argv) > 2 else 3000.0
slab_w = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # 0: library default (16-column sub-slabs on this grid)
fill = int(sys.argv[5]) if len(sys.argv) > 5 else 0
krylov = int(sys.argv[6]) if len(sys.argv) > 6 else 0    # 1: restarted GMRES instead of BiCGStab
native = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # 1: time loop inside the library (state resident in HBM)
s = setups.make_setup(nx, dx=100e3, stimulation=True, perturbation=0.0)
dev = setups.make_device(s)
opts = api.default_newton_opts(reduction=1e-5, abs_limit=1e-5, min_lin_reduction=1e-5, maxit=50, slab_w=slab_w, ilu_fill=fill, krylov=krylov, restart=50)
if native:
    tl = DeviceTimeLoop(dev, s["u0"], dtstart=1.0, newton_opts=opts, lower_limit=10, upper_limit=30)
else:
    tl = TimeLoop(dev, s["u0"], dtstart=1.0, newton_opts=opts, do_stimulation=True, lower_limit=10, upper_limit=30)
t0 = time.time(); rows = []; nits = 0; lin = 0
while tl.time < tend:
    r = tl.step(); nits += r.iterations; lin += r.lin_iterations
    vm = dev.membrane_potential_dev() if native else dev.membrane_potential(tl.u)
    rows.append((tl.time, tl.dt, r.iterations, float(vm[1]), float(vm[nx // 2]), float(vm[-1])))
wall = time.time() - t0
print("steps %d newton its %d lin its %d wall %.1fs -> %.1f Newton steps/s" % (len(rows), nits, lin, wall, nits / wall))
peak = max(rows, key=lambda x: x[3]); print("V_m peak at injection site: %.2f mV at t=%.1f us" % (peak[3], peak[0]))
peak = max(rows, key=lambda x: x[5]); print("V_m peak at far end: %.2f mV at t=%.1f us" % (peak[5], peak[0]))
json.dump({"nx": nx, "tend": tend, "steps": len(rows), "newton_its": nits, "lin_its": lin, "wall_s": wall,
           "columns": ["t_us", "dt_us", "newton_its", "vm_x150um_mV", "vm_mid_mV", "vm_end_mV"], "trace": rows[::5]},
          open("gpurun_out/ap_trace_nx%d.json" % nx, "w"))
