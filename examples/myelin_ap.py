"""BASELINE config 5 on the device: the 48 mm myelinated axon (48 nodes of Ranvier), adaptive time loop with
the config's own Newton/linear tolerances; records V_m at a few nodes (saltatory conduction) and the
Newton-steps/s of the converged run. usage: myelin_ap.py [tend_us] [ilu_fill] [slab_w] [max_steps]"""
import sys, time, json, numpy as np, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
from dune_ax1_b200.timeloop import TimeLoop
tend = float(sys.argv[1]) if len(sys.argv) > 1 else 500.0
fill = int(sys.argv[2]) if len(sys.argv) > 2 else 1
slab_w = int(sys.argv[3]) if len(sys.argv) > 3 else 128
max_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10 ** 9
amg = int(sys.argv[5]) if len(sys.argv) > 5 else 0
krylov = int(sys.argv[6]) if len(sys.argv) > 6 else 0
cx = int(sys.argv[7]) if len(sys.argv) > 7 else 4
s = setups.make_myelin_setup()
dev = setups.make_device(s)
# newtonReduction 5e-6, newtonAbsLimit 100, newtonMinLinReduction 1e-2, linearSolverMaxIt 1000, newtonMaxIt 50
opts = api.default_newton_opts(reduction=5e-6, abs_limit=100.0, min_lin_reduction=1e-2, maxit=50, lin_maxit=1000,
                               slab_w=slab_w, ilu_fill=fill, amg=amg, krylov=krylov, restart=100)
if amg:
    dev.set_amg(True, cx)
# maxTimeStep = maxTimeStepAP = 0.01e-3 s, minTimeStep = 0.05e-6 s, Newton-iteration limits 10 / 30, 2 restarts
tl = TimeLoop(dev, s["u0"], dtstart=1.0, newton_opts=opts, do_stimulation=True, tinj_start=20.0, tinj_end=3.0e3,
              max_dt=10.0, min_dt=0.05, max_dt_ap=10.0, lower_limit=10, upper_limit=30, max_restarts=2)
node = s["groups"].names.index("node_of_ranvier")
cols = np.nonzero(s["cell_group"] == node)[0]
node_cols = [int(cols[5 + 10 * k]) for k in (0, 1, 2, 5, 10, 20)]   # centre column of nodes 0, 1, 2, 5, 10, 20
t0 = time.time(); rows = []; nits = 0; lin = 0
while tl.time < tend and len(rows) < max_steps:
    r = tl.step(); nits += r.iterations; lin += r.lin_iterations
    vm = dev.membrane_potential(tl.u)
    rows.append([tl.time, tl.dt, r.iterations, r.lin_iterations] + [float(vm[c]) for c in node_cols])
    if len(rows) % 20 == 0:
        print("t=%.1f dt=%.2f its=%d lin=%d vm=%s" % (tl.time, tl.dt, r.iterations, r.lin_iterations,
              ["%.1f" % v for v in rows[-1][4:]]), flush=True)
wall = time.time() - t0
print("steps %d newton its %d lin its %d wall %.1fs -> %.1f Newton steps/s, %.1f BiCGStab its / Newton step"
      % (len(rows), nits, lin, wall, nits / wall, lin / max(nits, 1)))
json.dump({"workload": "myelin 48mm x 10mm, 48 nodes, 7251 x 180 cells", "tend": tl.time, "steps": len(rows),
           "ilu_fill": fill, "slab_w": slab_w, "amg": amg, "amg_cx": cx, "krylov": krylov, "newton_its": nits, "lin_its": lin, "wall_s": wall,
           "newton_steps_per_s": nits / wall,
           "columns": ["t_us", "dt_us", "newton_its", "lin_its"] + ["vm_node%d_mV" % k for k in (0, 1, 2, 5, 10, 20)],
           "trace": rows}, open("gpurun_out/myelin_ap_fill%d_w%d_amg%d_k%d_cx%d.json" % (fill, slab_w, amg, krylov, cx), "w"))
dev.close()
