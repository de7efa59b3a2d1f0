"""Where a Newton iteration of the paper configuration (square10mm.config, 73,124 DOF) spends its time: wall clock of
the library's time loop against the device time of assembly / linear solve (CUDA events inside ax1gpu_newton) and the
isolated sweep / factorisation kernels. Usage: python examples/paper_breakdown.py [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dune_ax1_b200 import api, setups
from dune_ax1_b200.timeloop import DeviceTimeLoop
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
s = setups.make_setup(100, dx=100e3, stimulation=True, perturbation=0.0)
out = {}
for name, split in (("default (far-field cut)", -1), ("no radial split", 0)):
    dev = setups.make_device(s)
    dev.set_ilu_split(split)
    tl = DeviceTimeLoop(dev, s["u0"], dtstart=1.0, newton_opts=api.default_newton_opts(slab_w=0), lower_limit=10, upper_limit=30)
    for _ in range(5):
        tl.step()
    dev.close()
    dev = setups.make_device(s)
    dev.set_ilu_split(split)
    tl = DeviceTimeLoop(dev, s["u0"], dtstart=1.0, newton_opts=api.default_newton_opts(slab_w=0), lower_limit=10, upper_limit=30)
    l0 = dev.launch_count(); t0 = time.perf_counter(); nits = lin = 0; ta = tls = 0.0
    for _ in range(steps):
        r = tl.step(); nits += r.newton_iterations; lin += r.lin_iterations; ta += r.t_assembler; tls += r.t_lin_solv
    wall = time.perf_counter() - t0
    out[name] = dict(newton=nits, krylov=lin, wall_ms=wall * 1e3, newton_per_s=nits / wall, launches=dev.launch_count() - l0,
                     ms_per_newton=wall * 1e3 / nits, t_assembler_ms_per_newton=ta * 1e3 / nits, t_lin_solv_ms_per_newton=tls * 1e3 / nits,
                     split_row=dev.ilu_split(), sweep_ms=dev.time_kernel("ilu_apply", 20), factor_ms=dev.time_kernel("ilu_factor", 20),
                     residual_ms=dev.time_kernel("residual", 20), jacobian_ms=dev.time_kernel("jacobian", 20),
                     spmv_ms=dev.time_kernel("spmv_dot1", 20))
    dev.close()
print(json.dumps(out, indent=1))
