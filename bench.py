#!/usr/bin/env python
"""bench.py -- PNP Newton hot path throughput (BASELINE.json metric) on N B200 GPUs of one node.

Workload (SURVEY 8(d) "synthetic weak scaling"): the paper's r-z tensor grid refined axially --
the shipped 181-point radial vector (Ny = 180) x Nx = 93,184 cell columns of dx = 100 nm PER GPU
(16.77 M cells, 16.87 M vertices, 67.5 M DOF per GPU), HH channels everywhere, implicit Euler
dt = 1 us, deterministic perturbation of the replicated equilibrium column. One bench "step" is one
time step = gating update + 5 Newton iterations, each = Jacobian assembly + block-ILU0
factorisation + 20 BiCGStab iterations + residual/line-search evaluation (fixed work, lambda = 1).

  value  GDOF/s = 4 * vertices * Newton iterations / time over all GPUs (Newton steps/s is reported
         beside it), state resident in HBM, timed with CUDA events on the library's stream
  e2e    the same step through the host-buffer C ABI call ax1gpu_newton (pinned host buffers,
         H2D of uold and u and D2H of u inside the timed region)
  --impl reference: the CPU restatement of the reference algorithm (oracle/, all host cores as
         x-slabs with one overlap column, BiCGStab + block ILU(1), analytic volume Jacobian) on a
         bounded sample of the same workload. The real DUNE binary cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1, so fd 1 is
# pointed at stderr for the whole run and the result line goes to the saved original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


METRIC = "PNP assembly+solve throughput (DOF x Newton steps per second)"
UNIT = "GDOF/s"

NX_PER_GPU = 93184          # 728 * 128 cell columns per GPU
DX = 100.0                  # nm
NEWTON_ITS = 5
KRYLOV_ITS = 20
SLAB_W = 128
# algorithmic bytes per vertex. SURVEY 8(d) table for dense 4x4 blocks: jacobian 1184, spmv 1256,
# ilu_factor 2384. The device stores the Jacobian blocks in their 10-entry "arrow" shape (80 B instead
# of 128 B, common.cuh), so the bytes these kernels must move are smaller; roofline.achieved uses the
# bytes of the format actually stored, the dense-equivalent figure is reported beside it.
BYTES = {"residual": 104, "jacobian": 32 + 720, "spmv": 720 + 32 + 32, "ilu_factor": 720 + 1152, "ilu_apply": 1384,
         "dot": 64, "axpy": 224}
BYTES_DENSE = {"residual": 104, "jacobian": 1184, "spmv": 1256, "ilu_factor": 2384, "ilu_apply": 1384,
               "dot": 64, "axpy": 224}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(nx_sample, threads, steps, warmup):
    """Time the oracle (reference algorithm restated) on a bounded sample: `threads` x-slabs."""
    from oracle import pyoracle as O
    from dune_ax1_b200 import setups
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_problem
    O.build()
    probs, us = [], []
    for rk in range(threads):
        s = setups.make_setup(nx_sample, dx=DX, rank=rk, nranks=threads, perturbation=1e-3)
        P = oracle_problem(O, s)
        P.set_time(0.0, 1.0); P.set_uold(s["u0"]); P.update_state(s["u0"]); P.accept_state()
        P.set_time(1.0, 1.0); P.set_uold(s["u0"])
        probs.append(P); us.append(s["u0"])
    nv_global = (nx_sample + 1) * probs[0].ny + (nx_sample + 1)
    opts = O.default_newton_opts(maxit=NEWTON_ITS, lin_maxit=KRYLOV_ITS, fixed_work=1, ilu_fill=1, slab_w=0)
    times = []
    res = None
    for it in range(warmup + steps):
        for P, u in zip(probs, us):
            P.update_state(u)
        t0 = time.perf_counter()
        _, res, _ = O.newton_mt(probs, us, opts)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return dict(sec_per_step=t, nv=nv_global, newton_its=NEWTON_ITS, threads=threads,
                t_assembler=res[0].t_assembler, t_lin_solv=res[0].t_lin_solv)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=NX_PER_GPU, help="cell columns per GPU (default = named workload)")
    ap.add_argument("--slab-w", type=int, default=SLAB_W)
    ap.add_argument("--cpu-nx", type=int, default=2048, help="cell columns of the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "myelin"],
                    help="synthetic: weak scaling, Nx columns PER GPU (headline); myelin: BASELINE config 5, the 48 mm "
                         "myelinated axon with 48 nodes of Ranvier (7,251 x 180 cells), STRONG scaling over z-slabs")
    ap.add_argument("--ilu-fill", type=int, default=0, choices=[0, 1])
    args = ap.parse_args()
    myelin = args.workload == "myelin"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if myelin and args.slab_w == SLAB_W:
        args.slab_w = 24 if args.ilu_fill == 1 else 16   # the library default on this grid: one-warp sub-slabs

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    nv_full = (args.nx + 1) * 181
    workload = ("synthetic refined r-z tensor grid: Ny=180 (shipped 181-point radial vector) x Nx=%d columns/GPU, "
                "dx=100 nm, HH channels, implicit Euler dt=1us, %d Newton x %d BiCGStab(ILU%d slab %d) per step"
                % (args.nx, NEWTON_ITS, KRYLOV_ITS, args.ilu_fill, args.slab_w))
    if myelin:
        workload = ("myelinated axon 48 mm x 10 mm, 48 nodes of Ranvier (acme2_cyl_par_fiona_myelin_48mmx10mm_48nodes "
                    "config): 7,251 x 180 cells in total, non-uniform x (100 nm in the nodes ... 10 um in the myelin), "
                    "implicit Euler dt=1us, %d Newton x %d BiCGStab(ILU%d slab %d) per step"
                    % (NEWTON_ITS, KRYLOV_ITS, args.ilu_fill, args.slab_w))

    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        steps = max(1, min(args.steps, 2))
        r = cpu_reference(args.cpu_nx, threads, steps, 1 if args.warmup > 0 else 0)
        # GDOF/s = 4 * vertices * Newton iterations / time is size independent (work is linear in the
        # number of vertices), so the bounded sample needs no extrapolation
        gdof = 4.0 * r["nv"] * NEWTON_ITS / r["sec_per_step"] / 1e9
        full_nv = nv_full * max(1, args.gpus)
        sample = ("CPU restatement of the reference algorithm (oracle/; not the DUNE binary), %d threads as %d "
                  "x-slabs with 1 overlap column, BiCGStab+block ILU(1), analytic volume Jacobian; sample Nx=%d "
                  "(%d vertices, %.2f s per step of %d Newton x %d BiCGStab iterations)"
                  % (threads, threads, args.cpu_nx, r["nv"], r["sec_per_step"], NEWTON_ITS, KRYLOV_ITS))
        out = {"impl": "reference", "metric": METRIC, "value": gdof, "unit": UNIT,
               "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload},
               "newton_steps_per_s_at_workload_size": NEWTON_ITS / r["sec_per_step"] * (r["nv"] / full_nv),
               "cpu_baseline": {"value": gdof, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
               "e2e": {"value": gdof, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return

    import torch
    import torch.distributed as dist
    from dune_ax1_b200 import api, setups
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "dune_ax1_b200", "libax1gpu.so")):
        ge.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)

    if myelin:
        s = setups.make_myelin_setup(rank=rank, nranks=world)
    else:
        s = setups.make_setup(args.nx * world, dx=DX, rank=rank, nranks=world, perturbation=1e-3)
    stream = torch.cuda.Stream(device=local_rank)
    dev = setups.make_device(s, device=local_rank, stream=stream.cuda_stream)
    if world > 1:
        ids = [api.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        dev.comm_init(ids[0], rank, world)
    nv = dev.nv
    opts = api.default_newton_opts(maxit=NEWTON_ITS, lin_maxit=KRYLOV_ITS, fixed_work=1, slab_w=args.slab_w,
                                   ilu_fill=args.ilu_fill)

    # pinned host buffers for the e2e leg
    u_host_t = torch.empty(dev.n, dtype=torch.float64, pin_memory=True)
    u_host = u_host_t.numpy()
    u_host[:] = s["u0"]
    uold_host_t = torch.empty(dev.n, dtype=torch.float64, pin_memory=True)
    uold_host = uold_host_t.numpy()
    uold_host[:] = s["u0"]

    # initialise channels (first time step of the reference does this) and accept
    dev.set_time(0.0, 1.0)
    dev.upload_u(u_host)
    dev.set_uold_dev()
    dev.update_state()
    dev.accept_state()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step_resident(t):
        dev.set_time(t, 1.0)
        dev.update_state()          # gfMembFlux.updateState(time, dt) of the time loop
        dev.set_uold_dev()          # r0 = -M(uold), uold = current device iterate
        return dev.newton_dev(opts)

    def one_step_e2e(t):
        import ctypes as C
        dev.set_time(t, 1.0)
        api._chk(api.lib().ax1gpu_update_state(dev.h, api._dp(u_host)))     # H2D u
        api._chk(api.lib().ax1gpu_set_uold(dev.h, api._dp(uold_host)))      # H2D uold
        res = api.NewtonResult()
        api._chk(api.lib().ax1gpu_newton(dev.h, C.byref(opts), api._dp(u_host), C.byref(res)))  # H2D u, D2H u
        return res

    t_sim = 1.0
    for _ in range(args.warmup):
        dev.upload_u(u_host)
        one_step_resident(t_sim)
    # ---- timed region: resident state ----
    dev.upload_u(u_host)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        last = None
        for k in range(args.steps):
            last = one_step_resident(t_sim)
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = dev.launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    # ---- e2e: host buffers through the C ABI ----
    u_host[:] = s["u0"]
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for k in range(args.steps):
            one_step_e2e(t_sim)
        ev1.record(stream)
    barrier()
    ms_e2e = ev0.elapsed_time(ev1)
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
    its = NEWTON_ITS * args.steps
    nv_total = nv_full * world
    if myelin:   # strong scaling: the global grid is fixed
        nv_total = len(s["xg"]) * 181
    newton_per_s = its / (ms * 1e-3)
    value = 4.0 * nv_total * its / (ms * 1e-3) / 1e9          # whole-job GDOF/s
    e2e_value = 4.0 * nv_total * its / (ms_e2e * 1e-3) / 1e9

    # ---- roofline of the dominant kernel, measured live with CUDA events on the library stream ----
    peak, peak_kind = measured_peak()
    # launches per step: residual = 1 + 2 per Newton iteration (line search + defect), dot = norms only
    # (the Krylov dot products are fused into the SpMV and the x,r updates), axpy = the fused x,r update
    counts = {"residual": 2 * NEWTON_ITS + 1, "jacobian": NEWTON_ITS, "ilu_factor": NEWTON_ITS,
              "spmv": 2 * KRYLOV_ITS * NEWTON_ITS, "ilu_apply": 2 * KRYLOV_ITS * NEWTON_ITS,
              "dot": 3 * NEWTON_ITS + 1, "axpy": 2 * KRYLOV_ITS * NEWTON_ITS}
    kern = {}
    bytes_k, bytes_dense = dict(BYTES), dict(BYTES_DENSE)
    if args.ilu_fill == 1:   # 13 instead of 9 blocks per factor row (ilu1.cu)
        bytes_k["ilu_factor"], bytes_k["ilu_apply"] = 720 + 1664, 1664 + 64
        bytes_dense["ilu_factor"], bytes_dense["ilu_apply"] = 1192 + 1664, 1664 + 64
    for name in counts:
        t_ms = dev.time_kernel(name, reps=3)
        kern[name] = {"ms": t_ms, "launches_per_step": counts[name],
                      "GBps": bytes_k[name] * nv / (t_ms * 1e-3) / 1e9,
                      "GBps_dense_equivalent": bytes_dense[name] * nv / (t_ms * 1e-3) / 1e9,
                      "share_of_step": counts[name] * t_ms / (ms / args.steps)}
    dom = max(kern, key=lambda k: kern[k]["share_of_step"])
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (named workload only)
    traffic = None
    if args.nx == NX_PER_GPU and args.slab_w == SLAB_W and not myelin and args.ilu_fill == 0:
        traffic = {"ilu_apply": 20.352319e9 + 1.073509e9}.get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "traffic": traffic,
                "traffic_source": "profiles/r01_top_kernels_ncu.json", "peak_kind": peak_kind,
                "algorithmic_bytes_per_launch": bytes_k[dom] * nv, "kernels": kern}
    # whole-step algorithmic traffic (SURVEY 8(d) formula, l = 1 line-search residual, k Krylov its)
    step_bytes = NEWTON_ITS * (bytes_k["jacobian"] + bytes_k["ilu_factor"] + 2 * 104 +
                               KRYLOV_ITS * (2 * bytes_k["spmv"] + 2 * bytes_k["ilu_apply"] + 416)) * nv
    step_bytes_dense = NEWTON_ITS * (1184 + bytes_dense["ilu_factor"] + 2 * 104 +
                                     KRYLOV_ITS * (2 * 1256 + 2 * bytes_dense["ilu_apply"] + 416)) * nv
    roofline["step_frac_of_peak"] = step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak
    roofline["step_frac_of_peak_dense_equivalent"] = step_bytes_dense / (ms / args.steps * 1e-3) / 1e9 / peak

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "strong" if myelin else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload, "cells_per_gpu": s["nx"] * 180, "dof_total": 4 * nv_total,
                      "l2": ("working set %.1f GB per GPU" % (nv * (720 + 1664 + 12 * 32) / 1e9)) if myelin else
                            "inputs larger than L2 (working set ~46 GB per GPU)",
                      "parallelism": "z-slab x%d" % world, "comm": {0: "none", 1: "nccl", 2: "nvlink peer windows"}[dev.comm_mode()]},
           "newton_steps_per_s": newton_per_s, "newton_iterations_per_step": NEWTON_ITS,
           "krylov_iterations_per_solve": KRYLOV_ITS,
           "last_step": {"defect": last.defect, "first_defect": last.first_defect, "lin_iterations": last.lin_iterations,
                         "t_assembler": last.t_assembler, "t_lin_solv": last.t_lin_solv},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * dev.n * 8,
                   "d2h_bytes_per_step": dev.n * 8, "ms_per_step": ms_e2e / args.steps},
           "roofline": roofline}
    if rank == 0 and world == 1 and not myelin:
        # secondary figure: the paper configuration (square10mm.config, 100 x 180 cells, 73,124 DOF), first 60
        # adaptive time steps of the stimulated run through the restated time loop (latency-bound regime)
        try:
            from dune_ax1_b200.timeloop import TimeLoop
            sp = setups.make_setup(100, dx=100.0e3, stimulation=True, perturbation=0.0)
            devp = setups.make_device(sp, device=local_rank)
            tl = TimeLoop(devp, sp["u0"], dtstart=1.0, do_stimulation=True, lower_limit=10, upper_limit=30,
                          newton_opts=api.default_newton_opts(slab_w=0))  # 0: library default width
            t0 = time.perf_counter()
            nits = sum(tl.step().iterations for _ in range(60))
            wallp = time.perf_counter() - t0
            out["paper_config"] = {"workload": "square10mm.config grid, 73,124 DOF, 60 adaptive time steps from t=0 "
                                               "(Na injection starts at 20 us), reference Newton tolerances 1e-5, library-default sub-slab width",
                                   "newton_steps_per_s": nits / wallp, "newton_iterations": nits, "t_end_us": tl.time,
                                   "timing": "host wall clock incl. H2D/D2H of every step"}
            devp.close()
        except Exception as e:  # never lose the headline line because of the secondary figure
            out["paper_config"] = {"error": str(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not myelin:
        threads = os.cpu_count() or 1
        r = cpu_reference(args.cpu_nx, threads, 1, 0)
        out["cpu_baseline"] = {
            "value": 4.0 * r["nv"] * NEWTON_ITS / r["sec_per_step"] / 1e9, "unit": UNIT, "cores": threads,
            "kind": "port",
            "sample": "oracle/ restatement of the reference algorithm (not the DUNE binary), %d threads as x-slabs with 1 "
                      "overlap column, BiCGStab+block ILU(1), analytic volume Jacobian, Nx=%d sample (%d vertices, %.1f s "
                      "per step of %d Newton x %d BiCGStab iterations)" % (threads, args.cpu_nx, r["nv"],
                                                                        r["sec_per_step"], NEWTON_ITS, KRYLOV_ITS),
            "newton_steps_per_s_at_workload_size": NEWTON_ITS / r["sec_per_step"] * (r["nv"] / nv_total)}
    if rank == 0:
        emit(out)
    dev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
