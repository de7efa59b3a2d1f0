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
  parity the device against the CPU oracle on a bounded sample of THIS workload (same dx, perturbation,
         sub-slab width and solver options), run before the timed region: one fixed-work step and one
         converged-tolerance step; with WORLD_SIZE > 1 also the multi-rank overlap path (a17) against the
         oracle's overlapping multi-rank solve on the same partition, in the communication mode the run uses
  --impl reference: the CPU restatement of the reference algorithm (oracle/ only -- the product library is
         not loaded), all host cores as x-slabs with one overlap column, the SAME solver options as the device
         arm (BiCGStab + block ILU0 on 128-column sub-slabs, analytic volume Jacobian, 5 x 20 iterations) on a
         bounded sample of the same workload; the reference-default ILU(1) is reported beside it. The real DUNE
         binary cannot be built here.
  --workload paper: BASELINE configs[1], the square10mm.config action-potential run (100 x 180 cells), a "step"
         is one adaptive time step of the library's own time loop; myelin: configs[4].
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


def native_so():
    """in-tree shared objects this process has mapped (evidence: the device arm runs libax1gpu.so, the reference
    arm only the oracle)"""
    out = set()
    try:
        for line in open("/proc/self/maps"):
            p = line.split()[-1]
            if p.startswith(ROOT) and ".so" in p:
                out.add(os.path.relpath(p, ROOT))
    except Exception:
        pass
    return sorted(out)


def emit(obj):
    obj["native_so"] = native_so()
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


METRIC = "PNP assembly+solve throughput (DOF x Newton steps per second)"
UNIT = "GDOF/s"

NX_PER_GPU = 93184          # 728 * 128 cell columns per GPU
DX = 100.0                  # nm
NEWTON_ITS = 5
KRYLOV_ITS = 20
SLAB_W = 128
PAPER_NX, PAPER_DX = 100, 100.0e3        # square10mm.config
PAPER_CTL = dict(dtstart=1.0, lower_limit=10, upper_limit=30)   # [timeStepping] of square10mm.config
PAPER_SPLIT = 91     # radial cut of the library-default preconditioner on the 181-point radial grid (driver.cu: auto_radial_cut)
# algorithmic bytes per vertex. SURVEY 8(d) table for dense 4x4 blocks: jacobian 1184, spmv 1256,
# ilu_factor 2384. The device stores the Jacobian blocks in their 10-entry "arrow" shape (80 B instead
# of 128 B, common.cuh), so the bytes these kernels must move are smaller; roofline.achieved uses the
# bytes of the format actually stored, the dense-equivalent figure is reported beside it.
BYTES = {"residual": 104, "jacobian": 32 + 720, "spmv": 720 + 32 + 32, "ilu_factor": 720 + 1152, "ilu_apply": 1384,
         "dot": 64, "axpy": 256, "rupdate": 96, "pupdate": 128}
BYTES_DENSE = {"residual": 104, "jacobian": 1184, "spmv": 1256, "ilu_factor": 2384, "ilu_apply": 1384,
               "dot": 64, "axpy": 256, "rupdate": 96, "pupdate": 128}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of THIS round's code
    (profiles/r02_traffic.json, written from the .ncu-rep by examples/ncu_summary.py); None if there is none."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        e = t["kernels"][kernel]
        return float(e["dram_bytes_read"]) + float(e["dram_bytes_write"]), "profiles/r02_traffic.json (" + t["capture"] + ")"
    except Exception:
        return None, None


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


# ------------------------------------------------------------------------------------------------
# workload description shared by both arms (the driver compares the two `config` objects)
def workload_config(args, world):
    nv_full = (args.nx + 1) * 181
    if args.workload == "myelin":
        wl = ("myelinated axon 48 mm x 10 mm, 48 nodes of Ranvier (acme2_cyl_par_fiona_myelin_48mmx10mm_48nodes "
              "config): 7,251 x 180 cells in total, non-uniform x (100 nm in the nodes ... 10 um in the myelin), "
              "implicit Euler dt=1us, %d Newton x %d BiCGStab(ILU%d slab %d) per step"
              % (NEWTON_ITS, KRYLOV_ITS, args.ilu_fill, args.slab_w))
        return wl, {"workload": wl, "cells_total": 7251 * 180, "parallelism": "z-slab x%d" % world,
                    "l2": "working set ~3 GB in total; strong scaling"}
    if args.workload == "paper":
        wl = ("square10mm.config (2013 Biophys J action-potential set-up): 100 x 180 cells, 73,124 DOF, Na injection at "
              "x = 150 um from t = 20 us, HH channels, adaptive implicit Euler from the shipped equilibrium state; one "
              "step = one adaptive time step (Newton to the config's tolerances, BiCGStab + block ILU%d)" % args.ilu_fill)
        return wl, {"workload": wl, "cells_total": 18000, "dof_total": 73124, "parallelism": "single rank",
                    "l2": "working set 50 MB: L2-resident by nature of the configuration (latency-bound regime)"}
    wl = ("synthetic refined r-z tensor grid: Ny=180 (shipped 181-point radial vector) x Nx=%d columns/GPU, "
          "dx=100 nm, HH channels, implicit Euler dt=1us, %d Newton x %d BiCGStab(ILU%d slab %d) per step"
          % (args.nx, NEWTON_ITS, KRYLOV_ITS, args.ilu_fill, args.slab_w))
    return wl, {"workload": wl, "cells_per_gpu": args.nx * 180, "dof_total": 4 * nv_full * world,
                "parallelism": "z-slab x%d" % world,
                "l2": "inputs larger than L2 (working set ~%.0f GB per GPU: Jacobian, its 64-byte-block copy, ILU factor, "
                      "15 vectors)" % ((720 + 1152 + (576 if nv_full >= (1 << 20) else 0) + 15 * 32) * nv_full / 1e9)}


# ------------------------------------------------------------------------------------------------
# CPU legs: oracle/ only (never the product library)
def host_threads():
    """The host threads this process may run on (the affinity mask, not the machine's core count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_synthetic(nx_sample, threads, steps, warmup, ilu_fill, slab_w):
    """The oracle (reference algorithm restated) on a bounded sample: `threads` overlapping x-slabs, one thread each."""
    from oracle import cpu_setup as CS
    from oracle import pyoracle as O
    O.build()
    probs, us = [], []
    for rk in range(threads):
        P, u0 = CS.make_problem(nx_sample, DX, rank=rk, nranks=threads, perturbation=1e-3)
        P.set_time(0.0, 1.0); P.set_uold(u0); P.update_state(u0); P.accept_state()
        P.set_time(1.0, 1.0); P.set_uold(u0)
        probs.append(P); us.append(u0)
    nv_global = (nx_sample + 1) * 181
    opts = O.default_newton_opts(maxit=NEWTON_ITS, lin_maxit=KRYLOV_ITS, fixed_work=1, ilu_fill=ilu_fill, slab_w=slab_w)
    times, res = [], None
    for it in range(warmup + steps):
        for P, u in zip(probs, us):
            P.update_state(u)
        t0 = time.perf_counter()
        _, res, _ = O.newton_mt(probs, us, opts)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    t = float(np.mean(times))
    return dict(sec_per_step=t, nv=nv_global, threads=threads, gdof=4.0 * nv_global * NEWTON_ITS / t / 1e9,
                t_assembler=res[0].t_assembler, t_lin_solv=res[0].t_lin_solv)


def cpu_paper(threads, nsteps, ilu_fill, slab_w, ilu_split=0):
    """The paper configuration on the oracle: `threads` MPI-style x-slabs (the reference's docker default runs this
    configuration on 10 ranks, bin/run.sh:72), the restated time loop, the config's Newton tolerances."""
    from oracle import cpu_setup as CS
    from oracle import pyoracle as O
    from dune_ax1_b200.timeloop import TimeLoop   # pure Python control flow, loads no native library
    O.build()
    M = CS.MultiRankProblem(PAPER_NX, PAPER_DX, threads, stimulation=True)
    M.set_ilu_split(ilu_split)       # the radial cut of the device's default preconditioner (ax1gpu_set_ilu_split)
    tl = TimeLoop(M, M.u0, do_stimulation=True, newton_opts=O.default_newton_opts(ilu_fill=ilu_fill, slab_w=slab_w),
                  **PAPER_CTL)
    t0 = time.perf_counter()
    nits = sum(tl.step().iterations for _ in range(nsteps))
    wall = time.perf_counter() - t0
    return dict(newton_iterations=nits, wall=wall, t_end_us=tl.time, steps=nsteps, threads=threads,
                newton_steps_per_s=nits / wall, gdof=73124.0 * nits / wall / 1e9, trace=tl.trace)


def run_reference(args, world):
    threads = host_threads()
    workload, config = workload_config(args, world)
    if args.workload == "paper":
        ranks = min(threads, 10)
        cpu_paper(ranks, min(5, args.warmup), args.ilu_fill, 16)
        r = cpu_paper(ranks, args.steps, args.ilu_fill, 16 if args.ilu_fill == 0 else 0,
                      PAPER_SPLIT if args.ilu_fill == 0 else 0)
        sample = ("oracle/ restatement of the reference algorithm (not the DUNE binary), %d threads as %d x-slabs "
                  "(docker default: 10 MPI ranks), the first %d adaptive time steps of the run, %d Newton iterations in "
                  "%.2f s" % (ranks, ranks, args.steps, r["newton_iterations"], r["wall"]))
        out = {"impl": "reference", "metric": METRIC, "value": r["gdof"], "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["wall"] / args.steps * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config, "newton_steps_per_s": r["newton_steps_per_s"],
               "cpu_baseline": {"value": r["gdof"], "unit": UNIT, "cores": ranks, "kind": "port", "sample": sample},
               "e2e": {"value": r["gdof"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return
    if args.workload == "myelin":
        # the oracle has no membrane-group grid generator of its own; the synthetic sample stands in
        pass
    r = cpu_synthetic(args.cpu_nx, threads, args.steps, args.warmup, args.ilu_fill, args.slab_w)
    # GDOF/s = 4 * vertices * Newton iterations / time is size independent (work is linear in the
    # number of vertices), so the bounded sample needs no extrapolation
    r1 = cpu_synthetic(args.cpu_nx, threads, 1, 0, 1, 0)   # the reference's own default: one ILU(1) per rank
    full_nv = (args.nx + 1) * 181 * max(1, args.gpus)
    sample = ("CPU restatement of the reference algorithm (oracle/; not the DUNE binary), %d threads as %d "
              "x-slabs with 1 overlap column, BiCGStab+block ILU%d on %d-column sub-slabs (the device arm's options), "
              "analytic volume Jacobian; sample Nx=%d (%d vertices, %.2f s per step of %d Newton x %d BiCGStab "
              "iterations)" % (threads, threads, args.ilu_fill, args.slab_w, args.cpu_nx, r["nv"], r["sec_per_step"],
                               NEWTON_ITS, KRYLOV_ITS))
    out = {"impl": "reference", "metric": METRIC, "value": r["gdof"], "unit": UNIT,
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec_per_step"] * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config,
           "newton_steps_per_s_at_workload_size": NEWTON_ITS / r["sec_per_step"] * (r["nv"] / full_nv),
           "t_assembler": r["t_assembler"], "t_lin_solv": r["t_lin_solv"],
           "reference_default_ilu1": {"value": r1["gdof"], "unit": UNIT,
                                      "note": "same sample and iteration counts with the reference's default "
                                              "preconditioner SeqILUn(n=1), one factorisation per rank"},
           "cpu_baseline": {"value": r["gdof"], "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": r["gdof"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


# ------------------------------------------------------------------------------------------------
# parity legs of the device arm (the oracle is the checker here, never the thing measured)
def _field_err(ug, uc):
    scale = np.max(np.abs(uc.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(ug - uc).reshape(-1, 4) / scale, axis=0)
    return float(err[:3].max()), float(err[3])


def sample_parity(api, setups, device, ilu_fill, slab_w, nx_fixed=512, nx_conv=256, compact=False):
    """Device vs oracle on bounded samples of the synthetic workload with exactly the same options: (a) the
    fixed-work step the headline times (5 Newton x 20 BiCGStab, lambda = 1), (b) a step converged to tight
    tolerances. The oracle legs run in threads beside the device legs. compact: the headline grid is large enough
    for the library to multiply with the 64-byte-block copy of the Jacobian; the samples are too small for the
    automatic rule, so the same operand is forced on them."""
    from oracle import cpu_setup as CS
    from oracle import pyoracle as O
    O.build()
    cases = {"fixed_work": (nx_fixed, dict(maxit=NEWTON_ITS, lin_maxit=KRYLOV_ITS, fixed_work=1)),
             "converged": (nx_conv, dict(reduction=1e-8, abs_limit=1e-14, min_lin_reduction=1e-5, maxit=30))}
    cpu, th = {}, []

    def cpu_leg(name, nx, kw):
        P, u0 = CS.make_problem(nx, DX, perturbation=1e-3)
        P.set_time(0.0, 1.0); P.set_uold(u0); P.update_state(u0); P.accept_state()
        P.set_time(1.0, 1.0); P.set_uold(u0); P.update_state(u0)
        t0 = time.perf_counter()
        u, res = P.newton(u0, O.default_newton_opts(ilu_fill=ilu_fill, slab_w=slab_w, **kw))
        cpu[name] = (u, res, time.perf_counter() - t0)

    for name, (nx, kw) in cases.items():
        t = threading.Thread(target=cpu_leg, args=(name, nx, kw))
        t.start(); th.append(t)
    gpu = {}
    for name, (nx, kw) in cases.items():
        s = setups.make_setup(nx, dx=DX, perturbation=1e-3)
        dev = setups.make_device(s, device=device)
        dev.set_time(0.0, 1.0); dev.set_uold(s["u0"]); dev.update_state(s["u0"]); dev.accept_state()
        dev.set_time(1.0, 1.0); dev.set_uold(s["u0"]); dev.update_state(s["u0"])
        if compact:
            dev.set_spmv_compact(1)
        t0 = time.perf_counter()
        u, res = dev.newton(s["u0"], api.default_newton_opts(ilu_fill=ilu_fill, slab_w=slab_w, **kw))
        gpu[name] = (u, res, time.perf_counter() - t0)
        operand_compact = dev.spmv_compact()
        dev.close()
    for t in th:
        t.join()
    out = {"checker": "oracle/ (CPU restatement, pinned on the reference's compiled operators)",
           "options": "ILU%d, %d-column sub-slabs, dx = 100 nm, perturbation 1e-3, dt = 1 us" % (ilu_fill, slab_w),
           "spmv_operand": "64-byte blocks (forced, as on the headline grid)" if operand_compact else "arrow blocks"}
    ok = True
    for name, (nx, kw) in cases.items():
        (ug, rg, tg), (uc, rc, tc) = gpu[name], cpu[name]
        econ, epot = _field_err(ug, uc)
        out[name] = {"nx": nx, "max_rel_err_con": econ, "max_rel_err_pot": epot,
                     "newton_its_gpu": rg.iterations, "newton_its_cpu": rc.iterations,
                     "krylov_its_gpu": rg.lin_iterations, "krylov_its_cpu": rc.lin_iterations,
                     "defect_gpu": rg.defect, "defect_cpu": rc.defect, "status_gpu": rg.status, "status_cpu": rc.status,
                     "wall_gpu_s": tg, "wall_cpu_s": tc}
        ok &= rg.status == 0 and rc.status == 0 and rg.iterations == rc.iterations
    out["max_rel_err_con"] = max(out[n]["max_rel_err_con"] for n in cases)
    out["max_rel_err_pot"] = max(out[n]["max_rel_err_pot"] for n in cases)
    # Tolerances: the north star's 1e-8 for converged concentrations. The potential carries one weakly determined
    # mode that a defect-based Newton stop does not see: two correct double-precision solves stopped by the same
    # criterion agree in it only to ~1e-6 of max|phi| on grids of this size, and the double-precision oracle itself
    # sits that far from the 80-bit solution of the same discrete problem (measured with the extended-precision
    # build of the oracle: tests/test_precision_floor.py, test_gpu_parity.py::test_fields_at_rounding_floor...).
    # The fixed-work step stops far from convergence (20 Krylov iterations per solve); it is held to equal
    # iteration counts and reported, not to the converged-field tolerance.
    c = out["converged"]
    out["tolerance_con"], out["tolerance_pot"] = 1e-8, 1e-6
    out["green"] = bool(ok and c["max_rel_err_con"] < 1e-8 and c["max_rel_err_pot"] < 1e-6 and
                        out["fixed_work"]["krylov_its_gpu"] == out["fixed_work"]["krylov_its_cpu"])
    return out


def multigpu_parity(api, setups, dist, rank, world, local_rank, compact=False):
    """a17 where the driver sees it: a small problem on the SAME ranks / communicator kind the bench uses, device vs
    the oracle's overlapping multi-rank Newton solve on the same partition (ILU0 on narrow sub-slabs and the
    reference-default ILU(1) per rank)."""
    nx, slab_w = 6 * world, 4
    s = setups.make_setup(nx, dx=100.0e3, rank=rank, nranks=world, perturbation=1e-3)
    dev = setups.make_device(s, device=local_rank)
    ids = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    dev.comm_init(ids[0], rank, world)
    if compact:     # the SpMV operand of the benchmarked grid (see sample_parity)
        dev.set_spmv_compact(1)
    kw = dict(reduction=1e-9, abs_limit=1e-14, min_lin_reduction=1e-10)
    runs = {}
    for name, o in (("ilu0_slab4", dict(slab_w=slab_w, ilu_fill=0)), ("ilu1_per_rank", dict(slab_w=64, ilu_fill=1))):
        dev.set_time(0.0, 1.0); dev.set_uold(s["u0"]); dev.update_state(s["u0"])
        u, r = dev.newton(s["u0"], api.default_newton_opts(**o, **kw))
        runs[name] = dict(u=u, its=r.iterations, lin=r.lin_iterations, status=r.status)
    _, nrm = dev.residual(s["u0"])
    mode = dev.comm_mode()
    operand_compact = dev.spmv_compact()
    dev.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(runs=runs, nrm=nrm, mode=mode))
    if rank != 0:
        return None
    from oracle import cpu_setup as CS
    from oracle import pyoracle as O
    O.build()
    out = {"ranks": world, "comm": {1: "nccl", 2: "nvlink peer windows"}.get(mode, str(mode)), "nx": nx,
           "tolerance_con": 1e-8, "tolerance_pot": 5e-7, "cases": {}}
    ok = len({g["nrm"] for g in gathered}) == 1 and len({g["mode"] for g in gathered}) == 1
    for name, o in (("ilu0_slab4", dict(slab_w=slab_w, ilu_fill=0)), ("ilu1_per_rank", dict(slab_w=0, ilu_fill=1))):
        probs, us = [], []
        for r in range(world):
            P, u0 = CS.make_problem(nx, 100.0e3, rank=r, nranks=world, perturbation=1e-3)
            P.set_time(0.0, 1.0); P.set_uold(u0); P.update_state(u0)
            probs.append(P); us.append(u0)
        uc, rc, st = O.newton_mt(probs, us, O.default_newton_opts(**o, **kw))
        econ = epot = 0.0
        its_ok = st == 0
        for r in range(world):
            g = gathered[r]["runs"][name]
            a, b = _field_err(g["u"], uc[r])
            econ, epot = max(econ, a), max(epot, b)
            its_ok &= g["status"] == 0 and g["its"] == rc[r].iterations
        out["cases"][name] = {"max_rel_err_con": econ, "max_rel_err_pot": epot,
                              "newton_its_gpu": gathered[0]["runs"][name]["its"], "newton_its_cpu": rc[0].iterations,
                              "krylov_its_gpu": gathered[0]["runs"][name]["lin"], "krylov_its_cpu": rc[0].lin_iterations}
        ok &= its_ok and econ < 1e-8 and epot < 5e-7     # potential: see sample_parity on its floor
    out["green"] = bool(ok)
    out["spmv_operand"] = "64-byte blocks (forced, as on the benchmarked grid)" if operand_compact else "arrow blocks"
    return out


# ------------------------------------------------------------------------------------------------
def run_paper(args, api, setups, local_rank):
    """BASELINE configs[1] as its own workload: K adaptive time steps of the stimulated run through the library's
    time loop (state resident in HBM), after W warm-up steps of the same run."""
    import torch
    from dune_ax1_b200.timeloop import DeviceTimeLoop
    workload, config = workload_config(args, 1)
    sp = setups.make_setup(PAPER_NX, dx=PAPER_DX, stimulation=True, perturbation=0.0)
    stream = torch.cuda.Stream(device=local_rank)
    dev = setups.make_device(sp, device=local_rank, stream=stream.cuda_stream)
    opts = api.default_newton_opts(slab_w=0, ilu_fill=args.ilu_fill)
    # warm-up: a throw-away copy of the run (module load, first launches)
    tlw = DeviceTimeLoop(dev, sp["u0"], newton_opts=opts, **PAPER_CTL)
    for _ in range(max(3, args.warmup)):
        tlw.step()
    dev.close()
    dev = setups.make_device(sp, device=local_rank, stream=stream.cuda_stream)
    tl = DeviceTimeLoop(dev, sp["u0"], newton_opts=opts, **PAPER_CTL)
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    l0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nits = lin = 0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            r = tl.step()
            nits += r.newton_iterations; lin += r.lin_iterations
        ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = dev.launch_count() - l0
    t_end = tl.time
    split = dev.ilu_split()
    dev.close()
    # e2e: the host-side loop -- u, uold up and u down through the C ABI every step
    from dune_ax1_b200.timeloop import TimeLoop
    dev = setups.make_device(sp, device=local_rank, stream=stream.cuda_stream)
    th = TimeLoop(dev, sp["u0"], do_stimulation=True, newton_opts=opts, **PAPER_CTL)
    t0 = time.perf_counter()
    nits_h = sum(th.step().iterations for _ in range(args.steps))
    wall_h = time.perf_counter() - t0
    n = dev.n
    dev.close()
    out = {"metric": METRIC, "value": 73124.0 * nits / (ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
           "newton_steps_per_s": nits / (ms * 1e-3), "newton_iterations": nits, "krylov_iterations": lin,
           "t_end_us": t_end, "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": 73124.0 * nits_h / wall_h / 1e9, "unit": UNIT, "newton_steps_per_s": nits_h / wall_h,
                   "h2d_bytes_per_step": 3 * n * 8, "d2h_bytes_per_step": n * 8, "timing": "host wall clock"}}
    if not args.no_cpu_baseline:
        ranks = min(host_threads(), 10)
        c = cpu_paper(ranks, args.steps, args.ilu_fill, 16 if args.ilu_fill == 0 else 0, split)
        out["cpu_baseline"] = {"value": c["gdof"], "unit": UNIT, "cores": ranks, "kind": "port",
                               "newton_steps_per_s": c["newton_steps_per_s"],
                               "sample": "oracle/ restatement, %d threads as x-slabs, the same %d time steps (%d Newton "
                                         "iterations, t_end %.3f us)" % (ranks, args.steps, c["newton_iterations"],
                                                                        c["t_end_us"])}
        out["parity"] = {"newton_its_gpu": nits, "newton_its_cpu": c["newton_iterations"],
                         "t_end_gpu_us": t_end, "t_end_cpu_us": c["t_end_us"],
                         "green": bool(nits == c["newton_iterations"] and abs(t_end - c["t_end_us"]) < 1e-9)}
    emit(out)


def run_mori(args, api, setups, local_rank):
    """BASELINE configs[2]: the electroneutral (Mori) operator-split model on the same synthetic grid (Nx columns,
    dx = 100 nm, shipped radial vector). One step = one time step of Acme2CylMoriSimulation::run: channel update,
    potential sub-problem (assemble + BiCGStab/ILU0), charge-layer update, concentration sub-problem, accept. Value =
    DOF x steps per second; every step solves all 4 DOF per vertex once (1 + 3), both linear solves to the config's
    reduction. A sample of the same steps runs on the oracle for parity and as CPU figure."""
    import torch
    workload = ("electroneutral (Mori) operator-split model, synthetic refined r-z grid Ny=180 x Nx=%d, dx=100 nm, "
                "dt=1us: per step one potential and one concentration sub-problem solve (BiCGStab + ILU0 slab %d, "
                "reduction 1e-5)" % (args.nx, args.slab_w))
    config = {"workload": workload, "cells_per_gpu": args.nx * 180, "dof_total": 4 * (args.nx + 1) * 181,
              "parallelism": "single rank", "l2": "inputs larger than L2" if args.nx >= 8192 else "reduced size"}
    s = setups.make_setup(args.nx, dx=DX, perturbation=1e-3)
    stream = torch.cuda.Stream(device=local_rank)
    dev = setups.make_device(s, device=local_rank, stream=stream.cuda_stream)
    opts = api.default_newton_opts(reduction=1e-5, abs_limit=1e-5, lin_maxit=5000, slab_w=args.slab_w, ilu_fill=0)
    u_host_t = torch.empty(dev.n, dtype=torch.float64, pin_memory=True)
    u_host = u_host_t.numpy()
    u_host[:] = s["u0"]
    dev.upload_u(u_host)
    t = [0.0]

    def step():
        dev.set_time(t[0], 1.0)
        dev.set_uold_dev()
        dev.update_state()
        _, r = dev.mori_step(None, opts)
        dev.accept_state()
        t[0] += 1.0
        return r

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.synchronize()
    l0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    its = [0, 0]
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            r = step()
            its[0] += r.pot_iterations; its[1] += r.con_iterations
        ev1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = dev.launch_count() - l0
    # e2e: host buffers in and out every step
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dev.set_time(t[0], 1.0); dev.set_uold(u_host); dev.update_state(u_host)
        un, _ = dev.mori_step(u_host, opts)
        u_host[:] = un
        dev.accept_state(); t[0] += 1.0
    wall = time.perf_counter() - t0
    n = dev.n
    dev.close()
    out = {"metric": "electroneutral (Mori) step throughput (DOF x time steps per second)", "value": n * args.steps / (ms * 1e-3) / 1e9,
           "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config, "krylov_iterations_per_step": {"potential": its[0] / args.steps, "concentration": its[1] / args.steps},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": n * args.steps / wall / 1e9, "unit": UNIT, "h2d_bytes_per_step": 3 * n * 8,
                   "d2h_bytes_per_step": n * 8, "timing": "host wall clock"}}
    if not args.no_cpu_baseline:
        from oracle import cpu_setup as CS
        from oracle import pyoracle as O
        O.build()
        nxs = min(args.nx, 512)
        sg = setups.make_setup(nxs, dx=DX, perturbation=1e-3)
        dg = setups.make_device(sg, device=local_rank)
        P, u0 = CS.make_problem(nxs, DX, perturbation=1e-3)
        oc = O.default_newton_opts(reduction=1e-5, abs_limit=1e-5, lin_maxit=5000, slab_w=args.slab_w, ilu_fill=0)
        ug = uc = u0
        tt, tcpu, itg, itc = 0.0, 0.0, [0, 0], [0, 0]
        for _ in range(3):
            for X, u in ((dg, ug), (P, uc)):
                X.set_time(tt, 1.0); X.set_uold(u); X.update_state(u)
            ug, rg = dg.mori_step(ug, opts)
            t0 = time.perf_counter()
            uc, rc = P.mori_step(uc, oc)
            tcpu += time.perf_counter() - t0
            dg.accept_state(); P.accept_state(); tt += 1.0
            itg[0] += rg.pot_iterations; itg[1] += rg.con_iterations; itc[0] += rc.pot_iterations; itc[1] += rc.con_iterations
        dg.close()
        econ, epot = _field_err(ug, uc)
        out["parity"] = {"nx": nxs, "steps": 3, "max_rel_err_con": econ, "max_rel_err_pot": epot, "krylov_its_gpu": itg,
                         "krylov_its_cpu": itc, "green": bool(econ < 1e-8 and epot < 1e-6)}
        out["cpu_baseline"] = {"value": 4.0 * (nxs + 1) * 181 * 3 / tcpu / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": "oracle/mori.cpp on one core, Nx=%d, 3 time steps in %.1f s" % (nxs, tcpu)}
    emit(out)


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
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-paper", action="store_true", help="skip the secondary paper-configuration figure")
    ap.add_argument("--no-converged", action="store_true", help="skip the converged-tolerance step at full size")
    ap.add_argument("--workload", default="synthetic", choices=["synthetic", "myelin", "paper", "mori"],
                    help="synthetic: weak scaling, Nx columns PER GPU (headline); myelin: BASELINE config 5, the 48 mm "
                         "myelinated axon with 48 nodes of Ranvier (7,251 x 180 cells), STRONG scaling over z-slabs; "
                         "paper: BASELINE config 2, square10mm.config action-potential run (single GPU); "
                         "mori: BASELINE config 3, the electroneutral (Mori) operator-split model on the synthetic grid")
    ap.add_argument("--ilu-fill", type=int, default=0, choices=[0, 1])
    args = ap.parse_args()
    myelin = args.workload == "myelin"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if myelin and args.slab_w == SLAB_W:
        args.slab_w = 0     # library default on this grid: one-warp sub-slabs (16 / 24 columns) + far-field radial cut
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, max(world, args.gpus))
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
    if args.workload == "paper":
        if rank == 0:
            run_paper(args, api, setups, local_rank)
        return
    if args.workload == "mori":
        if rank == 0:
            run_mori(args, api, setups, local_rank)
        return
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", rank=rank, world_size=world)
    workload, config = workload_config(args, world)
    nv_full = (args.nx + 1) * 181

    # ---- parity on the benchmarked configuration, before anything is timed ----
    parity = mg_parity = None
    if not args.no_parity and not myelin:
        if world > 1:
            mg_parity = multigpu_parity(api, setups, dist, rank, world, local_rank, compact=nv_full >= (1 << 20))
        elif rank == 0:
            parity = sample_parity(api, setups, local_rank, args.ilu_fill, args.slab_w, compact=nv_full >= (1 << 20))

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

    def one_step_resident(t, o=opts):
        dev.set_time(t, 1.0)
        dev.update_state()          # gfMembFlux.updateState(time, dt) of the time loop
        dev.set_uold_dev()          # r0 = -M(uold), uold = current device iterate
        return dev.newton_dev(o)

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
    # ---- one step solved to tolerance next to the fixed-work figure (what the fixed 5 x 20 iterations hide):
    # Newton to the reference's default reduction with BiCGStab converged to newtonMinLinReduction
    conv = None
    if not args.no_converged:
        u_host[:] = s["u0"]
        dev.upload_u(u_host)
        co = api.default_newton_opts(reduction=1e-5, abs_limit=1e-5, min_lin_reduction=1e-5, maxit=20, lin_maxit=5000,
                                     slab_w=args.slab_w, ilu_fill=args.ilu_fill)
        barrier()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            cr = one_step_resident(t_sim, co)
            ev1.record(stream)
        barrier()
        conv = {"newton_its": cr.iterations, "krylov_its": cr.lin_iterations, "status": cr.status, "ms": ev0.elapsed_time(ev1),
                "first_defect": cr.first_defect, "defect": cr.defect,
                "tolerances": "newtonReduction 1e-5, newtonAbsLimit 1e-5, newtonMinLinReduction 1e-5 (config defaults)"}
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, conv["ms"] if conv else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(tt[0]), float(tt[1])
        if conv:
            conv["ms"] = float(tt[2])
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt[0])
        # every rank's own GPU clock under load: the ranks synchronise in every reduction, so the step runs at the
        # pace of the slowest GPU -- under sw_power_cap the GPUs of one box do not all hold the same clock
        per_rank = [None] * world
        dist.all_gather_object(per_rank, {"rank": rank, "sm_mhz": clocks.get("sm_mhz"), "reasons": clocks.get("reasons")})
        clocks["per_rank_sm_mhz"] = [p["sm_mhz"] for p in per_rank]
        clocks["sm_mhz_min_over_ranks"] = min(p["sm_mhz"] for p in per_rank if p["sm_mhz"] is not None)
        clocks["reasons"] = sorted(set(r for p in per_rank for r in (p["reasons"] or [])))
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
    # (the Krylov dot products are fused into the SpMV and the x,r updates), rupdate = r -= alpha v of the first half
    # step, axpy = the fused x,r update of the second (x += alpha y + omega z, r -= omega t);
    # spmv_dot1 / spmv_dot2 are the two dot-fused SpMV variants BiCGStab launches (v = A y, t = A y)
    counts = {"residual": 2 * NEWTON_ITS + 1, "jacobian": NEWTON_ITS, "ilu_factor": NEWTON_ITS,
              "spmv_dot1": KRYLOV_ITS * NEWTON_ITS, "spmv_dot2": KRYLOV_ITS * NEWTON_ITS,
              "ilu_apply": 2 * KRYLOV_ITS * NEWTON_ITS,
              "dot": 3 * NEWTON_ITS + 1, "axpy": KRYLOV_ITS * NEWTON_ITS, "rupdate": KRYLOV_ITS * NEWTON_ITS,
              "pupdate": KRYLOV_ITS * NEWTON_ITS}
    kern = {}
    bytes_k, bytes_dense = dict(BYTES), dict(BYTES_DENSE)
    compact = bool(dev.spmv_compact())
    if compact:   # the solver multiplies with the 64-byte-block copy (576 B per block row) each factorisation writes
        bytes_k["spmv"] = 576 + 32 + 32
        bytes_k["compact"], bytes_dense["compact"] = 720 + 576, 720 + 576
        counts["compact"] = NEWTON_ITS
    for b in (bytes_k, bytes_dense):
        b["spmv_dot1"], b["spmv_dot2"] = b["spmv"] + 32, b["spmv"] + 32    # + the vector of the fused dot product
    if args.ilu_fill == 1:   # 13 instead of 9 blocks per factor row (ilu1.cu)
        bytes_k["ilu_factor"], bytes_k["ilu_apply"] = 720 + 1664, 1664 + 64
        bytes_dense["ilu_factor"], bytes_dense["ilu_apply"] = 1192 + 1664, 1664 + 64
    for name in counts:
        t_ms = dev.time_kernel(name, reps=3)
        kern[name] = {"ms": t_ms, "launches_per_step": counts[name],
                      "GBps": bytes_k[name] * nv / (t_ms * 1e-3) / 1e9,
                      "GBps_dense_equivalent": bytes_dense[name] * nv / (t_ms * 1e-3) / 1e9,
                      "share_of_step": counts[name] * t_ms / (ms / args.steps)}
    # the two exchange steps of the overlapping solve (scalar reduction incl. the cross-rank sum; additive halo sum),
    # timed in lock step on all ranks: what the weak-scaling loss is made of
    comm = {"reduce_ms": dev.time_kernel("reduce", reps=50),
            "reduce_ms_is": "one dot kernel over two vectors + the reduction of its partials incl. the cross-rank sum",
            "reductions_per_step": 4 * KRYLOV_ITS * NEWTON_ITS + 3 * NEWTON_ITS + 1}
    if world > 1:
        comm["halo_ms"] = dev.time_kernel("halo", reps=50)
        comm["halo_sums_per_step"] = 2 * KRYLOV_ITS * NEWTON_ITS
    dom = max(kern, key=lambda k: kern[k]["share_of_step"])
    traffic, traffic_src = (None, None)
    if args.nx == NX_PER_GPU and args.slab_w == SLAB_W and not myelin and args.ilu_fill == 0:
        traffic, traffic_src = ncu_traffic(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "frac_of_nominal_8TBs": kern[dom]["GBps"] / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                "peak_kind": peak_kind, "algorithmic_bytes_per_launch": bytes_k[dom] * nv, "kernels": kern}
    # whole-step algorithmic traffic (SURVEY 8(d) formula, l = 1 line-search residual, k Krylov its)
    step_bytes = NEWTON_ITS * (bytes_k["jacobian"] + bytes_k["ilu_factor"] + bytes_k.get("compact", 0) + 2 * 104 +
                               KRYLOV_ITS * (2 * bytes_k["spmv"] + 2 * bytes_k["ilu_apply"] + 416)) * nv
    step_bytes_dense = NEWTON_ITS * (1184 + bytes_dense["ilu_factor"] + 2 * 104 +
                                     KRYLOV_ITS * (2 * 1256 + 2 * bytes_dense["ilu_apply"] + 416)) * nv
    roofline["step_frac_of_peak"] = step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak
    roofline["step_frac_of_peak_dense_equivalent"] = step_bytes_dense / (ms / args.steps * 1e-3) / 1e9 / peak
    roofline["step_frac_of_nominal_8TBs"] = step_bytes / (ms / args.steps * 1e-3) / 1e9 / 8000.0

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
           "scaling": "strong" if myelin else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config, "comm": {0: "none", 1: "nccl", 2: "nvlink peer windows"}[dev.comm_mode()],
           "ilu_radial_split_row": dev.ilu_split(),
           "spmv_operand": "64-byte blocks (576 B per block row)" if compact else "arrow blocks (720 B per block row)",
           "newton_steps_per_s": newton_per_s, "newton_iterations_per_step": NEWTON_ITS,
           "krylov_iterations_per_solve": KRYLOV_ITS,
           "last_step": {"defect": last.defect, "first_defect": last.first_defect, "lin_iterations": last.lin_iterations,
                         "t_assembler": last.t_assembler, "t_lin_solv": last.t_lin_solv},
           "clocks": clocks, "gpu_launches": launches,
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 3 * dev.n * 8,
                   "d2h_bytes_per_step": dev.n * 8, "ms_per_step": ms_e2e / args.steps},
           "roofline": roofline, "exchange": comm}
    if conv:
        out["converged_step"] = conv
    if parity:
        out["parity"] = parity
    if mg_parity:
        out["multigpu_parity"] = mg_parity
    dev.close()
    if rank == 0 and world == 1 and not myelin and not args.no_paper:
        # secondary figure: the paper configuration (square10mm.config, 100 x 180 cells, 73,124 DOF), first 60
        # adaptive time steps of the stimulated run through the library's time loop (latency-bound regime), with
        # the oracle's run of the same steps beside it (10 x-slab threads = the reference's docker default)
        try:
            from dune_ax1_b200.timeloop import DeviceTimeLoop
            sp = setups.make_setup(PAPER_NX, dx=PAPER_DX, stimulation=True, perturbation=0.0)
            devp = setups.make_device(sp, device=local_rank)
            tl = DeviceTimeLoop(devp, sp["u0"], newton_opts=api.default_newton_opts(slab_w=0), **PAPER_CTL)
            for _ in range(3):
                tl.step()
            devp.close()
            devp = setups.make_device(sp, device=local_rank)
            tl = DeviceTimeLoop(devp, sp["u0"], newton_opts=api.default_newton_opts(slab_w=0), **PAPER_CTL)
            t0 = time.perf_counter()
            nits = sum(tl.step().iterations for _ in range(60))
            wallp = time.perf_counter() - t0
            split = devp.ilu_split()
            out["paper_config"] = {"workload": "square10mm.config grid, 73,124 DOF, 60 adaptive time steps from t=0 "
                                               "(Na injection starts at 20 us), reference Newton tolerances 1e-5, "
                                               "library-default sub-slab width, time loop inside the library",
                                   "newton_steps_per_s": nits / wallp, "newton_iterations": nits, "t_end_us": tl.time,
                                   "ilu_radial_split_row": split, "timing": "host wall clock"}
            devp.close()
            if not args.no_cpu_baseline:
                ranks = min(host_threads(), 10)
                c = cpu_paper(ranks, 60, 0, 16, split)
                out["paper_config"]["cpu_baseline"] = {
                    "newton_steps_per_s": c["newton_steps_per_s"], "cores": ranks, "kind": "port",
                    "newton_iterations": c["newton_iterations"], "t_end_us": c["t_end_us"],
                    "sample": "oracle/ restatement, %d threads as x-slabs, the same 60 time steps" % ranks}
                out["paper_config"]["parity"] = {
                    "same_newton_history": bool(nits == c["newton_iterations"] and abs(tl.time - c["t_end_us"]) < 1e-9)}
        except Exception as e:  # never lose the headline line because of the secondary figure
            out["paper_config"] = {"error": str(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not myelin:
        threads = host_threads()
        r = cpu_synthetic(args.cpu_nx, threads, 1, 0, args.ilu_fill, args.slab_w)
        out["cpu_baseline"] = {
            "value": r["gdof"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "oracle/ restatement of the reference algorithm (not the DUNE binary), %d threads as x-slabs with 1 "
                      "overlap column, BiCGStab+block ILU%d on %d-column sub-slabs (the device arm's options), analytic "
                      "volume Jacobian, Nx=%d sample (%d vertices, %.1f s per step of %d Newton x %d BiCGStab iterations)"
                      % (threads, args.ilu_fill, args.slab_w, args.cpu_nx, r["nv"], r["sec_per_step"], NEWTON_ITS,
                         KRYLOV_ITS),
            "newton_steps_per_s_at_workload_size": NEWTON_ITS / r["sec_per_step"] * (r["nv"] / nv_total)}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
