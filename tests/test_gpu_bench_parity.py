"""The benchmark line itself carries the parity of the benchmarked configuration: bench.py runs the device against
the oracle on a bounded sample of its own workload (same dx, perturbation, sub-slab width, solver options) before
the timed region. This test runs a reduced-size bench.py and checks that part of its contract."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_carries_green_parity_and_contract_keys():
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--nx", "4096", "--steps", "1", "--warmup", "3",
           "--cpu-nx", "256"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline", "parity"):
        assert k in d, k
    p = d["parity"]
    # converged concentrations within the north star's 1e-8, the potential within its double-precision floor
    # (tests/test_precision_floor.py); equal Newton / Krylov histories in the fixed-work leg
    assert p["green"], p
    assert p["converged"]["max_rel_err_con"] < 1e-8 and p["converged"]["max_rel_err_pot"] < 1e-6, p
    assert p["fixed_work"]["newton_its_gpu"] == p["fixed_work"]["newton_its_cpu"] == 5
    assert p["fixed_work"]["krylov_its_gpu"] == p["fixed_work"]["krylov_its_cpu"] == 100
    assert "dune_ax1_b200/libax1gpu.so" in d["native_so"]
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    assert d["paper_config"]["parity"]["same_newton_history"], d["paper_config"]
