"""oracle/cpu_setup.py (problems for the CPU oracle built without the product library, used by bench.py's
reference arm) against dune_ax1_b200/setups.py + the host model of the product: same grid, partition, Physics
maps, configuration and initial state -- the two arms of the benchmark run the same problem. CPU only."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import oracle_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nx,dx,nranks,stim", [(40, 100.0, 3, False), (100, 100.0e3, 1, True), (37, 100.0, 4, False)])
def test_cpu_setup_matches_product_setup(oracle, axlib, nx, dx, nranks, stim):
    from dune_ax1_b200 import setups
    from oracle import cpu_setup as CS
    for rk in range(nranks):
        P, u0 = CS.make_problem(nx, dx, rank=rk, nranks=nranks, perturbation=1e-3, stimulation=stim)
        s = setups.make_setup(nx, dx=dx, rank=rk, nranks=nranks, perturbation=1e-3, stimulation=stim)
        Q = oracle_problem(oracle, s)
        assert CS.slab(nx, nranks, rk) == axlib.slab(nx, nranks, rk) or nranks == 1
        assert np.array_equal(u0, s["u0"]) and np.array_equal(P.x, s["x"]) and np.array_equal(P.y, s["y"])
        for a, b in zip(P.maps(), Q.maps()):
            assert np.array_equal(a, b)
        assert bytes(P.cfg) == bytes(Q.cfg)
        # and the product's own host model hands the device the same maps
        assert np.array_equal(P.maps()[0], s["cell_subdomain"]) and np.array_equal(P.maps()[2], s["dirichlet"])
        assert np.array_equal(P.maps()[1], s["node_volume"])


def test_state_parser_matches_library_loader(axlib):
    from oracle import cpu_setup as CS
    a, b = CS.parse_state(CS.STATE_SQUARE10MM), axlib.load_state(CS.STATE_SQUARE10MM)
    for k in ("x", "y", "channelStates", "uold", "unew"):
        assert np.array_equal(a[k], b[k])
    assert a["elements"] == b["elements"] and a["time"] == b["time"] and a["dt"] == b["dt"]


def test_multirank_problem_scatter_gather_round_trip(oracle):
    from oracle import cpu_setup as CS
    M = CS.MultiRankProblem(23, 100.0e3, 4, perturbation=1e-3)
    P, u0 = CS.make_problem(23, 100.0e3, perturbation=1e-3)
    assert np.array_equal(M.u0, u0)
    assert np.array_equal(M.gather(M.scatter(u0)), u0)
    assert np.allclose(M.membrane_potential(u0), P.membrane_potential(u0), rtol=0, atol=0)


def test_reference_arm_runs_without_the_product_library():
    """bench.py --impl reference prints one JSON line of the contract, honours --steps, runs the device arm's solver
    options and maps no in-tree library except the oracle."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--cpu-nx", "48"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["unit"] == "GDOF/s"
    assert d["native_so"] == ["oracle/libax1oracle.so"]
    assert "ILU0 slab 128" in d["config"]["workload"] and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
    # the paper workload through the same arm
    out = subprocess.run(cmd[:2] + ["--impl", "reference", "--workload", "paper", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["steps"] == 3 and d["native_so"] == ["oracle/libax1oracle.so"]
