"""N>1 host logic on the CPU: z-slab partition, overlap constraints, owner mask, additive halo sum
and allreduce semantics, checked with the oracle -- in-process (one thread per rank)
and across two real processes over torch.distributed/gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import oracle_problem, small_setup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition(axlib):
    for nx, p in ((9, 3), (10, 4), (93184 * 8, 8), (7, 7)):
        cover = []
        for r in range(p):
            c0, c1, l0, l1 = axlib.slab(nx, p, r)
            assert l0 == c0 - (r > 0) and l1 == c1 + (r < p - 1)
            assert c1 > c0
            cover += list(range(c0, c1)) if nx < 100 else [c0, c1]
        if nx < 100:
            assert cover == list(range(nx))
    with pytest.raises(axlib.Ax1GpuError):
        axlib.slab(3, 4, 0)  # px must not exceed Nx (ax1_gridgenerator.hh:615-619)


def _rank_problems(oracle, nx, nranks, **kw):
    from dune_ax1_b200 import setups
    probs, us, ss = [], [], []
    for r in range(nranks):
        s = setups.make_setup(nx, dx=100.0e3, rank=r, nranks=nranks, perturbation=1e-3, **kw)
        P = oracle_problem(oracle, s)
        P.set_time(0.0, 1.0); P.set_uold(s["u0"]); P.update_state(s["u0"])
        probs.append(P); us.append(s["u0"]); ss.append(s)
    return probs, us, ss


def _gather(us, ss, nx, nvy):
    """Assemble the global vector from the owned vertex columns of every rank."""
    out = np.zeros((nvy, nx + 1, 4))
    for u, s in zip(us, ss):
        c0, c1, l0, l1 = s["slab"]
        loc = u.reshape(nvy, l1 - l0 + 1, 4)
        lo = 2 if s["left_pb"] else 0
        hi = (l1 - l0 - 1) if s["right_pb"] else (l1 - l0)
        out[:, l0 + lo:l0 + hi + 1, :] = loc[:, lo:hi + 1, :]
    return out.reshape(-1)


@pytest.mark.parametrize("nranks", [2, 3])
def test_overlapping_newton_matches_single_rank(oracle, axlib, nranks):
    nx = 9
    opts = oracle.default_newton_opts(reduction=1e-9, abs_limit=1e-14, min_lin_reduction=1e-10, ilu_fill=1)
    p1, u1, s1 = _rank_problems(oracle, nx, 1)
    ref, res1 = p1[0].newton(u1[0], opts)
    assert res1.status == 0
    probs, us, ss = _rank_problems(oracle, nx, nranks)
    # local masks: processor-boundary columns are fully constrained, ownership is disjoint
    owned = np.zeros(nx + 1, dtype=int)
    for P, s in zip(probs, ss):
        _, _, dm = P.maps()
        dm = dm.reshape(s["ny"] + 1, s["nx"] + 1)
        if s["left_pb"]:
            assert np.all(dm[:, 0] == 0xF)
        if s["right_pb"]:
            assert np.all(dm[:, -1] == 0xF)
        c0, c1, l0, l1 = s["slab"]
        lo = 2 if s["left_pb"] else 0
        hi = (l1 - l0 - 1) if s["right_pb"] else (l1 - l0)
        owned[l0 + lo:l0 + hi + 1] += 1
    assert np.all(owned == 1)
    out, res, st = oracle.newton_mt(probs, us, opts)
    assert st == 0 and all(r.status == 0 for r in res)
    assert len({r.iterations for r in res}) == 1 and len({r.lin_iterations for r in res}) == 1
    glob = _gather(out, ss, nx, s1[0]["ny"] + 1)
    scale = np.max(np.abs(ref.reshape(-1, 4)), axis=0)
    err = np.max(np.abs(glob - ref).reshape(-1, 4) / scale, axis=0)
    assert np.all(err[:3] < 1e-8) and err[3] < 2e-6, err
    # overlap copies are consistent between neighbours after the solve
    for a in range(nranks - 1):
        ua = out[a].reshape(s1[0]["ny"] + 1, -1, 4)
        ub = out[a + 1].reshape(s1[0]["ny"] + 1, -1, 4)
        assert np.max(np.abs(ua[:, -3:, :] - ub[:, :3, :])) < 1e-9 * np.max(scale)


def test_gloo_world_size_2(oracle, axlib, tmp_path):
    """Two processes over torch.distributed (gloo): halo add via send/recv, dots via all_reduce.
    Must reproduce the in-process two-rank run bit for bit (same reduction order)."""
    nx = 8
    opts = oracle.default_newton_opts(reduction=1e-8, abs_limit=1e-14, min_lin_reduction=1e-9, ilu_fill=1)
    probs, us, ss = _rank_problems(oracle, nx, 2)
    out, res, st = oracle.newton_mt(probs, us, opts)
    assert st == 0
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", PYTHONPATH=ROOT + ":" + os.path.join(ROOT, "tests"))
    outs = [str(tmp_path / ("u%d.npy" % r)) for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "gloo_rank.py"), str(r), "2", str(nx), outs[r]],
                              env=env) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    for r in range(2):
        got = np.load(outs[r])
        assert np.array_equal(got, out[r]), np.max(np.abs(got - out[r]))
