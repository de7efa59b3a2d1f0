"""Launch the multi-GPU parity script when the box has >= 2 GPUs (gpurun --gpus 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_gpu_parity(p2p):
    """p2p=1: halo sum and dot products through NVLink peer windows (p2p.cuh); p2p=0: NCCL send/recv + allreduce."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "run_multigpu_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, AX1_P2P=p2p))
    sys.stdout.write(out.stdout[-3000:])
    sys.stderr.write(out.stderr[-3000:])
    assert out.returncode == 0 and "MULTIGPU PARITY OK" in out.stdout


@pytest.mark.parametrize("compact", ["auto", "1"])
def test_two_ranks_on_one_gpu_peer_windows(compact):
    """(compact = "1": the solver's products on the 64-byte-block operand, forced by AX1_SPMV_COMPACT -- the operand of
    the named workload's ranks, with the constrained processor-boundary rows of the overlap.)
    The same parity script with BOTH ranks on GPU 0 (two processes, time-sliced contexts) and the peer windows
    bootstrapped without NCCL (ax1gpu_p2p_export / ax1gpu_p2p_attach over gloo): the overlap path of row a17 -- halo sum
    pushed by the sweep kernels into the neighbour's window, owner-masked dot products summed in rank order through the
    windows, constrained processor-boundary vertices -- runs on a 1-GPU box. Skipped (not failed) if the two contexts
    are not scheduled against each other within the exchange timeout."""
    import torch
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "run_multigpu_parity.py")]
    env = dict(os.environ, AX1_SAME_GPU="1", AX1_P2P="1", AX1_P2P_TIMEOUT_S="20")
    if compact == "1":
        env["AX1_SPMV_COMPACT"] = "1"
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env)
    except subprocess.TimeoutExpired:
        pytest.skip("two contexts on one GPU did not make progress against each other")
    sys.stdout.write(out.stdout[-3000:])
    sys.stderr.write(out.stderr[-2000:])
    if out.returncode != 0 and ("timed out" in out.stdout or "timed out" in out.stderr):
        pytest.skip("peer-window exchange timed out with both ranks on one GPU (no time slicing between the contexts)")
    assert out.returncode == 0 and "MULTIGPU PARITY OK" in out.stdout
    if compact == "1":
        assert "spmv operand compact True" in out.stdout
