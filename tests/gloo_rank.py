"""One rank of the two-process gloo test (tests/test_multirank_cpu.py): the oracle's overlapping
Newton solve with its collectives served by torch.distributed (gloo) instead of threads."""
import ctypes as C
import sys

import numpy as np
import torch
import torch.distributed as dist

from dune_ax1_b200 import setups
from helpers import oracle_problem
from oracle import pyoracle as O


def main():
    rank, world, nx, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = setups.make_setup(nx, dx=100.0e3, rank=rank, nranks=world, perturbation=1e-3)
    P = oracle_problem(O, s)
    P.set_time(0.0, 1.0); P.set_uold(s["u0"]); P.update_state(s["u0"])
    nvx, nvy = s["nx"] + 1, s["ny"] + 1

    def cols(v, c0):
        return v.reshape(nvy, nvx, 4)[:, c0:c0 + 3, :]

    def halo_add(ctx, rk, vptr):
        v = np.ctypeslib.as_array(vptr, shape=(4 * nvx * nvy,))
        reqs, recv = [], {}
        for side, nb, c0 in (("L", rank - 1, 0), ("R", rank + 1, nvx - 3)):
            if nb < 0 or nb >= world:
                continue
            send = torch.from_numpy(np.ascontiguousarray(cols(v, c0)))
            recv[side] = (torch.empty_like(send), c0)
            reqs.append(dist.isend(send, nb))
            reqs.append(dist.irecv(recv[side][0], nb))
        for r in reqs:
            r.wait()
        for side, (t, c0) in recv.items():
            cols(v, c0)[...] += t.numpy()

    def allreduce(ctx, rk, local):
        # gather + ordered sum = the same rank-order reduction as the in-process world
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([local], dtype=torch.float64))
        s_ = 0.0
        for p in parts:
            s_ += float(p[0])
        return s_

    opts = O.default_newton_opts(reduction=1e-8, abs_limit=1e-14, min_lin_reduction=1e-9, ilu_fill=1)
    u, res = P.newton_ext(s["u0"], opts, rank, halo_add, allreduce)
    assert res.status == 0, res.status
    np.save(out, u)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
