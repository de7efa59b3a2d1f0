"""Row f3, fidelity of the multigrid preconditioner: the x-aggregation cycle the device runs
(csrc/amg.cu; restated in oracle/gmres_amg.cpp) held against an INDEPENDENT generic algebraic aggregation multigrid
(tests/generic_amg.py: strength of connection from the matrix, greedy aggregation with Dune::Amg's anisotropic default
sizes, skip-isolated, non-smoothed prolongation, Galerkin operators, block Gauss-Seidel smoothing, direct coarse solve;
numpy / scipy only, no line of the oracle or of the device design). On the problem class where one-level ILU stalls
-- the axially refined grid, dx = 100 nm -- the same restarted GMRES(100) is run with (a) block ILU0 alone, (b) the
device's cycle, (c) the generic algebraic cycle, and iterations to a TRUE residual reduction are compared. The generic
aggregates follow the strong couplings the matrix shows (on this grid mostly axial, radial in the Debye layers); the
device's aggregates are cx whole vertex columns and leave the radial direction to the ILU smoother: two coarsening
philosophies that must land in the same convergence class. CPU only."""
import numpy as np
import scipy.sparse.linalg as sla

from generic_amg import GenericAMG, aggregate, block_strength, bsr_from_bcrs
from helpers import oracle_problem, small_setup


def _count(A, M, r, red):
    n = len(r)
    MA = sla.LinearOperator((n, n), matvec=lambda x: M(A @ x))
    nr = np.linalg.norm(r)
    for m in range(5, 400, 5):          # smallest m (step 5) whose m-step GMRES iterate reaches the true reduction
        z, _ = sla.gmres(MA, M(r), rtol=1e-30, atol=0.0, restart=m, maxiter=1)
        if np.linalg.norm(r - A @ z) <= red * nr:
            return m
    return 400


def test_device_cycle_is_in_the_convergence_class_of_a_generic_algebraic_amg(oracle):
    s = small_setup(nx=64, dx=100.0)
    P = oracle_problem(oracle, s)
    u0 = s["u0"]
    P.set_time(0.0, 1.0); P.set_uold(u0); P.update_state(u0)
    J, r = P.jacobian(u0), P.residual(u0)
    rowptr, col = P.pattern()
    A = bsr_from_bcrs(rowptr, col, J)
    assert np.max(np.abs(A @ u0 - P.spmv(J, u0))) <= 1e-12 * np.max(np.abs(P.spmv(J, u0)))
    G = GenericAMG(A)
    sizes = G.sizes()
    assert len(sizes) >= 2 and sizes[1] < 0.4 * sizes[0]          # it really coarsens (aggregates of ~4 vertices)
    # where does a purely algebraic aggregation coarsen on this grid? Mostly ALONG the axon: dx = 100 nm is far
    # smaller than the radial spacing everywhere but in the Debye layers, so the strong couplings run in x -- the
    # direction the device's structured aggregates (cx consecutive vertex columns) take by construction
    agg, nagg = aggregate(block_strength(A, P.nv))
    nvx = s["nx"] + 1
    spans = np.zeros((nagg, 4))
    spans[:, 0] = spans[:, 2] = 1e9
    for v in np.nonzero(agg >= 0)[0]:
        i, j, a = v % nvx, v // nvx, agg[v]
        spans[a] = [min(spans[a, 0], i), max(spans[a, 1], i), min(spans[a, 2], j), max(spans[a, 3], j)]
    wide = np.mean((spans[:, 1] - spans[:, 0]) >= (spans[:, 3] - spans[:, 2]))
    assert wide > 0.6, wide
    red = 1e-6
    its_ilu = _count(A, lambda d: P.ilu_apply(J, d, 0, 16), r, red)
    its_dev = _count(A, lambda d: P.amg_apply(J, d, 0, 16, 4), r, red)
    its_gen = _count(A, lambda d: G.cycle(d), r, red)
    print("GMRES iterations to a true reduction of 1e-6: ILU0 %d, device cycle (cx = 4) %d, generic algebraic AMG %d"
          % (its_ilu, its_dev, its_gen))
    assert its_gen < its_ilu and its_dev < its_ilu                 # both multigrids beat the one-level ILU
    assert its_dev <= 2 * its_gen and its_gen <= 2 * its_dev       # ... and sit in the same class
