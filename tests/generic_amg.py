"""An independent, generic aggregation multigrid in numpy / scipy -- TEST INFRASTRUCTURE.

Written to hold the device's multigrid cycle (csrc/amg.cu: aggregates = cx consecutive vertex columns, chosen from
the grid's structure) against what a purely ALGEBRAIC aggregation in the style of Dune::Amg would do on the same
matrix, the preconditioner the reference's myelin build configures (acme2_cyl_setup.hh:749-760,
ax1_solverbackend.hh:73-302): strength of connection from the matrix alone (symmetric criterion, alpha = 1/3, on the
Frobenius norm of the 4 x 4 blocks), greedy aggregation of strongly connected neighbours with the anisotropic default
sizes (min 4 / max 6 vertices, graph distance <= 3 from the seed; dune-istl Parameters::setDefaultValuesAnisotropic(2)),
isolated vertices skipped (setSkipIsolated), piecewise-constant prolongation per component (non-smoothed aggregation),
Galerkin coarse operators, V-cycles with a block (4 x 4) symmetric Gauss-Seidel smoother, direct solve below the coarsen
target. Nothing here knows about x-columns, the stencil or the device's data layout, and no line of the oracle is used.
It is NOT a bit-level restatement of dune-istl's aggregation heuristics (un-vendored, see DESIGN); it answers the
question how iteration counts of the device's structured cycle relate to a generic algebraic one."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla


def bsr_from_bcrs(rowptr, col, vals):
    return sp.bsr_matrix((vals.reshape(-1, 4, 4), col, rowptr)).tocsr()


def block_strength(A, nb, bs=4):
    """Frobenius norms of the bs x bs blocks as a CSR graph on the block rows"""
    Ab = sp.bsr_matrix(A, blocksize=(bs, bs))
    data = np.sqrt((Ab.data ** 2).sum(axis=(1, 2)))
    return sp.csr_matrix((data, Ab.indices, Ab.indptr), shape=(nb, nb))


def aggregate(S, alpha=1.0 / 3.0, min_size=4, max_size=6, max_dist=3):
    """greedy aggregation on the strength graph S (symmetric criterion); returns agg[v] (-1: isolated / skipped)"""
    n = S.shape[0]
    St = S.T.tocsr()
    d = S.diagonal()
    strong = [None] * n
    for i in range(n):
        cols = S.indices[S.indptr[i]:S.indptr[i + 1]]
        vals = S.data[S.indptr[i]:S.indptr[i + 1]]
        best, cand = 0.0, []
        for j, aij in zip(cols, vals):
            if j == i:
                continue
            aji = St[i, j]
            m = aij * aji / (d[i] * d[j]) if d[i] > 0 and d[j] > 0 else 0.0
            cand.append((m, j))
            best = max(best, m)
        strong[i] = [(m, j) for m, j in cand if best > 0 and m >= alpha * best]
    agg = -np.ones(n, dtype=np.int64)
    nagg = 0
    for seed in range(n):
        if agg[seed] >= 0 or not strong[seed]:
            continue
        members, dist = [seed], {seed: 0}
        agg[seed] = nagg
        frontier = [seed]
        while frontier and len(members) < max_size:
            # strongest unaggregated neighbour of the current aggregate within the distance bound
            cands = [(m, j, v) for v in members for m, j in strong[v] if agg[j] < 0 and dist[v] + 1 <= max_dist]
            if not cands:
                break
            m, j, v = max(cands)
            agg[j] = nagg
            dist[j] = dist[v] + 1
            members.append(j)
            if len(members) >= min_size and m < 0.5 * max(c[0] for c in cands):
                break
        nagg += 1
    # undersized aggregates (1 member) join their strongest aggregated neighbour
    sizes = np.bincount(agg[agg >= 0], minlength=nagg)
    for v in range(n):
        if agg[v] >= 0 and sizes[agg[v]] == 1 and strong[v]:
            m, j = max(strong[v])
            if agg[j] >= 0 and agg[j] != agg[v]:
                sizes[agg[v]] -= 1
                agg[v] = agg[j]
                sizes[agg[j]] += 1
    # renumber
    used = np.unique(agg[agg >= 0])
    ren = -np.ones(nagg, dtype=np.int64)
    ren[used] = np.arange(len(used))
    agg[agg >= 0] = ren[agg[agg >= 0]]
    return agg, len(used)


class GenericAMG:
    def __init__(self, A, bs=4, coarsen_target=2000, max_levels=15, nu=1, damp=1.0):
        self.levels, self.nu, self.damp = [], nu, damp
        A = A.tocsr()
        while True:
            n = A.shape[0]
            lev = {"A": A}
            if n // bs <= coarsen_target or len(self.levels) + 1 >= max_levels:
                lev["lu"] = sla.splu(A.tocsc())
                self.levels.append(lev)
                break
            nb = n // bs
            # block symmetric Gauss-Seidel: exact solves with the block lower / block upper triangle (incl. the 4 x 4
            # diagonal blocks), which scipy's sparse LU factorises without fill beyond the blocks
            blk = np.arange(n) // bs
            coo = A.tocoo()
            low = sp.csc_matrix((coo.data[blk[coo.row] >= blk[coo.col]], (coo.row[blk[coo.row] >= blk[coo.col]],
                                                                          coo.col[blk[coo.row] >= blk[coo.col]])), shape=A.shape)
            upp = sp.csc_matrix((coo.data[blk[coo.row] <= blk[coo.col]], (coo.row[blk[coo.row] <= blk[coo.col]],
                                                                          coo.col[blk[coo.row] <= blk[coo.col]])), shape=A.shape)
            lev["low"] = sla.splu(low, permc_spec="NATURAL", diag_pivot_thresh=0.0)
            lev["upp"] = sla.splu(upp, permc_spec="NATURAL", diag_pivot_thresh=0.0)
            agg, nagg = aggregate(block_strength(A, nb, bs))
            rows = np.nonzero(agg >= 0)[0]
            Pb = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(nb, nagg))
            P = sp.kron(Pb, sp.identity(bs), format="csr")
            lev["P"] = P
            self.levels.append(lev)
            Ac = (P.T @ A @ P).tocsr()
            if nagg >= 0.9 * nb:       # coarsening stalled
                self.levels.append({"A": Ac, "lu": sla.splu(Ac.tocsc())})
                break
            A = Ac

    def sizes(self):
        return [l["A"].shape[0] // 4 for l in self.levels]

    def cycle(self, b, l=0):
        lev = self.levels[l]
        if "lu" in lev:
            return lev["lu"].solve(b)
        A = lev["A"]
        x = np.zeros_like(b)
        for _ in range(self.nu):
            x += lev["low"].solve(b - A @ x)
        rc = lev["P"].T @ (b - A @ x)
        x += self.damp * (lev["P"] @ self.cycle(rc, l + 1))
        for _ in range(self.nu):
            x += lev["upp"].solve(b - A @ x)
        return x
