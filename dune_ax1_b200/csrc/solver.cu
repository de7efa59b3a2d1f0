// solver.cu -- linear-algebra kernels on the 9-slot block-row layout (4x4 FP64 blocks):
//   SpMV                       replaces BCRSMatrix::mv inside Dune::BiCGSTABSolver       [ext dune-istl]
//   block ILU0 factor / apply  replaces Dune::SeqILUn / bilu0_decomposition / bilu_backsolve
//                              (selected at acme2_cyl_setup.hh:764-765, 797-798)
//   row-norm equilibration     dune/ax1/common/rownorm_preconditioner.hh:65-155
//
// ILU parallelisation: the reference runs one ILU per MPI rank (additive Schwarz over x-slabs,
// ax1_solverbackend.hh:346-381). Here every GPU holds many x-sub-slabs of `slab_w` vertex columns;
// one CTA factors / solves one sub-slab, sweeping the wavefronts level = i + 2 j of the x-fastest
// 9-point stencil (row (i,j) depends on (i-1,j-1),(i,j-1),(i+1,j-1),(i-1,j)). Couplings across
// sub-slab borders are dropped from the preconditioner only (the operator keeps them).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "ilu_common.cuh"
#include "p2p.cuh"
#include "reduce.cuh"

namespace ax1 {

// ------------------------------------------------------------------------------------------
// y = A x ; 4 lanes per block row, lane r = scalar row r.
// DOT = 3: residual mode y = w - A x (multigrid defect), no reduction.
// DOT = 1: also partial <w, y> (BiCGStab h = <rt, v>);  DOT = 2: partial <y, w> and <y, y>
// (omega = <t, r> / <t, t>): the SpMV result is reduced while it is still in registers.
template <int DOT>
__global__ void __launch_bounds__(256) spmv_kernel(Dev g, const double* __restrict__ A, const double* __restrict__ x,
                                                   double* __restrict__ y, int zero_con, const double* __restrict__ w,
                                                   Own own, const Scalars* sc, double* partials) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(t >> 2), r = (int)(t & 3);
  double s[2] = {0.0, 0.0};
  const bool active = (v < g.nv) && !(DOT && DOT != 3 && sc && sc->done);
  if (active) {
    const int i = v % g.nvx, j = v / g.nvx;
    const double* Arow = A + (size_t)v * AX1_ROW;
    double acc = 0.0;
#pragma unroll
    for (int sl = 0; sl < 9; ++sl) {
      const int di = sl % 3 - 1, dj = sl / 3 - 1;
      const int ii = i + di, jj = j + dj;
      if (ii < 0 || ii >= g.nvx || jj < 0 || jj >= g.nvy) continue;
      double xv[4];
      ldg4(x + 4 * (size_t)(v + di + dj * g.nvx), xv);
      // branch-free arrow-row load: every lane issues two 16 B loads (lanes 0-2 read their (d,c) pair
      // twice) so that all 36 loads of the row can be in flight at once
      const double* blk = Arow + sl * AX1_BLK;
      const double2 a = __ldg(reinterpret_cast<const double2*>(blk + (r == 3 ? 6 : 2 * r)));
      const double2 b2 = __ldg(reinterpret_cast<const double2*>(blk + (r == 3 ? 8 : 2 * r)));
      const double xa = r == 3 ? xv[0] : (r == 0 ? xv[0] : (r == 1 ? xv[1] : xv[2]));
      const double xb = r == 3 ? xv[1] : xv[3];
      acc += a.x * xa;
      acc += a.y * xb;
      acc += (r == 3 ? b2.x : 0.0) * xv[2];
      acc += (r == 3 ? b2.y : 0.0) * xv[3];
    }
    if (DOT == 3) acc = __ldg(w + 4 * (size_t)v + r) - acc;
    if (zero_con && ((g.dmask[v] >> r) & 1)) acc = 0.0;
    y[4 * (size_t)v + r] = acc;
    if (DOT == 1 || DOT == 2) {
      if (!own.masked || (i >= own.lo && i <= own.hi)) {
        const double wv = __ldg(w + 4 * (size_t)v + r);
        s[0] = acc * wv;
        if (DOT == 2) s[1] = acc * acc;
      }
    }
  }
  if (DOT == 1 || DOT == 2) block_reduce_store<2>(s, partials);
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lev_jlo(int L, int wc) {
  const int a = L - wc + 1;
  return a <= 0 ? 0 : (a + 1) / 2;
}

// ------------------------------------------------------------------------------------------
// Wavefront-packed factor layout. The factor is private to the preconditioner, so it is stored in
// the order the sweeps consume it: sub-slab major, then level L = il + 2 j, then row within the
// level (q = j - jlo(L)). LF holds the 4 lower blocks of a row (512 B), LB the inverted diagonal
// block followed by the 4 upper blocks (640 B). A CTA therefore streams two contiguous arrays
// instead of touching Ny+1 pages that lie (Nx+1)*1152 B apart at every level.
//   row position  p(il, j) = lvl[L] + j - jlo(L),  lvl = prefix sum of the level sizes
//   LF[(slab_off + p) * 64 + s * 16 + r * 4 + c],  LB[(slab_off + p) * 80 + t * 16 + r * 4 + c]
struct SlabGeom {
  int i0, wc, nlev;
  size_t row_off;   // packed row offset of this sub-slab
  const int* lvl;   // level prefix table (shared memory)
};
__device__ __forceinline__ int packed_pos(const SlabGeom& sg, int il, int j) {
  const int L = il + 2 * j;
  return sg.lvl[L] + j - lev_jlo(L, sg.wc);
}
__device__ __forceinline__ SlabGeom slab_geom(const Dev& g, int slab_w, const int* __restrict__ lvl_tab, int* lvl_sh) {
  SlabGeom sg;
  sg.i0 = blockIdx.x * slab_w;
  sg.wc = min(slab_w, g.nvx - sg.i0);
  sg.nlev = sg.wc + 2 * (g.nvy - 1);
  sg.row_off = (size_t)blockIdx.x * slab_w * g.nvy;
  const int stride = slab_w + 2 * (g.nvy - 1) + 1;
  const int* src = lvl_tab + (sg.wc == slab_w ? 0 : stride);
  for (int k = threadIdx.x; k <= sg.nlev; k += blockDim.x) lvl_sh[k] = src[k];
  __syncthreads();
  sg.lvl = lvl_sh;
  return sg;
}

struct LevelRow {
  int v, il, j;  // v < 0: this quad has no row at this level
};
__device__ __forceinline__ LevelRow level_row(const Dev& g, int L, int nlev, int i0, int wc, int q) {
  LevelRow o;
  o.v = -1; o.il = 0; o.j = 0;
  if (L < 0 || L >= nlev) return o;
  const int j = lev_jlo(L, wc) + q;
  if (j > min(g.nvy - 1, L >> 1)) return o;
  o.j = j; o.il = L - 2 * j;
  o.v = i0 + o.il + g.nvx * j;
  return o;
}

// Radial split of the sub-slabs (jc > 0; additive Schwarz in y as well): blockIdx.y = 0 handles rows
// [0, jc), 1 rows [jc, nvy) of its sub-slab as an independent grid -- couplings across the cut are dropped from
// the preconditioner. The kernels below then simply see a grid of g.nvy = the part's rows: vertex-indexed arrays
// are shifted to the part's first row, the packed factor stores all lower parts first and all upper parts after
// them, and the level tables hold two tables (full / last sub-slab width) per part. Placed in the uniform far
// field of the radial grid the cut costs no Krylov iterations and halves the dependency depth of a sweep
// (w + 2 Ny wavefronts -> w + Ny).
struct YPart { int j0, nvy_glob, lvl_off; };
__device__ __forceinline__ YPart ypart_select(Dev& g, int jc, int slab_w) {
  YPart yp{0, g.nvy, 0};
  if (jc > 0) {
    if (blockIdx.y) { yp.lvl_off = 2 * (slab_w + 2 * (jc - 1) + 1); yp.j0 = jc; g.nvy -= jc; }
    else g.nvy = jc;
  }
  return yp;
}

// block ILU0 of one sub-slab per CTA (bilu0_decomposition: L_ij = A_ij * A_jj^-1, diagonal stored
// inverted). One quad per wavefront row. The 720 B Jacobian row of level L+2 is streamed into shared
// memory with cp.async while level L is eliminated; the finished [Dinv | U] rows of the last three
// levels -- all a row ever needs from its lower neighbours -- are kept in a shared-memory ring, so
// the critical path of a level contains no global load.
#define AX1_RING_ROW 82  // 80 doubles + 2 padding: 8 quads of a warp hit 8 different bank groups
__global__ void __launch_bounds__(256, 1) ilu_factor_kernel(Dev g, const double* __restrict__ A, double* LF, double* LB,
                                                            int slab_w, const int* __restrict__ lvl_tab, int jc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  {
    const YPart yp = ypart_select(g, jc, slab_w);
    const size_t roff = (size_t)yp.j0 * g.nvx;
    A += roff * AX1_ROW; LF += roff * 64; LB += roff * 80; lvl_tab += yp.lvl_off;
  }
  const int tid = threadIdx.x, q = tid >> 2, r = tid & 3, nq = blockDim.x >> 2;
  double* stage = reinterpret_cast<double*>(smem_raw);            // [2][nq][AX1_ROW]
  double* ring = stage + 2 * (size_t)nq * AX1_ROW;                 // [3][nq][AX1_RING_ROW]
  const SlabGeom sg = slab_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 3 * (size_t)nq * AX1_RING_ROW));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  for (int k = tid; k < 3 * nq * AX1_RING_ROW; k += blockDim.x) ring[k] = 0.0;  // finite don't-care rows (see below)
  auto issue = [&](int L) {
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const double* src = A + (size_t)w.v * AX1_ROW;
      double* dst = stage + ((size_t)(L & 1) * nq + q) * AX1_ROW;
      for (int c = r; c < AX1_ROW / 2; c += 4) cp_async16(dst + 2 * c, src + 2 * c);
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  for (int L = 0; L < nlev; ++L) {
    cp_async_wait<1>();
    cta_sync();  // staged row of this level and the ring rows of level L-1 are visible
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    const bool act = w.v >= 0;
    double row[9][4];
    row[4][0] = row[4][1] = row[4][2] = row[4][3] = 0.0;
    if (act) {
      const int il = w.il, j = w.j;
      bool ex[9];
      const double* arow = stage + ((size_t)(L & 1) * nq + q) * AX1_ROW;
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        ex[s] = (il + di >= 0) && (il + di < wc) && (j + dj >= 0) && (j + dj < g.nvy);
        // expand the arrow block row (common.cuh), branch-free in the lane index
        const double* blk = arow + s * AX1_BLK;
        const double2 a = *reinterpret_cast<const double2*>(blk + (r == 3 ? 6 : 2 * r));
        const double2 b2 = *reinterpret_cast<const double2*>(blk + (r == 3 ? 8 : 2 * r));
        row[s][0] = (ex[s] && (r == 3 || r == 0)) ? a.x : 0.0;
        row[s][1] = ex[s] ? (r == 3 ? a.y : (r == 1 ? a.x : 0.0)) : 0.0;
        row[s][2] = ex[s] ? (r == 3 ? b2.x : (r == 2 ? a.x : 0.0)) : 0.0;
        row[s][3] = ex[s] ? (r == 3 ? b2.y : a.y) : 0.0;
      }
      // Branch-free elimination: a lower neighbour that does not exist has a zero block row (expanded above), so its
      // L row is 0 x (any finite ring row) = 0 and every update it feeds subtracts 0; and a block U_k[tt] that points at
      // a vertex which does not exist is zero in k's row as well, so targets outside the grid / sub-slab stay zero.
      // Quads of a warp differ in which neighbours exist on almost every level; predication would serialise them.
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int dis = s % 3 - 1, djs = s / 3 - 1;
        const int Ln = L + dis + 2 * djs;
        const int qn = (j + djs) - lev_jlo(Ln, wc);
        const double* kb = ex[s] ? ring + ((size_t)(Ln % 3) * nq + qn) * AX1_RING_ROW : ring;  // row k: [Dinv, U5..U8]
        double lrow[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double sum = 0.0;
#pragma unroll
          for (int m = 0; m < 4; ++m) sum += row[s][m] * kb[4 * m + c];
          lrow[c] = sum;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) row[s][c] = lrow[c];
#pragma unroll
        for (int tt = 5; tt < 9; ++tt) {
          const int di = dis + (tt % 3 - 1), dj = djs + (tt / 3 - 1);
          if (di < -1 || di > 1 || dj < -1 || dj > 1) continue;
          const int s2 = (dj + 1) * 3 + (di + 1);
          const double* ub = kb + (tt - 4) * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum += lrow[m] * ub[4 * m + c];
            row[s2][c] -= sum;
          }
        }
      }
    }
    // invert the pivot block (every lane of the quad holds one row of it). The exchange runs converged with the
    // full mask -- quads without a row on this level shuffle don't-care values -- so it compiles to bare SHFLs
    // instead of one collective re-convergence sequence per shuffle
    double d[4][4], dinv[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c) d[m][c] = __shfl_sync(0xffffffffu, row[4][c], (tid & 28) + m);
    if (act) {
      invert4(d, dinv);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double val = dinv[0][c];
        if (r == 1) val = dinv[1][c];
        if (r == 2) val = dinv[2][c];
        if (r == 3) val = dinv[3][c];
        row[4][c] = val;
      }
      const size_t p = sg.row_off + (size_t)(sg.lvl[L] + q);
#pragma unroll
      for (int s = 0; s < 4; ++s) st4(LF + p * 64 + s * 16 + r * 4, row[s]);
#pragma unroll
      for (int s = 4; s < 9; ++s) st4(LB + p * 80 + (s - 4) * 16 + r * 4, row[s]);
    }
    cta_sync();  // every read of ring slot L%3 (it held level L-3) is done
    if (w.v >= 0) {
      double* kr = ring + ((size_t)(L % 3) * nq + q) * AX1_RING_ROW;
#pragma unroll
      for (int s = 4; s < 9; ++s) st4(kr + (s - 4) * 16 + r * 4, row[s]);
    }
    issue(L + 2);
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// The same factorisation for narrow sub-slabs (<= 24 vertex columns: the paper grid and the coarse multigrid
// levels), where a level holds <= 12 rows and the sweep is bound by the dependent-instruction latency
// of a single warp rather than by memory. Sixteen lanes share a vertex row -- lane (r, c) owns element (r, c) of
// each of its 9 blocks -- so a level costs every thread a quarter of the multiply-adds, the pivot block is inverted
// by its 16 owners together (adjugate formula, ilu_common.cuh: equal to the quad kernel's pivoted Gauss-Jordan to
// rounding), and a 16-column sub-slab
// keeps four warp schedulers busy instead of one. Rows of a block row are exchanged with full-mask shuffles; the code
// is uniform across the CTA (rows that do not exist on a level run on zero blocks and store nothing).
#define AX1_RING16_ROW 84  // 80 doubles + 4: the two row groups of a warp read disjoint banks
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// Levels are software-pipelined: of the four lower neighbours of a row on level L, SW and S lie on levels L-3 and
// L-2 and only SE and W on level L-1. Rows are eliminated in increasing column order SW, S, SE, W, so the first two
// steps of level L+1 can run while level L still does its last two steps and inverts its pivot block -- two
// independent dependency chains per thread instead of one, and a single barrier per level (the ring slot written at
// the end of iteration L was last read in iteration L-1). Operation order per row is unchanged.
__global__ void __launch_bounds__(256) ilu_factor16_kernel(Dev g, const double* __restrict__ A, double* LF, double* LB,
                                                           int slab_w, int nq, const int* __restrict__ lvl_tab, int jc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  {
    const YPart yp = ypart_select(g, jc, slab_w);
    const size_t roff = (size_t)yp.j0 * g.nvx;
    A += roff * AX1_ROW; LF += roff * 64; LB += roff * 80; lvl_tab += yp.lvl_off;
  }
  const int tid = threadIdx.x, q = tid >> 4, e = tid & 15, r = e >> 2, c = e & 3;
  const int lane = tid & 31, rowl = lane & ~3, grp = lane & 16;  // first lane of my block row / of my vertex row
  double* stage = reinterpret_cast<double*>(smem_raw);            // [3][nq][AX1_ROW]
  double* ring = stage + 3 * (size_t)nq * AX1_ROW;                 // [3][nq][AX1_RING16_ROW]
  const SlabGeom sg = slab_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 3 * (size_t)nq * AX1_RING16_ROW + 2 * (size_t)nq * 16));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  for (int k = tid; k < 3 * nq * AX1_RING16_ROW; k += blockDim.x) ring[k] = 0.0;  // finite don't-care rows
  double* pscratch = ring + 3 * (size_t)nq * AX1_RING16_ROW + (size_t)q * 16;  // [2][nq][16]: pivot block of my row, by level parity
  const Adj16 adj = adj16_setup(r, c);
  // element (r, c) inside the 10-double arrow block (common.cuh): rows 0..2 hold (r,r), (r,3); row 3 is full
  const bool nz = (r == 3) || (c == 3) || (r == c);
  const int aoff = (r == 3) ? 6 + c : (c == 3 ? 2 * r + 1 : 2 * r);

  auto issue = [&](int L) {
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const double* src = A + (size_t)w.v * AX1_ROW;
      double* dst = stage + ((size_t)(L % 3) * nq + q) * AX1_ROW;
      for (int k = e; k < AX1_ROW / 2; k += 16) cp_async16(dst + 2 * k, src + 2 * k);
    }
    cp_async_commit();
  };
  // which neighbours of the row this quad-of-quads holds on level L exist (bit s), 0 if it holds none
  auto exists = [&](const LevelRow& w) {
    unsigned m = 0;
    if (w.v >= 0) {
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0) && (w.j + dj < g.nvy)) m |= 1u << s;
      }
    }
    return m;
  };
  // one elimination step of row x (level L, grid row j) against its lower neighbour s; branch-free: a missing
  // neighbour is a zero block row, so its L row and every update it feeds are zero
  auto eliminate = [&](double (&x)[9], int s_, int L, int j, unsigned exm) {
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      if (s != s_) continue;
      const int dis = s % 3 - 1, djs = s / 3 - 1;
      const int Ln = L + dis + 2 * djs;
      const int qn = (j + djs) - lev_jlo(Ln, wc);
      const double* kb = ((exm >> s) & 1u) ? ring + ((size_t)(Ln % 3) * nq + qn) * AX1_RING16_ROW : ring;  // [Dinv, U5..U8]
      double xs[4], lr[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) xs[m] = shfl_d(x[s], rowl + m);
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) sum += xs[m] * kb[4 * m + c];
      x[s] = sum;
#pragma unroll
      for (int m = 0; m < 4; ++m) lr[m] = shfl_d(sum, rowl + m);
#pragma unroll
      for (int tt = 5; tt < 9; ++tt) {
        const int di = dis + (tt % 3 - 1), dj = djs + (tt / 3 - 1);
        if (di < -1 || di > 1 || dj < -1 || dj > 1) continue;
        const int s2 = (dj + 1) * 3 + (di + 1);
        const double* ub = kb + (tt - 4) * 16;
        double su = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) su += lr[m] * ub[4 * m + c];
        x[s2] -= su;
      }
    }
  };
  auto expand = [&](double (&x)[9], int L, unsigned exm) {
    const double* arow = stage + ((size_t)(L % 3) * nq + q) * AX1_ROW;
#pragma unroll
    for (int s = 0; s < 9; ++s) x[s] = (((exm >> s) & 1u) && nz) ? arow[s * AX1_BLK + aoff] : 0.0;
  };

  issue(0);
  issue(1);
  issue(2);
  double xn[9];
  LevelRow wn = level_row(g, 0, nlev, i0, wc, q);
  unsigned exn = exists(wn);
  cp_async_wait<2>();
  cta_sync();  // row of level 0 staged, ring zeroed
  expand(xn, 0, exn);
  eliminate(xn, 0, 0, wn.j, exn);
  eliminate(xn, 1, 0, wn.j, exn);
  for (int L = 0; L < nlev; ++L) {
    double x[9];
#pragma unroll
    for (int s = 0; s < 9; ++s) x[s] = xn[s];
    const LevelRow w = wn;
    const unsigned exm = exn;
    const bool act = w.v >= 0;
    cp_async_wait<1>();
    cta_sync();  // row of level L+1 staged; ring row of level L-1 visible
    wn = level_row(g, L + 1, nlev, i0, wc, q);
    exn = exists(wn);
    // level L: SE, W          level L+1: load, SW, S
    eliminate(x, 2, L, w.j, exm);
    expand(xn, L + 1, exn);
    eliminate(x, 3, L, w.j, exm);
    eliminate(xn, 0, L + 1, wn.j, exn);
    // pivot block of level L: adjugate inverse over its 16 owners (ilu_common.cuh); the last early step of level L+1
    // is independent of it
    const double a = act ? x[4] : ((r == c) ? 1.0 : 0.0);
    eliminate(xn, 1, L + 1, wn.j, exn);
    const double inv = invert4_adjugate16(a, pscratch + (size_t)(L & 1) * nq * 16, e, adj);
    x[4] = inv;
    if (act) {
      const size_t p = sg.row_off + (size_t)(sg.lvl[L] + q);
#pragma unroll
      for (int s = 0; s < 4; ++s) LF[p * 64 + s * 16 + e] = x[s];
#pragma unroll
      for (int s = 4; s < 9; ++s) LB[p * 80 + (s - 4) * 16 + e] = x[s];
      double* kr = ring + ((size_t)(L % 3) * nq + q) * AX1_RING16_ROW;  // slot last read in iteration L-1
#pragma unroll
      for (int s = 4; s < 9; ++s) kr[(s - 4) * 16 + e] = x[s];
    }
    issue(L + 3);
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// v = (LU)^-1 d for one sub-slab per CTA (bilu_backsolve): forward sweep with unit lower blocks,
// backward sweep with the stored inverse diagonal; both sweeps in one launch.
//
// The sweep is strictly sequential over ~W+2Ny wavefront levels, so HBM latency must be hidden by
// a deep software pipeline: every thread (lane r of the quad that owns a wavefront row) streams
// "its" scalar row of the 4-5 factor blocks of the next D levels into a private shared-memory slot
// with cp.async (LDGSTS) and consumes it D levels later -- no barrier is needed for the staged
// factor data, only cp.async.wait_group. The solution values of the last three levels, which the
// next level depends on, are exchanged through a small shared-memory ring (one __syncthreads per
// level) instead of an L2 round trip.
// P2P: multi-GPU run with peer windows -- the final values of the three vertex columns shared with a
// neighbour rank are also stored straight into that rank's memory over NVLink (p2p.cuh).
template <int D, bool P2P>
__global__ void __launch_bounds__(256) ilu_apply_kernel(Dev g, const double* __restrict__ LF,
                                                        const double* __restrict__ LB, const double* __restrict__ d,
                                                        double* v_out, int slab_w, const int* __restrict__ lvl_tab,
                                                        HaloPush hp, int jc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const YPart yp = ypart_select(g, jc, slab_w);
  {
    const size_t roff = (size_t)yp.j0 * g.nvx;
    LF += roff * 64; LB += roff * 80; d += 4 * roff; v_out += 4 * roff; lvl_tab += yp.lvl_off;
  }
  const int nt = blockDim.x, tid = threadIdx.x;
  const int q = tid >> 2, r = tid & 3, nq = nt >> 2;
  double* stageM = reinterpret_cast<double*>(smem_raw);   // [D][5][nt][4]
  double* stageR = stageM + (size_t)D * 5 * nt * 4;       // [D][nt]
  double* ring = stageR + (size_t)D * nt;                 // [4][nq][4]
  const SlabGeom sg = slab_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 4 * (size_t)nq * 4));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;

  // ---------------- forward: y = L^-1 d ----------------
  auto issue_fwd = [&](int L) {
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = L % D;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0)) {
          const double* src = LF + (sg.row_off + (size_t)(sg.lvl[L] + q)) * 64 + s * 16 + r * 4;
          double* dst = stageM + ((size_t)(st * 5 + s) * nt + tid) * 4;
          cp_async16(dst, src);
          cp_async16(dst + 2, src + 2);
        }
      }
      cp_async8(stageR + st * nt + tid, d + 4 * (size_t)w.v + r);
    }
    cp_async_commit();
  };
  for (int L = 0; L < D; ++L) issue_fwd(L);
  for (int L = 0; L < nlev; ++L) {
    cp_async_wait<D - 1>();
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = L % D;
      double acc = stageR[st * nt + tid];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0)) {
          const double* lp = stageM + ((size_t)(st * 5 + s) * nt + tid) * 4;
          const double2 l01 = *reinterpret_cast<const double2*>(lp);
          const double2 l23 = *reinterpret_cast<const double2*>(lp + 2);
          const int Ln = L + di + 2 * dj;
          const int qn = (w.j + dj) - lev_jlo(Ln, wc);
          const double* yp = ring + ((size_t)(Ln & 3) * nq + qn) * 4;
          const double2 y01 = *reinterpret_cast<const double2*>(yp);
          const double2 y23 = *reinterpret_cast<const double2*>(yp + 2);
          double t = l01.x * y01.x;
          t += l01.y * y01.y;
          t += l23.x * y23.x;
          t += l23.y * y23.y;
          acc -= t;
        }
      }
      ring[((size_t)(L & 3) * nq + q) * 4 + r] = acc;
      v_out[4 * (size_t)w.v + r] = acc;
    }
    cta_sync();
    issue_fwd(L + D);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---------------- backward: x = U^-1 y ----------------
  auto issue_bwd = [&](int L) {  // L counts down; stage index by distance from the top level
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = (nlev - 1 - L) % D;
#pragma unroll
      for (int s = 4; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if (s == 4 || ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj < g.nvy))) {
          const double* src = LB + (sg.row_off + (size_t)(sg.lvl[L] + q)) * 80 + (s - 4) * 16 + r * 4;
          double* dst = stageM + ((size_t)(st * 5 + (s - 4)) * nt + tid) * 4;
          cp_async16(dst, src);
          cp_async16(dst + 2, src + 2);
        }
      }
      cp_async8(stageR + st * nt + tid, v_out + 4 * (size_t)w.v + r);  // own forward result
    }
    cp_async_commit();
  };
  for (int k = 0; k < D; ++k) issue_bwd(nlev - 1 - k);
  for (int L = nlev - 1; L >= 0; --L) {
    cp_async_wait<D - 1>();
    const LevelRow w = level_row(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = (nlev - 1 - L) % D;
      double acc = stageR[st * nt + tid];
#pragma unroll
      for (int s = 5; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj < g.nvy)) {
          const double* up = stageM + ((size_t)(st * 5 + (s - 4)) * nt + tid) * 4;
          const double2 u01 = *reinterpret_cast<const double2*>(up);
          const double2 u23 = *reinterpret_cast<const double2*>(up + 2);
          const int Ln = L + di + 2 * dj;
          const int qn = (w.j + dj) - lev_jlo(Ln, wc);
          const double* xp = ring + ((size_t)(Ln & 3) * nq + qn) * 4;
          const double2 x01 = *reinterpret_cast<const double2*>(xp);
          const double2 x23 = *reinterpret_cast<const double2*>(xp + 2);
          double t = u01.x * x01.x;
          t += u01.y * x01.y;
          t += u23.x * x23.x;
          t += u23.y * x23.y;
          acc -= t;
        }
      }
      // x_r = sum_c Dinv[r][c] * acc_c : exchange acc within the quad through the ring slot
      const double* dp = stageM + ((size_t)(st * 5 + 0) * nt + tid) * 4;
      const double2 d01 = *reinterpret_cast<const double2*>(dp);
      const double2 d23 = *reinterpret_cast<const double2*>(dp + 2);
      const unsigned qmask = 0xFu << (tid & 28);
      const int qb = tid & 28;
      double xo = d01.x * __shfl_sync(qmask, acc, qb + 0);
      xo += d01.y * __shfl_sync(qmask, acc, qb + 1);
      xo += d23.x * __shfl_sync(qmask, acc, qb + 2);
      xo += d23.y * __shfl_sync(qmask, acc, qb + 3);
      ring[((size_t)(L & 3) * nq + q) * 4 + r] = xo;
      v_out[4 * (size_t)w.v + r] = xo;
      if (P2P) halo_push_value(hp, g.nvx, yp.nvy_glob, i0 + w.il, w.j + yp.j0, r, xo);
    }
    cta_sync();
    issue_bwd(L - D);
  }
  cp_async_wait<0>();
  if (P2P) halo_push_finish(hp, g.nvx, i0, wc);
}

// ------------------------------------------------------------------------------------------
// The same sweeps as a warp-specialised bulk-TMA pipeline. The packed layout stores the rows of a wavefront
// level contiguously, so ONE cp.async.bulk brings a whole level of the factor (rows x 512 B forward, rows x 640 B
// backward) into a shared-memory stage. A dedicated producer warp issues it, gathers the level's right-hand side
// with cp.async, and lets both complete on the stage's `full` mbarrier; the consumer warps only wait on that
// barrier, multiply, and hand the stage back through an `empty` mbarrier. What bounds the sweeps on small grids
// (paper / myelin configurations, multigrid coarse levels) is the instruction count on the per-level critical
// path of the consumers; besides moving all load issue off that path:
//  * quad q owns the vertex-column pair (2q, 2q+1) of the sub-slab for the whole sweep (column 2q + (L & 1) at
//    level L), so the ring of recent solution values is indexed by column and no neighbour needs lev_jlo;
//  * blocks of non-existent neighbours are stored as zeros by the factorisation and the ring is zero-initialised
//    and padded by one quad on either side, so the four block-row x vector products run without predicates;
//  * the level barrier is a __syncwarp when the sub-slab is swept by a single consumer warp.
// Pipeline step k: forward level L = k, backward level L = 2 nlev - 1 - k; stage k % D, use number k / D.
template <int D, bool P2P>
__global__ void __launch_bounds__(160) ilu_apply_tma_kernel(Dev g, const double* __restrict__ LF,
                                                            const double* __restrict__ LB, const double* __restrict__ d,
                                                            double* v_out, int slab_w, const int* __restrict__ lvl_tab,
                                                            HaloPush hp, int jc) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const YPart yp = ypart_select(g, jc, slab_w);
  {
    const size_t roff = (size_t)yp.j0 * g.nvx;
    LF += roff * 64; LB += roff * 80; d += 4 * roff; v_out += 4 * roff; lvl_tab += yp.lvl_off;
  }
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;                                     // consumer threads; the last warp produces
  const int nq = nt >> 2;
  double* stage = reinterpret_cast<double*>(smem_raw);               // [D][nq * 80]
  double* stageR = stage + (size_t)D * nq * 80;                       // [D][nt]
  double* ring = stageR + (size_t)D * nt;                             // [4][nq + 2][4]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + 4 * (size_t)(nq + 2) * 4);   // [D]
  unsigned long long* empty = full + D;                                                                 // [D]
  int* lvl_sh = reinterpret_cast<int*>(empty + D);
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) { mbar_init(&full[k], 33); mbar_init(&empty[k], (unsigned)(blockDim.x - 32) >> 5); }
    mbar_init_fence();
  }
  for (int k = tid; k < 4 * (nq + 2) * 4; k += blockDim.x) ring[k] = 0.0;
  const SlabGeom sg = slab_geom(g, slab_w, lvl_tab, lvl_sh);          // ends with __syncthreads
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  const bool producer = tid >= nt;

  if (producer) {
    const int lane = tid - nt;
    auto produce = [&](int k) {
      const int st = k % D, use = k / D;
      if (use > 0) mbar_wait(&empty[st], (unsigned)(use - 1) & 1u);
      const bool fwd = k < nlev;
      const int L = fwd ? k : 2 * nlev - 1 - k;
      if (lane == 0) {                                                // the level's factor rows: one bulk copy
        const unsigned rows = (unsigned)(sg.lvl[L + 1] - sg.lvl[L]);
        const unsigned bytes = rows * (fwd ? 512u : 640u);
        mbar_arrive_expect_tx(&full[st], bytes);
        const double* gsrc = fwd ? LF + (sg.row_off + (size_t)sg.lvl[L]) * 64 : LB + (sg.row_off + (size_t)sg.lvl[L]) * 80;
        bulk_g2s(stage + (size_t)st * nq * 80, gsrc, bytes, &full[st]);
      }
      const double* src = fwd ? d : v_out;                            // backward: the forward result y
      for (int t = lane; t < nt; t += 32) {                           // the level's right-hand side values
        const int il = 2 * (t >> 2) + (L & 1), tj = L - il, j = tj >> 1;
        if (il < wc && tj >= 0 && j < g.nvy)
          cp_async8(stageR + st * nt + t, src + 4 * (size_t)(i0 + il + g.nvx * j) + (t & 3));
      }
      cp_async_mbar_arrive_noinc(&full[st]);
    };
    for (int k = 0; k < nlev; ++k) produce(k);
    __syncthreads();          // the consumers' forward results are in global memory
    for (int k = nlev; k < 2 * nlev; ++k) produce(k);
  } else {
    // Consumers. Everything that does not change from level to level is computed here once; inside the loops the
    // level index only advances a handful of running values (stage, phase, ring slots, row index, row pointers).
    const int q = tid >> 2, r = tid & 3;
    const int rstride = (nq + 2) * 4;
    const unsigned full0 = smem_u32(full), empty0 = smem_u32(empty);
    const bool lane0 = (tid & 31) == 0, one_warp = nt == 32;
    const bool okE = 2 * q < wc, okO = 2 * q + 1 < wc;                 // my column at even / odd levels exists
    const int capE = ((wc + 1) >> 1) - 1, capO = (wc >> 1) - 1;        // row index in the stage: min(m, cap) - q
    // Shared memory is addressed through 32-bit shared-window addresses (ld/st.shared), and what changes from level
    // to level is kept as running byte offsets -- stage, right-hand side, barrier, ring slots, the global output
    // element -- so the per-level address arithmetic is a handful of 32-bit adds instead of generic-pointer math.
    const unsigned stage_u = smem_u32(stage) + (unsigned)r * 32u;      // my scalar row inside a block row
    const unsigned rhs_u = smem_u32(stageR) + (unsigned)tid * 8u;
    const unsigned ringq_u = smem_u32(ring) + (unsigned)(q + 1) * 32u; // my quad's 4 values inside a ring slot
    const unsigned ringown_u = ringq_u + (unsigned)r * 8u;
    const unsigned stage_stride_b = (unsigned)nq * 640u, rhs_stride_b = (unsigned)nt * 8u;
    const unsigned rstride_b = (unsigned)rstride * 8u;
    const long long vstride = 4LL * g.nvx;
    const long long vbase = 4LL * (i0 + 2 * q) + r;                    // my even column, grid row 0
    // ring neighbour quads (relative to my own slot) per level parity: column il + di lives in quad (il + di) >> 1
    //   forward  s = 0..3: di = -1, 0, +1, -1      backward s = 5..8: di = +1, -1, 0, +1
    //   forward  {-1, 0, 0, -1} (even) {0, 0, +1, 0} (odd) quads;  backward {0, -1, 0, 0} (even) {+1, 0, 0, +1} (odd)
    int st = 0;
    unsigned ph = 0, st_stage = 0, st_rhs = 0, st_bar = 0;
    unsigned so0 = 0, so1 = 3 * rstride_b, so2 = 2 * rstride_b, so3 = rstride_b;  // ring slots of levels L, L-1, L-2, L-3
    auto advance = [&]() {
      st_stage += stage_stride_b; st_rhs += rhs_stride_b; st_bar += 8u;
      if (++st == D) { st = 0; st_stage = 0; st_rhs = 0; st_bar = 0; ph ^= 1u; }
    };

    // ---------------- forward: y = L^-1 d ----------------
    long long voff = vbase - vstride * q;                              // element of (column 2q + b, row m - q), L = 0
    for (int L = 0; L < nlev; ++L) {
      const int b = L & 1, m = L >> 1, j = m - q;
      mbar_wait_a(full0 + st_bar, ph);
      const bool act = (b ? okO : okE) && (unsigned)j < (unsigned)g.nvy;
      double2 lv[4][2], yv[4][2];
      double acc = 0.0;
      if (act) {
        const unsigned la = stage_u + st_stage + (unsigned)(min(m, b ? capO : capE) - q) * 512u;
        const unsigned y0 = ringq_u + so3 - (b ? 0u : 32u);   // (-1,-1): level L-3
        const unsigned y1 = ringq_u + so2;                     // ( 0,-1): level L-2
        const unsigned y2 = ringq_u + so1 + (b ? 32u : 0u);   // (+1,-1): level L-1
        const unsigned y3 = ringq_u + so1 - (b ? 0u : 32u);   // (-1, 0): level L-1
#pragma unroll
        for (int s = 0; s < 4; ++s) { lv[s][0] = lds_d2(la + s * 128u); lv[s][1] = lds_d2(la + s * 128u + 16u); }
        yv[0][0] = lds_d2(y0); yv[0][1] = lds_d2(y0 + 16u);
        yv[1][0] = lds_d2(y1); yv[1][1] = lds_d2(y1 + 16u);
        yv[2][0] = lds_d2(y2); yv[2][1] = lds_d2(y2 + 16u);
        yv[3][0] = lds_d2(y3); yv[3][1] = lds_d2(y3 + 16u);
        acc = lds_d(rhs_u + st_rhs);
      }
      // hand the stage back as soon as the operands are in registers (mbarrier.arrive releases the warp's reads)
      __syncwarp();
      if (lane0) mbar_arrive_a(empty0 + st_bar);
      if (act) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          double t = lv[s][0].x * yv[s][0].x;
          t += lv[s][0].y * yv[s][0].y;
          t += lv[s][1].x * yv[s][1].x;
          t += lv[s][1].y * yv[s][1].y;
          acc -= t;
        }
        sts_d(ringown_u + so0, acc);
        v_out[voff] = acc;
      }
      if (one_warp) __syncwarp();
      else named_bar_sync(1, nt);
      advance();
      { const unsigned t = so3; so3 = so2; so2 = so1; so1 = so0; so0 = t; }
      voff += b ? vstride - 4 : 4;
    }
    __syncthreads();
    // ---------------- backward: x = U^-1 y ----------------
    // slot offsets now refer to levels L, L+1, L+2, L+3
    unsigned su0, su1, su2, su3;
    {
      const int top = nlev - 1;
      su0 = (unsigned)(top & 3) * rstride_b; su1 = (unsigned)((top + 1) & 3) * rstride_b;
      su2 = (unsigned)((top + 2) & 3) * rstride_b; su3 = (unsigned)((top + 3) & 3) * rstride_b;
      voff = vbase + vstride * ((top >> 1) - q) + 4 * (top & 1);
    }
    for (int L = nlev - 1; L >= 0; --L) {
      const int b = L & 1, m = L >> 1, j = m - q;
      mbar_wait_a(full0 + st_bar, ph);
      const bool act = (b ? okO : okE) && (unsigned)j < (unsigned)g.nvy;
      double2 uv[5][2], xv[4][2];
      double acc = 0.0;
      if (act) {
        const unsigned ua = stage_u + st_stage + (unsigned)(min(m, b ? capO : capE) - q) * 640u;
        const unsigned x0 = ringq_u + su1 + (b ? 32u : 0u);   // (+1, 0): level L+1
        const unsigned x1 = ringq_u + su1 - (b ? 0u : 32u);   // (-1,+1): level L+1
        const unsigned x2 = ringq_u + su2;                     // ( 0,+1): level L+2
        const unsigned x3 = ringq_u + su3 + (b ? 32u : 0u);   // (+1,+1): level L+3
#pragma unroll
        for (int s = 0; s < 5; ++s) { uv[s][0] = lds_d2(ua + s * 128u); uv[s][1] = lds_d2(ua + s * 128u + 16u); }
        xv[0][0] = lds_d2(x0); xv[0][1] = lds_d2(x0 + 16u);
        xv[1][0] = lds_d2(x1); xv[1][1] = lds_d2(x1 + 16u);
        xv[2][0] = lds_d2(x2); xv[2][1] = lds_d2(x2 + 16u);
        xv[3][0] = lds_d2(x3); xv[3][1] = lds_d2(x3 + 16u);
        acc = lds_d(rhs_u + st_rhs);
      }
      __syncwarp();
      if (lane0) mbar_arrive_a(empty0 + st_bar);
      if (act) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          double t = uv[s + 1][0].x * xv[s][0].x;
          t += uv[s + 1][0].y * xv[s][0].y;
          t += uv[s + 1][1].x * xv[s][1].x;
          t += uv[s + 1][1].y * xv[s][1].y;
          acc -= t;
        }
        // x_r = sum_c Dinv[r][c] * acc_c : exchange acc within the quad
        const unsigned qmask = 0xFu << (tid & 28);
        const int qb = tid & 28;
        double xo = uv[0][0].x * __shfl_sync(qmask, acc, qb + 0);
        xo += uv[0][0].y * __shfl_sync(qmask, acc, qb + 1);
        xo += uv[0][1].x * __shfl_sync(qmask, acc, qb + 2);
        xo += uv[0][1].y * __shfl_sync(qmask, acc, qb + 3);
        sts_d(ringown_u + su0, xo);
        v_out[voff] = xo;
        if (P2P) halo_push_value(hp, g.nvx, yp.nvy_glob, i0 + 2 * q + b, j + yp.j0, r, xo);
      }
      if (one_warp) __syncwarp();
      else named_bar_sync(1, nt);
      advance();
      { const unsigned t = su3; su3 = su2; su2 = su1; su1 = su0; su0 = t; }
      voff -= b ? 4 : vstride - 4;
    }
  }
  if (P2P) halo_push_finish(hp, g.nvx, i0, wc);
}

// ------------------------------------------------------------------------------------------
__global__ void rownorm_kernel(Dev g, double* A, double* rv) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(t >> 2), r = (int)(t & 3);
  if (v >= g.nv) return;
  double* Arow = A + (size_t)v * AX1_ROW;
  const int off = r == 3 ? 6 : 2 * r, cnt = r == 3 ? 4 : 2;
  double n = 0.0;
  for (int c = 0; c < cnt; ++c) n = fmax(n, fabs(Arow[4 * AX1_BLK + off + c]));
  if (n) {
    for (int s = 0; s < 9; ++s)
      for (int c = 0; c < cnt; ++c) Arow[s * AX1_BLK + off + c] /= n;
    rv[4 * (size_t)v + r] /= n;
  } else {
    Arow[4 * AX1_BLK + (r == 3 ? 9 : 2 * r)] = 1.0;
  }
}

// 9-slot layout <-> compact BCRS (download / upload of the Jacobian in the reference's layout)
// arrow-block position of dense entry (rr, cc), -1 if structurally zero
__device__ __forceinline__ int arrow_pos(int rr, int cc) {
  if (rr == 3) return 6 + cc;
  if (cc == rr) return 2 * rr;
  if (cc == 3) return 2 * rr + 1;
  return -1;
}
// device layout <-> compact BCRS with dense 4x4 blocks (the reference's BCRSMatrix layout). On upload
// (to_bcrs = 0) a non-zero outside the arrow shape sets *viol (such a matrix is not a PNP Jacobian).
__global__ void bcrs_copy_kernel(Dev g, double* A9, double* blocks, const int* __restrict__ rowptr, int to_bcrs, int* viol) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long vs = t >> 4;
  const int e = (int)(t & 15);
  if (vs >= 9LL * g.nv) return;
  const int v = (int)(vs / 9), s = (int)(vs % 9);
  const int i = v % g.nvx, j = v / g.nvx;
  int pos = 0;
  bool mine = false;
  for (int k = 0; k < 9; ++k) {
    const int ii = i + k % 3 - 1, jj = j + k / 3 - 1;
    const bool inb = ii >= 0 && ii < g.nvx && jj >= 0 && jj < g.nvy;
    if (k == s) { mine = inb; break; }
    pos += inb;
  }
  const int ap = arrow_pos(e >> 2, e & 3);
  if (!mine) {
    if (!to_bcrs && ap >= 0) A9[vs * AX1_BLK + ap] = 0.0;
    return;
  }
  const size_t b = ((size_t)rowptr[v] + pos) * 16 + e;
  if (to_bcrs) blocks[b] = ap >= 0 ? A9[vs * AX1_BLK + ap] : 0.0;
  else if (ap >= 0) A9[vs * AX1_BLK + ap] = blocks[b];
  else if (blocks[b] != 0.0) atomicExch(viol, 1);
}

// The solver's SpMV on the 64-byte-block copy of A (compact_rows_kernel below; 40 products per Newton iteration):
// a block is [d0 c0 | d1 c1 | d2 c2 | q p3], the potential row's three concentration entries folded into q (they are
// q times the valences: the charge-density term, acme2_cyl_operator_fully_coupled.hh:535) -- 576 B instead of 720 B
// per block row. The copy is tiled for the load pattern of this kernel: [8 vertices][9 slots][vertex in group][8],
// so the 16 B loads of a warp (8 vertices x 4 lanes, one slot) cover 512 contiguous bytes = 4 full cache lines
// instead of 8 half-used ones, and each lane loads ONE component of the neighbour's x (a coalesced 256 B per warp);
// x3 and the valence-weighted sum come from the quad by shuffles. L1 wavefronts per row: 54 instead of 108 --
// they, not DRAM, held the row-layout variant of this kernel at 83 % of the HBM peak.
template <int DOT>
__global__ void __launch_bounds__(256, 8) spmv_compact_kernel(Dev g, const double* __restrict__ Ac, const double* __restrict__ x,
                                                              double* __restrict__ y, int zero_con,
                                                              const double* __restrict__ w, Own own, const Scalars* sc,
                                                              double* partials) {
  double s[2] = {0.0, 0.0};
  if (!(DOT && sc && sc->done)) {                  // uniform: all lanes of a warp take part in the shuffles
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int r = (int)(t & 3);
    const double zr = r < 3 ? (double)g.valence[r] : 0.0;
    const int vt = (int)(t >> 2);
    const int v = vt < g.nv ? vt : g.nv - 1;       // lanes past the end compute the last row and do not store
    const int i = v % g.nvx, j = v / g.nvx;
    const double* Arow = Ac + ((size_t)(v >> 3) * 9) * (8 * AX1_CBLK) + (v & 7) * AX1_CBLK + 2 * r;
    double acc = 0.0;
#pragma unroll
    for (int sl = 0; sl < 9; ++sl) {
      const int di = sl % 3 - 1, dj = sl / 3 - 1;
      const int ii = i + di, jj = j + dj;
      const bool ok = ii >= 0 && ii < g.nvx && jj >= 0 && jj < g.nvy;      // missing neighbours hold zero blocks
      const double xr = ok ? __ldg(x + 4 * (size_t)(v + di + dj * g.nvx) + r) : 0.0;
      const double2 a = __ldg(reinterpret_cast<const double2*>(Arow + sl * (8 * AX1_CBLK)));
      const double x3 = __shfl_sync(0xffffffffu, xr, 3, 4);
      double zx = zr * xr;
      zx += __shfl_xor_sync(0xffffffffu, zx, 1, 4);
      zx += __shfl_xor_sync(0xffffffffu, zx, 2, 4);
      acc += a.x * (r == 3 ? zx : xr);
      acc += a.y * x3;
    }
    if (vt < g.nv) {
      if (zero_con && ((g.dmask[v] >> r) & 1)) acc = 0.0;
      y[4 * (size_t)v + r] = acc;
      if (DOT == 1 || DOT == 2) {
        if (!own.masked || (i >= own.lo && i <= own.hi)) {
          const double wv = __ldg(w + 4 * (size_t)v + r);
          s[0] = acc * wv;
          if (DOT == 2) s[1] = acc * acc;
        }
      }
    }
  }
  if (DOT == 1 || DOT == 2) block_reduce_store<2>(s, partials);
}
// ------------------------------------------------------------------------------------------
int spmv_blocks(const Dev& g) { return (int)((4LL * g.nv + 255) / 256); }
void launch_spmv(const Dev& g, const double* A, const double* x, double* y, int zero_con, cudaStream_t st,
                 const double* Ac) {
  if (Ac) spmv_compact_kernel<0><<<spmv_blocks(g), 256, 0, st>>>(g, Ac, x, y, zero_con, nullptr, Own{0, 0, g.nvx, 0}, nullptr, nullptr);
  else spmv_kernel<0><<<spmv_blocks(g), 256, 0, st>>>(g, A, x, y, zero_con, nullptr, Own{0, 0, g.nvx, 0}, nullptr, nullptr);
}
// r = b - A x
void launch_spmv_residual(const Dev& g, const double* A, const double* x, const double* b, double* r, int zero_con,
                          cudaStream_t st) {
  spmv_kernel<3><<<spmv_blocks(g), 256, 0, st>>>(g, A, x, r, zero_con, b, Own{0, 0, g.nvx, 0}, nullptr, nullptr);
}
void launch_spmv_dot(const Dev& g, const double* A, const double* x, double* y, int zero_con, int dot_mode,
                     const double* w, int own_lo, int own_hi, int masked, const Scalars* sc, double* partials,
                     cudaStream_t st, const double* Ac) {
  const Own own{own_lo, own_hi, g.nvx, masked};
  if (Ac) {
    if (dot_mode == 1) spmv_compact_kernel<1><<<spmv_blocks(g), 256, 0, st>>>(g, Ac, x, y, zero_con, w, own, sc, partials);
    else spmv_compact_kernel<2><<<spmv_blocks(g), 256, 0, st>>>(g, Ac, x, y, zero_con, w, own, sc, partials);
    return;
  }
  if (dot_mode == 1) spmv_kernel<1><<<spmv_blocks(g), 256, 0, st>>>(g, A, x, y, zero_con, w, own, sc, partials);
  else spmv_kernel<2><<<spmv_blocks(g), 256, 0, st>>>(g, A, x, y, zero_con, w, own, sc, partials);
}

// Ac = the 64-byte-block copy of the arrow rows of A for the solver's products, tiled by groups of 8 vertices (see
// spmv_compact_kernel; the caller allocates nv rounded up to a multiple of 8 rows). One thread per (vertex, slot). *flag is raised if a block's potential row is not EXACTLY q times the valences (finite-difference
// volume Jacobians, uploaded matrices): the solver then multiplies with A itself.
__global__ void __launch_bounds__(256) compact_rows_kernel(Dev g, const double* __restrict__ A, double* __restrict__ Ac,
                                                           int* flag) {
  // the CTA's 256 blocks (80 B each) are staged through shared memory with coalesced 16 B loads: a thread reading its
  // own block straight from global memory touches 20 cache lines per warp instruction and the kernel runs at the L1's
  // pace (78 % L1 throughput under ncu, 4.8 ms); an 80 B stride is conflict-free for LDS.128 (8 lanes per wavefront)
  __shared__ double2 sh[256 * 5];
  const long long nb = 9LL * g.nv;
  const long long b0 = (long long)blockIdx.x * 256;
  const double2* A2 = reinterpret_cast<const double2*>(A) + b0 * 5;
  const long long left = (nb - b0) * 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int e = k * 256 + threadIdx.x;
    if (e < left) sh[e] = __ldg(A2 + e);
  }
  __syncthreads();
  const long long t = b0 + threadIdx.x;
  if (t >= nb) return;
  const double2* p = sh + threadIdx.x * 5;
  const double2 dc0 = p[0], dc1 = p[1], dc2 = p[2], p01 = p[3], p23 = p[4];
  const double z0 = g.valence[0], z1 = g.valence[1], z2 = g.valence[2];
  const double q = p01.x / z0;
  if (z0 * q != p01.x || z1 * q != p01.y || z2 * q != p23.x) atomicExch(flag, 1);
  const long long vv = t / 9;
  const int sl = (int)(t - 9 * vv);
  double2* o = reinterpret_cast<double2*>(Ac + ((size_t)(vv >> 3) * 9 + sl) * (8 * AX1_CBLK) + (vv & 7) * AX1_CBLK);
  o[0] = dc0;
  o[1] = dc1;
  o[2] = dc2;
  o[3] = make_double2(q, p23.y);
}
void launch_compact_rows(const Dev& g, const double* A, double* Ac, int* flag, cudaStream_t st) {
  const long long nt = 9LL * g.nv;
  compact_rows_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(g, A, Ac, flag);
}
// one quad per wavefront row: a level of a slab of width w has at most ceil(w/2) rows
static int slab_threads(int slab_w) {
  int rows = (slab_w + 1) / 2;
  int t = 4 * rows;
  t = ((t + 31) / 32) * 32;
  if (t > 256) t = 256;
  if (t < 32) t = 32;
  return t;
}
// level prefix tables for full-width sub-slabs and for the (narrower) last one: 2 x (slab_w + 2 Ny + 1) ints;
// with a radial split (jc > 0) one such pair per part, the lower part (rows [0, jc)) first
static void level_tables_part(int nvx, int nvy, int slab_w, std::vector<int>& tab) {
  const int stride = slab_w + 2 * (nvy - 1) + 1;
  const size_t base = tab.size();
  tab.resize(base + 2 * (size_t)stride, 0);
  const int last_wc = nvx - ((nvx - 1) / slab_w) * slab_w;
  const int wcs[2] = {slab_w, last_wc};
  for (int t = 0; t < 2; ++t) {
    const int wc = wcs[t], nlev = wc + 2 * (nvy - 1);
    int acc = 0;
    for (int L = 0; L < nlev; ++L) {
      tab[base + (size_t)t * stride + L] = acc;
      const int a = L - wc + 1;
      const int jlo = a <= 0 ? 0 : (a + 1) / 2;
      const int jhi = std::min(nvy - 1, L >> 1);
      if (jhi >= jlo) acc += jhi - jlo + 1;
    }
    tab[base + (size_t)t * stride + nlev] = acc;
  }
}
void ilu_level_tables(const Dev& g, int slab_w, std::vector<int>& tab, int jc) {
  tab.clear();
  if (jc > 0) {
    level_tables_part(g.nvx, jc, slab_w, tab);
    level_tables_part(g.nvx, g.nvy - jc, slab_w, tab);
  } else {
    level_tables_part(g.nvx, g.nvy, slab_w, tab);
  }
}
// AX1_ILU_FACTOR16=0 keeps the quad-per-row factorisation for narrow sub-slabs too (A/B)
static bool ilu_use_factor16() {
  static const int v = getenv("AX1_ILU_FACTOR16") ? atoi(getenv("AX1_ILU_FACTOR16")) : 1;
  return v != 0;
}
void launch_ilu_factor(const Dev& g, const double* A, double* LU, int slab_w, const int* lvl_tab, cudaStream_t st,
                       int jc) {
  const dim3 nslab((g.nvx + slab_w - 1) / slab_w, jc > 0 ? 2 : 1);
  // measured (same box): 16 columns 0.74 -> 0.66 ms on the paper grid, 24 columns 0.81 -> 0.79 ms; at 32 columns
  // (256 threads) the quad-per-row kernel is faster again (1.17 vs 1.28 ms on the myelin grid)
  if (slab_w <= 24 && ilu_use_factor16()) {
    const int nq = (((slab_w + 1) / 2) + 1) & ~1;  // rows per level, even: whole warps
    const size_t smem = (size_t)nq * (3 * AX1_ROW + 3 * AX1_RING16_ROW + 32) * sizeof(double) + (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
    static bool attr16 = false;
    if (!attr16) {
      cudaFuncSetAttribute(ilu_factor16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr16 = true;
    }
    ilu_factor16_kernel<<<nslab, 16 * nq, smem, st>>>(g, A, LU, LU + (size_t)g.nv * 64, slab_w, nq, lvl_tab, jc);
    return;
  }
  const int nt = slab_threads(slab_w), nq = nt / 4;
  const size_t smem = (size_t)nq * (2 * AX1_ROW + 3 * AX1_RING_ROW) * sizeof(double) + (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ilu_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  ilu_factor_kernel<<<nslab, nt, smem, st>>>(g, A, LU, LU + (size_t)g.nv * 64, slab_w, lvl_tab, jc);
}
// Depth of the cp.async pipeline of the sweeps. A level is consumed D levels after its loads were issued, so
// the sweep advances one level per (HBM latency / D) unless bandwidth limits it first. With 256-thread CTAs
// (sub-slabs of 128 columns) a stage is 43 KB and two stages x two CTAs per SM already saturate HBM; narrow
// sub-slabs / small grids (paper and myelin configurations) are latency bound and get a deeper pipeline.
int ilu_pick_stages(int nt, int bytes_per_thread_stage) {
  static const int env = getenv("AX1_ILU_STAGES") ? atoi(getenv("AX1_ILU_STAGES")) : 0;
  if (env > 0) return env;
  const int d = (110 * 1024) / (nt * bytes_per_thread_stage);
  return d >= 8 ? 8 : (d >= 6 ? 6 : (d >= 4 ? 4 : (d >= 3 ? 3 : 2)));
}
// AX1_ILU_TMA=0 selects the cp.async (LDGSTS) sweeps instead of the bulk-TMA ones
static bool ilu_use_tma() {
  static const int v = getenv("AX1_ILU_TMA") ? atoi(getenv("AX1_ILU_TMA")) : 1;
  return v != 0;
}
template <int D, bool P2P>
static void launch_ilu_apply_t(const Dev& g, const double* LU, const double* d, double* v, int slab_w,
                               const int* lvl_tab, const HaloPush& hp, cudaStream_t st, int jc) {
  const dim3 nslab((g.nvx + slab_w - 1) / slab_w, jc > 0 ? 2 : 1);
  const int nt = slab_threads(slab_w);
  // Sub-slabs of up to 64 columns (small grids: the sweep is bound by per-level latency) take the warp-specialised
  // TMA pipeline; at 128 columns the sweep is bandwidth-bound, both kernels reach the same ~0.94 of the copy peak
  // (3.7-3.9 ms at 16.9 M vertices) and the cp.async kernel needs fewer registers for its 2 CTAs per SM.
  if (ilu_use_tma() && nt <= 128) {
    const int nq = nt / 4;
    const size_t smem_t = (size_t)D * nq * 640 + (size_t)D * nt * 8 + 4 * (size_t)(nq + 2) * 32 + (size_t)D * 16 +
                          (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
    static bool attr_t = false;
    if (!attr_t) {
      cudaFuncSetAttribute(ilu_apply_tma_kernel<D, P2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr_t = true;
    }
    ilu_apply_tma_kernel<D, P2P><<<nslab, nt + 32, smem_t, st>>>(g, LU, LU + (size_t)g.nv * 64, d, v, slab_w, lvl_tab, hp, jc);
    return;
  }
  const size_t smem = (size_t)D * nt * (5 * 32 + 8) + 4 * (size_t)(nt / 4) * 32 + (size_t)(slab_w + 2 * g.nvy) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ilu_apply_kernel<D, P2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  ilu_apply_kernel<D, P2P><<<nslab, nt, smem, st>>>(g, LU, LU + (size_t)g.nv * 64, d, v, slab_w, lvl_tab, hp, jc);
}
template <bool P2P>
static void launch_ilu_apply_d(const Dev& g, const double* LU, const double* d, double* v, int slab_w,
                               const int* lvl_tab, const HaloPush& hp, cudaStream_t st, int jc) {
  switch (ilu_pick_stages(slab_threads(slab_w), 5 * 32 + 8)) {
    case 8: launch_ilu_apply_t<8, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st, jc); break;
    case 6: launch_ilu_apply_t<6, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st, jc); break;
    case 4: launch_ilu_apply_t<4, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st, jc); break;
    case 3: launch_ilu_apply_t<3, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st, jc); break;
    default: launch_ilu_apply_t<2, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st, jc); break;
  }
}
void launch_ilu_apply(const Dev& g, const double* LU, const double* d, double* v, int slab_w, const int* lvl_tab,
                      cudaStream_t st, const HaloPush* hp, int jc) {
  // the inputs of the preconditioner are already zero on constrained DOFs (see driver.cu)
  if (hp) launch_ilu_apply_d<true>(g, LU, d, v, slab_w, lvl_tab, *hp, st, jc);
  else launch_ilu_apply_d<false>(g, LU, d, v, slab_w, lvl_tab, HaloPush{}, st, jc);
}
// sub-slab CTAs that own part of the left / right three halo vertex columns
void halo_contributors(const Dev& g, int slab_w, int* left, int* right, int jc) {
  const int nslab = (g.nvx + slab_w - 1) / slab_w;
  int l = 0, r = 0;
  for (int b = 0; b < nslab; ++b) {
    const int i0 = b * slab_w, wc = std::min(slab_w, g.nvx - i0);
    if (i0 < 3) ++l;
    if (i0 + wc > g.nvx - 3) ++r;
  }
  if (jc > 0) { l *= 2; r *= 2; }   // radial split: two CTAs per sub-slab
  *left = l; *right = r;
}
void launch_rownorm(const Dev& g, double* A, double* r, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 4LL * g.nv;
  rownorm_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, A, r);
}
void launch_compact_bcrs(const Dev& g, const double* A9, double* blocks, const int* rowptr, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 9LL * 16 * g.nv;
  bcrs_copy_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, const_cast<double*>(A9), blocks, rowptr, 1, nullptr);
}
void launch_expand_bcrs(const Dev& g, const double* blocks, double* A9, const int* rowptr, int* viol, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 9LL * 16 * g.nv;
  bcrs_copy_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, A9, const_cast<double*>(blocks), rowptr, 0, viol);
}

}  // namespace ax1
