// ilu1.cu -- block ILU(1) factor / apply: the reference's default preconditioner
//   ISTLBackend_OVLP_BCGS_ILUn(multigfs, cc, n = 1, ...) / ISTLBackend_SEQ_BCGS_ILUn(1, 1.0, ...)
//   (dune/ax1/acme2_cyl/common/acme2_cyl_setup.hh:764-765, 797-798) = Dune::SeqILUn ->
//   bilu_decomposition(A, 1, ILU) + bilu_backsolve (dune-istl 2.3 ilu.hh)                          [ext]
//
// Symbolic phase, done by hand for the x-fastest 9-point vertex stencil: a level-1 fill entry (r,c)
// appears where a level-0 lower entry (r,k) meets a level-0 upper entry (k,c). With r = (i,j):
//   k = (i-1,j-1) -> (i-2,j);  k = (i+1,j-1) -> (i+2,j-1), (i+2,j);  k = (i-1,j) -> (i-2,j+1)
// so a row has 13 blocks (6 lower, diagonal, 6 upper), in ascending column order
//   lower (i-1,j-1) (i,j-1) (i+1,j-1) (i+2,j-1) (i-2,j) (i-1,j) | (i,j) | upper (i+1,j) (i+2,j) (i-2,j+1) (i-1,j+1) (i,j+1) (i+1,j+1)
// ("13 blocks/row"). Rows whose generating neighbour does not exist (j = 0) keep
// an exactly-zero block in the fill slot, which is numerically the same as not having the entry.
// The numeric phase is bilu0_decomposition on that pattern.
//
// Parallelisation as for ILU0 (solver.cu): one CTA per x-sub-slab; row (i,j) depends on
// (i+2,j-1) and (i-1,j), so the wavefronts are level = il + 3 j (<= ceil(w/3) rows each); the factor
// is stored wavefront-packed: LF 6 lower blocks (768 B), LB inverted diagonal + 6 upper blocks (896 B).
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "ilu_common.cuh"
#include "p2p.cuh"

namespace ax1 {

#define P1_DI(p) ((p) < 4 ? (p)-1 : ((p) < 9 ? (p)-6 : (p)-11))
#define P1_DJ(p) ((p) < 4 ? -1 : ((p) < 9 ? 0 : 1))
__host__ __device__ constexpr int p1_idx(int di, int dj) {
  return dj == -1 ? ((di >= -1 && di <= 2) ? di + 1 : -1)
                  : dj == 0 ? ((di >= -2 && di <= 2) ? di + 6 : -1)
                            : dj == 1 ? ((di >= -2 && di <= 1) ? di + 11 : -1) : -1;
}

__device__ __forceinline__ int lev1_jlo(int L, int wc) {
  const int a = L - wc + 1;
  return a <= 0 ? 0 : (a + 2) / 3;
}

struct Slab1 {
  int i0, wc, nlev;
  size_t row_off;
  const int* lvl;
};
__device__ __forceinline__ Slab1 slab1_geom(const Dev& g, int slab_w, const int* __restrict__ lvl_tab, int* lvl_sh) {
  Slab1 sg;
  sg.i0 = blockIdx.x * slab_w;
  sg.wc = min(slab_w, g.nvx - sg.i0);
  sg.nlev = sg.wc + 3 * (g.nvy - 1);
  sg.row_off = (size_t)blockIdx.x * slab_w * g.nvy;
  const int stride = slab_w + 3 * (g.nvy - 1) + 1;
  const int* src = lvl_tab + (sg.wc == slab_w ? 0 : stride);
  for (int k = threadIdx.x; k <= sg.nlev; k += blockDim.x) lvl_sh[k] = src[k];
  __syncthreads();
  sg.lvl = lvl_sh;
  return sg;
}

struct Row1 { int v, il, j; };
__device__ __forceinline__ Row1 level_row1(const Dev& g, int L, int nlev, int i0, int wc, int q) {
  Row1 o;
  o.v = -1; o.il = 0; o.j = 0;
  if (L < 0 || L >= nlev) return o;
  const int j = lev1_jlo(L, wc) + q;
  if (j > min(g.nvy - 1, L / 3)) return o;
  o.j = j; o.il = L - 3 * j;
  o.v = i0 + o.il + g.nvx * j;
  return o;
}

// ------------------------------------------------------------------------------------------
// factorisation of one sub-slab per CTA, one quad per wavefront row (lane = scalar row of the blocks).
// The Jacobian row of level L+2 is staged with cp.async; the finished [Dinv | U x 6] rows of the last
// four levels (all a row needs from its lower neighbours) sit in a shared-memory ring.
#define AX1_RING1_ROW 114  // 7 blocks = 112 doubles + 2 padding
__global__ void __launch_bounds__(256, 1) ilu1_factor_kernel(Dev g, const double* __restrict__ A, double* LF, double* LB,
                                                             int slab_w, int nq, const int* __restrict__ lvl_tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, q = tid >> 2, r = tid & 3;  // nq = rows a level can hold = ceil(slab_w / 3)
  double* stage = reinterpret_cast<double*>(smem_raw);   // [2][nq][AX1_ROW]
  double* ring = stage + 2 * (size_t)nq * AX1_ROW;        // [4][nq][AX1_RING1_ROW]
  const Slab1 sg = slab1_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 4 * (size_t)nq * AX1_RING1_ROW));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  for (int k = tid; k < 4 * nq * AX1_RING1_ROW; k += blockDim.x) ring[k] = 0.0;  // finite don't-care rows (see below)
  auto issue = [&](int L) {
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
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
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    const bool act = w.v >= 0;
    double row[13][4];
    row[6][0] = row[6][1] = row[6][2] = row[6][3] = 0.0;
    if (act) {
      const int il = w.il, j = w.j;
      bool ex[13];
      const double* arow = stage + ((size_t)(L & 1) * nq + q) * AX1_ROW;
#pragma unroll
      for (int p = 0; p < 13; ++p) {
        const int di = P1_DI(p), dj = P1_DJ(p);
        ex[p] = (il + di >= 0) && (il + di < wc) && (j + dj >= 0) && (j + dj < g.nvy);
        if (di >= -1 && di <= 1) {
          // level-0 entry: expand the arrow block row of the Jacobian (common.cuh), branch-free in the lane index
          const double* blk = arow + ((dj + 1) * 3 + di + 1) * AX1_BLK;
          const double2 a = *reinterpret_cast<const double2*>(blk + (r == 3 ? 6 : 2 * r));
          const double2 b2 = *reinterpret_cast<const double2*>(blk + (r == 3 ? 8 : 2 * r));
          row[p][0] = (ex[p] && (r == 3 || r == 0)) ? a.x : 0.0;
          row[p][1] = ex[p] ? (r == 3 ? a.y : (r == 1 ? a.x : 0.0)) : 0.0;
          row[p][2] = ex[p] ? (r == 3 ? b2.x : (r == 2 ? a.x : 0.0)) : 0.0;
          row[p][3] = ex[p] ? (r == 3 ? b2.y : a.y) : 0.0;
        } else {
          row[p][0] = row[p][1] = row[p][2] = row[p][3] = 0.0;  // level-1 fill entry
        }
      }
      // branch-free elimination, as in ilu_factor_kernel (solver.cu): a missing lower neighbour has a zero block row,
      // so its L row and every update it feeds are zero; a U block that points at a missing vertex is zero in k's row too
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int dik = P1_DI(k), djk = P1_DJ(k);
        const int Ln = L + dik + 3 * djk;
        const int qn = (j + djk) - lev1_jlo(Ln, wc);
        const double* kb = ex[k] ? ring + ((size_t)(Ln & 3) * nq + qn) * AX1_RING1_ROW : ring;  // row k: [Dinv, U7..U12]
        double lrow[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double sum = 0.0;
#pragma unroll
          for (int m = 0; m < 4; ++m) sum += row[k][m] * kb[4 * m + c];
          lrow[c] = sum;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) row[k][c] = lrow[c];
#pragma unroll
        for (int t = 7; t < 13; ++t) {
          const int p2 = p1_idx(dik + P1_DI(t), djk + P1_DJ(t));
          if (p2 < 0) continue;
          const double* ub = kb + (t - 6) * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum += lrow[m] * ub[4 * m + c];
            row[p2][c] -= sum;
          }
        }
      }
    }
    // invert the pivot block (every lane of the quad holds one row of it); converged full-mask exchange
    double d[4][4], dinv[4][4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int c = 0; c < 4; ++c) d[m][c] = __shfl_sync(0xffffffffu, row[6][c], (tid & 28) + m);
    if (act) {
      invert4(d, dinv);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double val = dinv[0][c];
        if (r == 1) val = dinv[1][c];
        if (r == 2) val = dinv[2][c];
        if (r == 3) val = dinv[3][c];
        row[6][c] = val;
      }
      const size_t p = sg.row_off + (size_t)(sg.lvl[L] + q);
#pragma unroll
      for (int k = 0; k < 6; ++k) st4(LF + p * 96 + k * 16 + r * 4, row[k]);
#pragma unroll
      for (int k = 6; k < 13; ++k) st4(LB + p * 112 + (k - 6) * 16 + r * 4, row[k]);
    }
    cta_sync();  // every read of ring slot L&3 (it held level L-4) is done
    if (w.v >= 0) {
      double* kr = ring + ((size_t)(L & 3) * nq + q) * AX1_RING1_ROW;
#pragma unroll
      for (int k = 6; k < 13; ++k) st4(kr + (k - 6) * 16 + r * 4, row[k]);
    }
    issue(L + 2);
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// The same factorisation for narrow sub-slabs, 16 lanes per vertex row (lane (r, c) owns element (r, c) of each of
// the row's 13 blocks) with the levels software-pipelined, as ilu_factor16_kernel (solver.cu) does for ILU0: of the
// six lower neighbours, taken in ascending column order, the first three lie on levels L-4, L-3, L-2 and are
// eliminated one iteration early, while level L-1 still runs its last three steps and inverts its pivot block.
// Same elimination in the same order per row; the pivot blocks are inverted by the adjugate formula (ilu_common.cuh),
// equal to ilu1_factor_kernel's pivoted Gauss-Jordan to rounding.
#define AX1_RING1_16_ROW 116  // 7 blocks = 112 doubles + 4: the two row groups of a warp read disjoint banks
__device__ __forceinline__ double shfl1_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__global__ void __launch_bounds__(256) ilu1_factor16_kernel(Dev g, const double* __restrict__ A, double* LF, double* LB,
                                                            int slab_w, int nq, const int* __restrict__ lvl_tab) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, q = tid >> 4, e = tid & 15, r = e >> 2, c = e & 3;
  const int lane = tid & 31, rowl = lane & ~3, grp = lane & 16;
  double* stage = reinterpret_cast<double*>(smem_raw);   // [3][nq][AX1_ROW]
  double* ring = stage + 3 * (size_t)nq * AX1_ROW;        // [4][nq][AX1_RING1_16_ROW]
  const Slab1 sg = slab1_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 4 * (size_t)nq * AX1_RING1_16_ROW + 2 * (size_t)nq * 16));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;
  for (int k = tid; k < 4 * nq * AX1_RING1_16_ROW; k += blockDim.x) ring[k] = 0.0;  // finite don't-care rows
  double* pscratch = ring + 4 * (size_t)nq * AX1_RING1_16_ROW + (size_t)q * 16;  // [2][nq][16]: pivot block of my row, by level parity
  const Adj16 adj = adj16_setup(r, c);
  const bool nz = (r == 3) || (c == 3) || (r == c);
  const int aoff = (r == 3) ? 6 + c : (c == 3 ? 2 * r + 1 : 2 * r);

  auto issue = [&](int L) {
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const double* src = A + (size_t)w.v * AX1_ROW;
      double* dst = stage + ((size_t)(L % 3) * nq + q) * AX1_ROW;
      for (int k = e; k < AX1_ROW / 2; k += 16) cp_async16(dst + 2 * k, src + 2 * k);
    }
    cp_async_commit();
  };
  auto exists = [&](const Row1& w) {
    unsigned m = 0;
    if (w.v >= 0) {
#pragma unroll
      for (int p = 0; p < 13; ++p) {
        const int di = P1_DI(p), dj = P1_DJ(p);
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0) && (w.j + dj < g.nvy)) m |= 1u << p;
      }
    }
    return m;
  };
  auto expand = [&](double (&x)[13], int L, unsigned exm) {
    const double* arow = stage + ((size_t)(L % 3) * nq + q) * AX1_ROW;
#pragma unroll
    for (int p = 0; p < 13; ++p) {
      const int di = P1_DI(p), dj = P1_DJ(p);
      if (di >= -1 && di <= 1)   // level-0 entry: element (r, c) of the arrow block of the Jacobian row
        x[p] = (((exm >> p) & 1u) && nz) ? arow[((dj + 1) * 3 + di + 1) * AX1_BLK + aoff] : 0.0;
      else
        x[p] = 0.0;              // level-1 fill entry
    }
  };
  // one elimination step against lower neighbour k_; branch-free (a missing neighbour is a zero block row)
  auto eliminate = [&](double (&x)[13], int k_, int L, int j, unsigned exm) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (k != k_) continue;
      const int dik = P1_DI(k), djk = P1_DJ(k);
      const int Ln = L + dik + 3 * djk;
      const int qn = (j + djk) - lev1_jlo(Ln, wc);
      const double* kb = ((exm >> k) & 1u) ? ring + ((size_t)(Ln & 3) * nq + qn) * AX1_RING1_16_ROW : ring;  // [Dinv, U7..U12]
      double xs[4], lr[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) xs[m] = shfl1_d(x[k], rowl + m);
      double sum = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) sum += xs[m] * kb[4 * m + c];
      x[k] = sum;
#pragma unroll
      for (int m = 0; m < 4; ++m) lr[m] = shfl1_d(sum, rowl + m);
#pragma unroll
      for (int t = 7; t < 13; ++t) {
        const int p2 = p1_idx(dik + P1_DI(t), djk + P1_DJ(t));
        if (p2 < 0) continue;
        const double* ub = kb + (t - 6) * 16;
        double su = 0.0;
#pragma unroll
        for (int m = 0; m < 4; ++m) su += lr[m] * ub[4 * m + c];
        x[p2] -= su;
      }
    }
  };

  issue(0);
  issue(1);
  issue(2);
  double xn[13];
  Row1 wn = level_row1(g, 0, nlev, i0, wc, q);
  unsigned exn = exists(wn);
  cp_async_wait<2>();
  cta_sync();  // row of level 0 staged, ring zeroed
  expand(xn, 0, exn);
  eliminate(xn, 0, 0, wn.j, exn);
  eliminate(xn, 1, 0, wn.j, exn);
  eliminate(xn, 2, 0, wn.j, exn);
  for (int L = 0; L < nlev; ++L) {
    double x[13];
#pragma unroll
    for (int p = 0; p < 13; ++p) x[p] = xn[p];
    const Row1 w = wn;
    const unsigned exm = exn;
    const bool act = w.v >= 0;
    cp_async_wait<1>();
    cta_sync();  // row of level L+1 staged; ring row of level L-1 visible
    wn = level_row1(g, L + 1, nlev, i0, wc, q);
    exn = exists(wn);
    // level L: (i+2,j-1), (i-2,j), (i-1,j)        level L+1: load, (i-1,j-1), (i,j-1), (i+1,j-1)
    eliminate(x, 3, L, w.j, exm);
    expand(xn, L + 1, exn);
    eliminate(x, 4, L, w.j, exm);
    eliminate(xn, 0, L + 1, wn.j, exn);
    eliminate(x, 5, L, w.j, exm);
    eliminate(xn, 1, L + 1, wn.j, exn);
    // pivot block of level L: adjugate inverse over its 16 owners (ilu_common.cuh); the last early step of level L+1
    // is independent of it
    const double a = act ? x[6] : ((r == c) ? 1.0 : 0.0);
    eliminate(xn, 2, L + 1, wn.j, exn);
    const double inv = invert4_adjugate16(a, pscratch + (size_t)(L & 1) * nq * 16, e, adj);
    x[6] = inv;
    if (act) {
      const size_t p = sg.row_off + (size_t)(sg.lvl[L] + q);
#pragma unroll
      for (int k = 0; k < 6; ++k) LF[p * 96 + k * 16 + e] = x[k];
#pragma unroll
      for (int k = 6; k < 13; ++k) LB[p * 112 + (k - 6) * 16 + e] = x[k];
      double* kr = ring + ((size_t)(L & 3) * nq + q) * AX1_RING1_16_ROW;  // slot last read in iteration L-1
#pragma unroll
      for (int k = 6; k < 13; ++k) kr[(k - 6) * 16 + e] = x[k];
    }
    issue(L + 3);
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
// v = (LU)^-1 d, forward and backward sweep in one launch; same software pipeline as ilu_apply_kernel
// (solver.cu): every lane streams its scalar row of the next D levels' blocks into a private
// shared-memory slot with cp.async, solution values of the last levels go through a small ring.
template <int D, bool P2P>
__global__ void __launch_bounds__(256) ilu1_apply_kernel(Dev g, const double* __restrict__ LF,
                                                         const double* __restrict__ LB, const double* __restrict__ d,
                                                         double* v_out, int slab_w, int nq,
                                                         const int* __restrict__ lvl_tab, HaloPush hp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nt = 4 * nq, tid = threadIdx.x;  // nq = rows a level can hold; threads beyond 4 nq never own a row
  const int q = tid >> 2, r = tid & 3;
  double* stageM = reinterpret_cast<double*>(smem_raw);   // [D][7][nt][4]
  double* stageR = stageM + (size_t)D * 7 * nt * 4;       // [D][nt]
  double* ring = stageR + (size_t)D * nt;                 // [8][nq][4]
  const Slab1 sg = slab1_geom(g, slab_w, lvl_tab, reinterpret_cast<int*>(ring + 8 * (size_t)nq * 4));
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;

  // ---------------- forward: y = L^-1 d ----------------
  auto issue_fwd = [&](int L) {
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = L % D;
      const double* base = LF + (sg.row_off + (size_t)(sg.lvl[L] + q)) * 96 + r * 4;
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int di = P1_DI(k), dj = P1_DJ(k);
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0)) {
          double* dst = stageM + ((size_t)(st * 7 + k) * nt + tid) * 4;
          cp_async16(dst, base + k * 16);
          cp_async16(dst + 2, base + k * 16 + 2);
        }
      }
      cp_async8(stageR + st * nt + tid, d + 4 * (size_t)w.v + r);
    }
    cp_async_commit();
  };
  for (int L = 0; L < D; ++L) issue_fwd(L);
  for (int L = 0; L < nlev; ++L) {
    cp_async_wait<D - 1>();
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = L % D;
      double acc = stageR[st * nt + tid];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int di = P1_DI(k), dj = P1_DJ(k);
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj >= 0)) {
          const double* lp = stageM + ((size_t)(st * 7 + k) * nt + tid) * 4;
          const double2 l01 = *reinterpret_cast<const double2*>(lp);
          const double2 l23 = *reinterpret_cast<const double2*>(lp + 2);
          const int Ln = L + di + 3 * dj;
          const int qn = (w.j + dj) - lev1_jlo(Ln, wc);
          const double* yp = ring + ((size_t)(Ln & 7) * nq + qn) * 4;
          const double2 y01 = *reinterpret_cast<const double2*>(yp);
          const double2 y23 = *reinterpret_cast<const double2*>(yp + 2);
          double t = l01.x * y01.x;
          t += l01.y * y01.y;
          t += l23.x * y23.x;
          t += l23.y * y23.y;
          acc -= t;
        }
      }
      ring[((size_t)(L & 7) * nq + q) * 4 + r] = acc;
      v_out[4 * (size_t)w.v + r] = acc;
    }
    cta_sync();
    issue_fwd(L + D);
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---------------- backward: x = U^-1 y ----------------
  auto issue_bwd = [&](int L) {  // L counts down; stage index by distance from the top level
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = (nlev - 1 - L) % D;
      const double* base = LB + (sg.row_off + (size_t)(sg.lvl[L] + q)) * 112 + r * 4;
#pragma unroll
      for (int k = 6; k < 13; ++k) {
        const int di = P1_DI(k), dj = P1_DJ(k);
        if (k == 6 || ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj < g.nvy))) {
          double* dst = stageM + ((size_t)(st * 7 + (k - 6)) * nt + tid) * 4;
          cp_async16(dst, base + (k - 6) * 16);
          cp_async16(dst + 2, base + (k - 6) * 16 + 2);
        }
      }
      cp_async8(stageR + st * nt + tid, v_out + 4 * (size_t)w.v + r);  // own forward result
    }
    cp_async_commit();
  };
  for (int k = 0; k < D; ++k) issue_bwd(nlev - 1 - k);
  for (int L = nlev - 1; L >= 0; --L) {
    cp_async_wait<D - 1>();
    const Row1 w = level_row1(g, L, nlev, i0, wc, q);
    if (w.v >= 0) {
      const int st = (nlev - 1 - L) % D;
      double acc = stageR[st * nt + tid];
#pragma unroll
      for (int k = 7; k < 13; ++k) {
        const int di = P1_DI(k), dj = P1_DJ(k);
        if ((w.il + di >= 0) && (w.il + di < wc) && (w.j + dj < g.nvy)) {
          const double* up = stageM + ((size_t)(st * 7 + (k - 6)) * nt + tid) * 4;
          const double2 u01 = *reinterpret_cast<const double2*>(up);
          const double2 u23 = *reinterpret_cast<const double2*>(up + 2);
          const int Ln = L + di + 3 * dj;
          const int qn = (w.j + dj) - lev1_jlo(Ln, wc);
          const double* xp = ring + ((size_t)(Ln & 7) * nq + qn) * 4;
          const double2 x01 = *reinterpret_cast<const double2*>(xp);
          const double2 x23 = *reinterpret_cast<const double2*>(xp + 2);
          double t = u01.x * x01.x;
          t += u01.y * x01.y;
          t += u23.x * x23.x;
          t += u23.y * x23.y;
          acc -= t;
        }
      }
      const double* dp = stageM + ((size_t)(st * 7 + 0) * nt + tid) * 4;
      const double2 d01 = *reinterpret_cast<const double2*>(dp);
      const double2 d23 = *reinterpret_cast<const double2*>(dp + 2);
      const unsigned qmask = 0xFu << (tid & 28);
      const int qb = tid & 28;
      double xo = d01.x * __shfl_sync(qmask, acc, qb + 0);
      xo += d01.y * __shfl_sync(qmask, acc, qb + 1);
      xo += d23.x * __shfl_sync(qmask, acc, qb + 2);
      xo += d23.y * __shfl_sync(qmask, acc, qb + 3);
      ring[((size_t)(L & 7) * nq + q) * 4 + r] = xo;
      v_out[4 * (size_t)w.v + r] = xo;
      if (P2P) halo_push_value(hp, g.nvx, g.nvy, i0 + w.il, w.j, r, xo);
    }
    cta_sync();
    issue_bwd(L - D);
  }
  cp_async_wait<0>();
  if (P2P) halo_push_finish(hp, g.nvx, i0, wc);
}

// ------------------------------------------------------------------------------------------
// The same sweeps as a warp-specialised bulk-TMA pipeline (see ilu_apply_tma_kernel in solver.cu for the
// scheme): a producer warp brings each wavefront level of the packed factor (rows x 768 B forward, rows x 896 B
// backward) into a shared-memory stage with one cp.async.bulk and gathers the level's right-hand side; consumer
// quad q owns the column triple (3q, 3q+1, 3q+2) of the sub-slab (column 3q + L mod 3 at level L), the ring of
// recent solution values is indexed by column triple and zero-padded, blocks of absent neighbours are zeros.
template <int D, bool P2P>
__global__ void __launch_bounds__(288) ilu1_apply_tma_kernel(Dev g, const double* __restrict__ LF,
                                                             const double* __restrict__ LB, const double* __restrict__ d,
                                                             double* v_out, int slab_w, const int* __restrict__ lvl_tab,
                                                             HaloPush hp) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nt = blockDim.x - 32;                                     // consumer threads; the last warp produces
  const int nq = nt >> 2;
  double* stage = reinterpret_cast<double*>(smem_raw);               // [D][nq * 112]
  double* stageR = stage + (size_t)D * nq * 112;                      // [D][nt]
  double* ring = stageR + (size_t)D * nt;                             // [8][nq + 2][4]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + 8 * (size_t)(nq + 2) * 4);   // [D]
  unsigned long long* empty = full + D;                                                                 // [D]
  int* lvl_sh = reinterpret_cast<int*>(empty + D);
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) { mbar_init(&full[k], 33); mbar_init(&empty[k], (unsigned)nt >> 5); }
    mbar_init_fence();
  }
  for (int k = tid; k < 8 * (nq + 2) * 4; k += blockDim.x) ring[k] = 0.0;
  const Slab1 sg = slab1_geom(g, slab_w, lvl_tab, lvl_sh);            // ends with __syncthreads
  const int i0 = sg.i0, wc = sg.wc, nlev = sg.nlev;

  if (tid >= nt) {
    const int lane = tid - nt;
    auto produce = [&](int k) {
      const int st = k % D, use = k / D;
      if (use > 0) mbar_wait(&empty[st], (unsigned)(use - 1) & 1u);
      const bool fwd = k < nlev;
      const int L = fwd ? k : 2 * nlev - 1 - k;
      if (lane == 0) {
        const unsigned rows = (unsigned)(sg.lvl[L + 1] - sg.lvl[L]);
        const unsigned bytes = rows * (fwd ? 768u : 896u);
        mbar_arrive_expect_tx(&full[st], bytes);
        const double* gsrc = fwd ? LF + (sg.row_off + (size_t)sg.lvl[L]) * 96 : LB + (sg.row_off + (size_t)sg.lvl[L]) * 112;
        bulk_g2s(stage + (size_t)st * nq * 112, gsrc, bytes, &full[st]);
      }
      const double* src = fwd ? d : v_out;
      const int m3 = L % 3;
      for (int t = lane; t < nt; t += 32) {
        const int il = 3 * (t >> 2) + m3, tj = L - il, j = tj / 3;
        if (il < wc && tj >= 0 && j < g.nvy)
          cp_async8(stageR + st * nt + t, src + 4 * (size_t)(i0 + il + g.nvx * j) + (t & 3));
      }
      cp_async_mbar_arrive_noinc(&full[st]);
    };
    for (int k = 0; k < nlev; ++k) produce(k);
    __syncthreads();          // the consumers' forward results are in global memory
    for (int k = nlev; k < 2 * nlev; ++k) produce(k);
  } else {
    // Quad q owns the vertex-column triple (3q, 3q+1, 3q+2): column 3q + L % 3 at level L, grid row L / 3 - q. The level
    // loop is unrolled by three so that the column phase c3 is a compile-time constant (neighbour quads and block
    // slots fold), shared memory is addressed through 32-bit shared-window addresses, and stage / barrier / phase
    // are running values, as in ilu_apply_tma_kernel (solver.cu).
    const int q = tid >> 2, r = tid & 3;
    const unsigned full0 = smem_u32(full), empty0 = smem_u32(empty);
    const bool lane0 = (tid & 31) == 0, one_warp = nt == 32;
    const unsigned stage_u = smem_u32(stage) + (unsigned)r * 32u;
    const unsigned rhs_u = smem_u32(stageR) + (unsigned)tid * 8u;
    const unsigned rstride_b = (unsigned)(nq + 2) * 32u;
    const unsigned ringq_u = smem_u32(ring) + (unsigned)(q + 1) * 32u;
    const unsigned ringown_u = ringq_u + (unsigned)r * 8u;
    const unsigned stage_stride_b = (unsigned)nq * 896u, rhs_stride_b = (unsigned)nt * 8u;
    const long long vstride = 4LL * g.nvx;
    const long long vbase = 4LL * (i0 + 3 * q) + r;                    // column 3q, grid row 0
    int cap[3];
    bool ok[3];
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) { cap[c3] = (wc - c3 + 2) / 3 - 1; ok[c3] = 3 * q + c3 < wc; }  // stage row = min(L/3, cap) - q
    int st = 0;
    unsigned ph = 0, st_stage = 0, st_rhs = 0, st_bar = 0;
    auto advance = [&]() {
      st_stage += stage_stride_b; st_rhs += rhs_stride_b; st_bar += 8u;
      if (++st == D) { st = 0; st_stage = 0; st_rhs = 0; st_bar = 0; ph ^= 1u; }
    };
    auto level_done = [&]() {
      if (one_warp) __syncwarp();
      else named_bar_sync(1, nt);
    };
    // ---------------- forward: y = L^-1 d ----------------
    for (int L0 = 0, m3 = 0; L0 < nlev; L0 += 3, ++m3) {
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        const int L = L0 + c3;
        if (L >= nlev) break;
        const int j = m3 - q;
        mbar_wait_a(full0 + st_bar, ph);
        const bool act = ok[c3] && (unsigned)j < (unsigned)g.nvy;
        double2 lv[6][2], yv[6][2];
        double acc = 0.0;
        if (act) {
          const unsigned la = stage_u + st_stage + (unsigned)(min(m3, cap[c3]) - q) * 768u;
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int di = P1_DI(k), dj = P1_DJ(k);
            const int dq = (c3 + di + 3) / 3 - 1;                       // neighbour's quad relative to mine
            const unsigned yp = ringq_u + ((unsigned)(L + di + 3 * dj) & 7u) * rstride_b + (unsigned)(dq * 32);
            lv[k][0] = lds_d2(la + k * 128u); lv[k][1] = lds_d2(la + k * 128u + 16u);
            yv[k][0] = lds_d2(yp); yv[k][1] = lds_d2(yp + 16u);
          }
          acc = lds_d(rhs_u + st_rhs);
        }
        __syncwarp();
        if (lane0) mbar_arrive_a(empty0 + st_bar);
        if (act) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            double t = lv[k][0].x * yv[k][0].x;
            t += lv[k][0].y * yv[k][0].y;
            t += lv[k][1].x * yv[k][1].x;
            t += lv[k][1].y * yv[k][1].y;
            acc -= t;
          }
          sts_d(ringown_u + ((unsigned)L & 7u) * rstride_b, acc);
          v_out[vbase + 4 * c3 + vstride * j] = acc;
        }
        level_done();
        advance();
      }
    }
    __syncthreads();
    // ---------------- backward: x = U^-1 y ----------------
    for (int m3 = (nlev - 1) / 3; m3 >= 0; --m3) {
#pragma unroll
      for (int c3 = 2; c3 >= 0; --c3) {
        const int L = 3 * m3 + c3;
        if (L >= nlev) continue;
        const int j = m3 - q;
        mbar_wait_a(full0 + st_bar, ph);
        const bool act = ok[c3] && (unsigned)j < (unsigned)g.nvy;
        double2 uv[7][2], xv[6][2];
        double acc = 0.0;
        if (act) {
          const unsigned ua = stage_u + st_stage + (unsigned)(min(m3, cap[c3]) - q) * 896u;
#pragma unroll
          for (int p = 6; p < 13; ++p) {
            uv[p - 6][0] = lds_d2(ua + (p - 6) * 128u); uv[p - 6][1] = lds_d2(ua + (p - 6) * 128u + 16u);
            if (p > 6) {
              const int di = P1_DI(p), dj = P1_DJ(p);
              const int dq = (c3 + di + 3) / 3 - 1;
              const unsigned xp = ringq_u + ((unsigned)(L + di + 3 * dj) & 7u) * rstride_b + (unsigned)(dq * 32);
              xv[p - 7][0] = lds_d2(xp); xv[p - 7][1] = lds_d2(xp + 16u);
            }
          }
          acc = lds_d(rhs_u + st_rhs);
        }
        __syncwarp();
        if (lane0) mbar_arrive_a(empty0 + st_bar);
        if (act) {
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            double t = uv[p + 1][0].x * xv[p][0].x;
            t += uv[p + 1][0].y * xv[p][0].y;
            t += uv[p + 1][1].x * xv[p][1].x;
            t += uv[p + 1][1].y * xv[p][1].y;
            acc -= t;
          }
          const unsigned qmask = 0xFu << (tid & 28);
          const int qb = tid & 28;
          double xo = uv[0][0].x * __shfl_sync(qmask, acc, qb + 0);
          xo += uv[0][0].y * __shfl_sync(qmask, acc, qb + 1);
          xo += uv[0][1].x * __shfl_sync(qmask, acc, qb + 2);
          xo += uv[0][1].y * __shfl_sync(qmask, acc, qb + 3);
          sts_d(ringown_u + ((unsigned)L & 7u) * rstride_b, xo);
          v_out[vbase + 4 * c3 + vstride * j] = xo;
          if (P2P) halo_push_value(hp, g.nvx, g.nvy, i0 + 3 * q + c3, j, r, xo);
        }
        level_done();
        advance();
      }
    }
  }
  if (P2P) halo_push_finish(hp, g.nvx, i0, wc);
}

// ------------------------------------------------------------------------------------------
// one quad per wavefront row: a level of a slab of width w has at most ceil(w/3) rows
static int slab1_threads(int slab_w) {
  int rows = (slab_w + 2) / 3;
  int t = 4 * rows;
  t = ((t + 31) / 32) * 32;
  if (t > 256) t = 256;
  if (t < 32) t = 32;
  return t;
}
static int slab1_rows(int slab_w) { return (slab_w + 2) / 3; }
static size_t factor1_smem(const Dev& g, int slab_w) {
  const int nq = slab1_rows(slab_w);
  return (size_t)nq * (2 * AX1_ROW + 4 * AX1_RING1_ROW) * sizeof(double) + (size_t)(slab_w + 3 * g.nvy) * sizeof(int);
}
// widest sub-slab the factor kernel can hold in shared memory for this grid (<= 192: 64 quads per CTA)
int ilu1_max_slab_w(const Dev& g) {
  int w = 192;
  while (w > 3 && factor1_smem(g, w) > 227 * 1024) w -= 3;
  return w;
}
// level prefix tables for full-width sub-slabs and for the (narrower) last one: 2 x (slab_w + 3 Ny + 1) ints
void ilu1_level_tables(const Dev& g, int slab_w, std::vector<int>& tab) {
  const int stride = slab_w + 3 * (g.nvy - 1) + 1;
  tab.assign(2 * (size_t)stride, 0);
  const int last_wc = g.nvx - ((g.nvx - 1) / slab_w) * slab_w;
  const int wcs[2] = {slab_w, last_wc};
  for (int t = 0; t < 2; ++t) {
    const int wc = wcs[t], nlev = wc + 3 * (g.nvy - 1);
    int acc = 0;
    for (int L = 0; L < nlev; ++L) {
      tab[(size_t)t * stride + L] = acc;
      const int a = L - wc + 1;
      const int jlo = a <= 0 ? 0 : (a + 2) / 3;
      const int jhi = std::min(g.nvy - 1, L / 3);
      if (jhi >= jlo) acc += jhi - jlo + 1;
    }
    tab[(size_t)t * stride + nlev] = acc;
  }
}
void launch_ilu1_factor(const Dev& g, const double* A, double* LU, int slab_w, const int* lvl_tab, cudaStream_t st) {
  const int nslab = (g.nvx + slab_w - 1) / slab_w;
  static const int f16_max_w = getenv("AX1_ILU1_FACTOR16_W") ? atoi(getenv("AX1_ILU1_FACTOR16_W")) : 24;
  if (slab_w <= f16_max_w && slab1_rows(slab_w) <= 16) {
    const int nq = (slab1_rows(slab_w) + 1) & ~1;  // rows per level, even: whole warps
    const size_t smem = (size_t)nq * (3 * AX1_ROW + 4 * AX1_RING1_16_ROW + 32) * sizeof(double) + (size_t)(slab_w + 3 * g.nvy) * sizeof(int);
    static bool attr16 = false;
    if (!attr16) {
      cudaFuncSetAttribute(ilu1_factor16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      attr16 = true;
    }
    ilu1_factor16_kernel<<<nslab, 16 * nq, smem, st>>>(g, A, LU, LU + (size_t)g.nv * 96, slab_w, nq, lvl_tab);
    return;
  }
  const int nt = slab1_threads(slab_w);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ilu1_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  ilu1_factor_kernel<<<nslab, nt, factor1_smem(g, slab_w), st>>>(g, A, LU, LU + (size_t)g.nv * 96, slab_w,
                                                                 slab1_rows(slab_w), lvl_tab);
}
int ilu_pick_stages(int nt, int bytes_per_thread_stage);  // solver.cu
template <int D, bool P2P>
static void launch_ilu1_apply_t(const Dev& g, const double* LU, const double* d, double* v, int slab_w,
                                const int* lvl_tab, const HaloPush& hp, cudaStream_t st) {
  const int nslab = (g.nvx + slab_w - 1) / slab_w;
  const int nt = slab1_threads(slab_w), nq = slab1_rows(slab_w);
  static const int use_tma = getenv("AX1_ILU_TMA") ? atoi(getenv("AX1_ILU_TMA")) : 1;
  // narrow sub-slabs (latency-bound small grids) take the warp-specialised TMA pipeline; at 128+ columns the sweep
  // is bandwidth-bound and the cp.async kernel below measured faster for ILU(1) (5.6 vs 7.0 ms at 16.9 M vertices)
  if (use_tma && nt <= 64) {
    const int nqc = nt / 4;
    const size_t smem_t = (size_t)D * nqc * 896 + (size_t)D * nt * 8 + 8 * (size_t)(nqc + 2) * 32 + (size_t)D * 16 +
                          (size_t)(slab_w + 3 * g.nvy) * sizeof(int);
    if (smem_t <= 227 * 1024) {
      static bool attr_t = false;
      if (!attr_t) {
        cudaFuncSetAttribute(ilu1_apply_tma_kernel<D, P2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_t = true;
      }
      ilu1_apply_tma_kernel<D, P2P><<<nslab, nt + 32, smem_t, st>>>(g, LU, LU + (size_t)g.nv * 96, d, v, slab_w, lvl_tab, hp);
      return;
    }
  }
  const size_t smem = (size_t)D * 4 * nq * (7 * 32 + 8) + 8 * (size_t)nq * 32 + (size_t)(slab_w + 3 * g.nvy) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(ilu1_apply_kernel<D, P2P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  ilu1_apply_kernel<D, P2P><<<nslab, nt, smem, st>>>(g, LU, LU + (size_t)g.nv * 96, d, v, slab_w, nq, lvl_tab, hp);
}
template <bool P2P>
static void launch_ilu1_apply_d(const Dev& g, const double* LU, const double* d, double* v, int slab_w,
                                const int* lvl_tab, const HaloPush& hp, cudaStream_t st) {
  switch (ilu_pick_stages(4 * slab1_rows(slab_w), 7 * 32 + 8)) {
    case 8: launch_ilu1_apply_t<8, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st); break;
    case 6: launch_ilu1_apply_t<6, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st); break;
    case 4: launch_ilu1_apply_t<4, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st); break;
    case 3: launch_ilu1_apply_t<3, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st); break;
    default: launch_ilu1_apply_t<2, P2P>(g, LU, d, v, slab_w, lvl_tab, hp, st); break;
  }
}
void launch_ilu1_apply(const Dev& g, const double* LU, const double* d, double* v, int slab_w, const int* lvl_tab,
                       cudaStream_t st, const HaloPush* hp) {
  if (hp) launch_ilu1_apply_d<true>(g, LU, d, v, slab_w, lvl_tab, *hp, st);
  else launch_ilu1_apply_d<false>(g, LU, d, v, slab_w, lvl_tab, HaloPush{}, st);
}

}  // namespace ax1
