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
#include "kernels.h"

namespace ax1 {

__device__ __forceinline__ void ldg4(const double* p, double* o) {
  const double2 lo = __ldg(reinterpret_cast<const double2*>(p));
  const double2 hi = __ldg(reinterpret_cast<const double2*>(p + 2));
  o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
}
// L2-coherent loads for data written earlier in the same kernel by other threads of the CTA
__device__ __forceinline__ void ldcg4(const double* p, double* o) {
  const double2 lo = __ldcg(reinterpret_cast<const double2*>(p));
  const double2 hi = __ldcg(reinterpret_cast<const double2*>(p + 2));
  o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
}
__device__ __forceinline__ void st4(double* p, const double* v) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

// ------------------------------------------------------------------------------------------
// y = A x ; 4 lanes per block row, lane r = scalar row r
__global__ void __launch_bounds__(256) spmv_kernel(Dev g, const double* __restrict__ A, const double* __restrict__ x,
                                                   double* __restrict__ y, int zero_con) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(t >> 2), r = (int)(t & 3);
  if (v >= g.nv) return;
  const int i = v % g.nvx, j = v / g.nvx;
  const double* Arow = A + (size_t)v * 144 + r * 4;
  double acc = 0.0;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const int di = s % 3 - 1, dj = s / 3 - 1;
    const int ii = i + di, jj = j + dj;
    if (ii < 0 || ii >= g.nvx || jj < 0 || jj >= g.nvy) continue;
    double a[4], xv[4];
    ldg4(Arow + s * 16, a);
    ldg4(x + 4 * (size_t)(v + di + dj * g.nvx), xv);
    acc += a[0] * xv[0];
    acc += a[1] * xv[1];
    acc += a[2] * xv[2];
    acc += a[3] * xv[3];
  }
  if (zero_con && ((g.dmask[v] >> r) & 1)) acc = 0.0;
  y[4 * (size_t)v + r] = acc;
}

// ------------------------------------------------------------------------------------------
// 4x4 inverse, Gauss-Jordan with partial pivoting, fully unrolled (static register indexing)
__device__ __forceinline__ void invert4(double a[4][4], double inv[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) inv[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int r = c + 1; r < 4; ++r) {
      if (fabs(a[r][c]) > fabs(a[c][c])) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          double t = a[c][k]; a[c][k] = a[r][k]; a[r][k] = t;
          t = inv[c][k]; inv[c][k] = inv[r][k]; inv[r][k] = t;
        }
      }
    }
    const double p = 1.0 / a[c][c];
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[c][k] *= p; inv[c][k] *= p; }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      const double f = a[r][c];
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[r][k] -= f * a[c][k]; inv[r][k] -= f * inv[c][k]; }
    }
  }
}

__device__ __forceinline__ int lev_jlo(int L, int wc) {
  const int a = L - wc + 1;
  return a <= 0 ? 0 : (a + 1) / 2;
}

// ------------------------------------------------------------------------------------------
// block ILU0 of one sub-slab per CTA (bilu0_decomposition: L_ij = A_ij * A_jj^-1, diagonal stored inverted)
__global__ void __launch_bounds__(256) ilu_factor_kernel(Dev g, const double* __restrict__ A, double* LU, int slab_w) {
  const int i0 = blockIdx.x * slab_w;
  const int wc = min(slab_w, g.nvx - i0);
  const int q = threadIdx.x >> 2, r = threadIdx.x & 3, nq = blockDim.x >> 2;
  const unsigned qmask = 0xFu << (threadIdx.x & 28);
  const int nlev = wc + 2 * (g.nvy - 1);
  for (int L = 0; L < nlev; ++L) {
    const int jlo = lev_jlo(L, wc), jhi = min(g.nvy - 1, L >> 1);
    for (int j = jlo + q; j <= jhi; j += nq) {
      const int il = L - 2 * j;
      const int v = i0 + il + g.nvx * j;
      bool ex[9];
      double row[9][4];
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        ex[s] = (il + di >= 0) && (il + di < wc) && (j + dj >= 0) && (j + dj < g.nvy);
        if (ex[s]) ldg4(A + ((size_t)v * 9 + s) * 16 + r * 4, row[s]);
        else row[s][0] = row[s][1] = row[s][2] = row[s][3] = 0.0;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (!ex[s]) continue;
        const int dis = s % 3 - 1, djs = s / 3 - 1;
        const size_t k = (size_t)(v + dis + djs * g.nvx);
        double lrow[4];
        {
          double dinv[16];
#pragma unroll
          for (int m = 0; m < 4; ++m) ldcg4(LU + (k * 9 + 4) * 16 + m * 4, dinv + 4 * m);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum += row[s][m] * dinv[4 * m + c];
            lrow[c] = sum;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) row[s][c] = lrow[c];
#pragma unroll
        for (int tt = 5; tt < 9; ++tt) {
          const int di = dis + (tt % 3 - 1), dj = djs + (tt / 3 - 1);
          if (di < -1 || di > 1 || dj < -1 || dj > 1) continue;
          const int s2 = (dj + 1) * 3 + (di + 1);
          if (!ex[s2]) continue;
          double ub[16];
#pragma unroll
          for (int m = 0; m < 4; ++m) ldcg4(LU + (k * 9 + tt) * 16 + m * 4, ub + 4 * m);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            double sum = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) sum += lrow[m] * ub[4 * m + c];
            row[s2][c] -= sum;
          }
        }
      }
      // invert the pivot block (every lane of the quad holds one row of it)
      double d[4][4], dinv[4][4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int c = 0; c < 4; ++c) d[m][c] = __shfl_sync(qmask, row[4][c], (threadIdx.x & 28) + m);
      invert4(d, dinv);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double val = dinv[0][c];
        if (r == 1) val = dinv[1][c];
        if (r == 2) val = dinv[2][c];
        if (r == 3) val = dinv[3][c];
        row[4][c] = val;
      }
#pragma unroll
      for (int s = 0; s < 9; ++s) st4(LU + ((size_t)v * 9 + s) * 16 + r * 4, row[s]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// v = (LU)^-1 d for one sub-slab per CTA (bilu_backsolve): forward sweep with unit lower blocks,
// backward sweep with the stored inverse diagonal; both sweeps in one launch.
__global__ void __launch_bounds__(256) ilu_apply_kernel(Dev g, const double* __restrict__ LU, const double* __restrict__ d,
                                                        double* v_out, int slab_w, int zero_con) {
  const int i0 = blockIdx.x * slab_w;
  const int wc = min(slab_w, g.nvx - i0);
  const int q = threadIdx.x >> 2, r = threadIdx.x & 3, nq = blockDim.x >> 2;
  const unsigned qmask = 0xFu << (threadIdx.x & 28);
  const int nlev = wc + 2 * (g.nvy - 1);
  for (int L = 0; L < nlev; ++L) {
    const int jlo = lev_jlo(L, wc), jhi = min(g.nvy - 1, L >> 1);
    for (int j = jlo + q; j <= jhi; j += nq) {
      const int il = L - 2 * j;
      const int v = i0 + il + g.nvx * j;
      double acc = d[4 * (size_t)v + r];
      if (zero_con && ((g.dmask[v] >> r) & 1)) acc = 0.0;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if (il + di < 0 || il + di >= wc || j + dj < 0) continue;
        double l[4], yk[4];
        ldg4(LU + ((size_t)v * 9 + s) * 16 + r * 4, l);
        ldcg4(v_out + 4 * (size_t)(v + di + dj * g.nvx), yk);
        acc -= l[0] * yk[0];
        acc -= l[1] * yk[1];
        acc -= l[2] * yk[2];
        acc -= l[3] * yk[3];
      }
      v_out[4 * (size_t)v + r] = acc;
    }
    __syncthreads();
  }
  for (int L = nlev - 1; L >= 0; --L) {
    const int jlo = lev_jlo(L, wc), jhi = min(g.nvy - 1, L >> 1);
    for (int j = jlo + q; j <= jhi; j += nq) {
      const int il = L - 2 * j;
      const int v = i0 + il + g.nvx * j;
      double acc = __ldcg(v_out + 4 * (size_t)v + r);
#pragma unroll
      for (int s = 5; s < 9; ++s) {
        const int di = s % 3 - 1, dj = s / 3 - 1;
        if (il + di < 0 || il + di >= wc || j + dj >= g.nvy) continue;
        double ub[4], xk[4];
        ldg4(LU + ((size_t)v * 9 + s) * 16 + r * 4, ub);
        ldcg4(v_out + 4 * (size_t)(v + di + dj * g.nvx), xk);
        acc -= ub[0] * xk[0];
        acc -= ub[1] * xk[1];
        acc -= ub[2] * xk[2];
        acc -= ub[3] * xk[3];
      }
      double dv[4];
      ldg4(LU + ((size_t)v * 9 + 4) * 16 + r * 4, dv);
      double xr = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) xr += dv[c] * __shfl_sync(qmask, acc, (threadIdx.x & 28) + c);
      v_out[4 * (size_t)v + r] = xr;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
__global__ void rownorm_kernel(Dev g, double* A, double* rv) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int v = (int)(t >> 2), r = (int)(t & 3);
  if (v >= g.nv) return;
  double* Arow = A + (size_t)v * 144 + r * 4;
  double n = 0.0;
#pragma unroll
  for (int c = 0; c < 4; ++c) n = fmax(n, fabs(Arow[4 * 16 + c]));
  if (n) {
#pragma unroll
    for (int s = 0; s < 9; ++s)
#pragma unroll
      for (int c = 0; c < 4; ++c) Arow[s * 16 + c] /= n;
    rv[4 * (size_t)v + r] /= n;
  } else {
    Arow[4 * 16 + r] = 1.0;
  }
}

// 9-slot layout <-> compact BCRS (download / upload of the Jacobian in the reference's layout)
__global__ void bcrs_copy_kernel(Dev g, double* A9, double* blocks, const int* __restrict__ rowptr, int to_bcrs) {
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
  if (!mine) {
    if (!to_bcrs) A9[vs * 16 + e] = 0.0;
    return;
  }
  const size_t b = ((size_t)rowptr[v] + pos) * 16 + e;
  if (to_bcrs) blocks[b] = A9[vs * 16 + e];
  else A9[vs * 16 + e] = blocks[b];
}

// ------------------------------------------------------------------------------------------
void launch_spmv(const Dev& g, const double* A, const double* x, double* y, int zero_con, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 4LL * g.nv;
  spmv_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, A, x, y, zero_con);
}
static int slab_threads(int slab_w) {
  int rows = (slab_w + 1) / 2;
  int t = 4 * rows;
  t = ((t + 31) / 32) * 32;
  if (t > 256) t = 256;
  if (t < 32) t = 32;
  return t;
}
void launch_ilu_factor(const Dev& g, const double* A, double* LU, int slab_w, cudaStream_t st) {
  const int nslab = (g.nvx + slab_w - 1) / slab_w;
  ilu_factor_kernel<<<nslab, slab_threads(slab_w), 0, st>>>(g, A, LU, slab_w);
}
void launch_ilu_apply(const Dev& g, const double* LU, const double* d, double* v, int slab_w, int zero_con,
                      cudaStream_t st) {
  const int nslab = (g.nvx + slab_w - 1) / slab_w;
  ilu_apply_kernel<<<nslab, slab_threads(slab_w), 0, st>>>(g, LU, d, v, slab_w, zero_con);
}
void launch_rownorm(const Dev& g, double* A, double* r, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 4LL * g.nv;
  rownorm_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, A, r);
}
void launch_compact_bcrs(const Dev& g, const double* A9, double* blocks, const int* rowptr, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 9LL * 16 * g.nv;
  bcrs_copy_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, const_cast<double*>(A9), blocks, rowptr, 1);
}
void launch_expand_bcrs(const Dev& g, const double* blocks, double* A9, const int* rowptr, cudaStream_t st) {
  const int bs = 256;
  const long long nt = 9LL * 16 * g.nv;
  bcrs_copy_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, A9, const_cast<double*>(blocks), rowptr, 0);
}

}  // namespace ax1
