// amg.cu -- transfer operators of the aggregation multigrid preconditioner (AMG with ILU0 smoother):
// the device counterpart of the reference's myelin backend ISTLBackend_GMRES_AMG_ILU0
// (dune/ax1/common/ax1_solverbackend.hh:73-302: Dune::Amg::AMG, non-smoothed aggregation, SeqILU0
// smoother, selected at acme2_cyl_setup.hh:749-760).
//
// Dune's AMG builds aggregates greedily from a strength-of-connection graph [ext, dune-istl 2.3
// aggregates.hh]; that graph algorithm is neither in /root/reference nor reproducible bit for bit, and a
// preconditioner only changes iteration counts, not converged answers. The device version uses what the
// structure of this problem offers: the block-ILU0 smoother already resolves the stiff radial (y)
// direction, what it cannot do is carry information along the axon (x) further than one sub-slab per
// application. Aggregates are therefore `cx` consecutive vertex columns at full radial resolution
// (piecewise-constant prolongation per component, like Dune's non-smoothed aggregation), so every
// Galerkin coarse operator P^T A P is again a 9-point block stencil with arrow-shaped 4x4 blocks on a
// tensor grid with nvx/cx columns -- SpMV, ILU0 factor and ILU0 sweeps run unchanged on every level.
// Constrained (Dirichlet / processor-boundary) DOFs of the finest level get no coarse correction.
#include <algorithm>
#include "kernels.h"

namespace ax1 {

// A_c[(I,j),(I+DI,j+dj)] = sum over fine i in aggregate I, fine di with (i+di)/cx == I+DI of A_f[(i,j),(i+di,j+dj)],
// rows / columns of constrained fine DOFs left out; a coarse scalar row without any free fine DOF becomes a unit row.
__global__ void galerkin_x_kernel(int nvx_f, int nvx_c, int nvy, int cx, const double* __restrict__ Af, double* Ac,
                                  const uint8_t* __restrict__ dmask_f) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nvc = (long long)nvx_c * nvy;
  if (t >= nvc * AX1_ROW) return;
  const int vc = (int)(t / AX1_ROW), rem = (int)(t % AX1_ROW);
  const int sc = rem / AX1_BLK, e = rem % AX1_BLK;
  const int I = vc % nvx_c, j = vc / nvx_c;
  const int DI = sc % 3 - 1, dj = sc / 3 - 1;
  // arrow layout [d0 c0 | d1 c1 | d2 c2 | p0 p1 p2 p3] -> (row a, column b) of the 4x4 block
  const int a = e < 6 ? e / 2 : 3;
  const int b = e < 6 ? ((e & 1) ? 3 : e / 2) : e - 6;
  const int jj = j + dj;
  double sum = 0.0;
  int free_rows = 0;
  const int i_end = min(cx * I + cx, nvx_f);
  for (int i = cx * I; i < i_end; ++i) {
    const size_t vf = (size_t)i + (size_t)nvx_f * j;
    if (dmask_f && ((dmask_f[vf] >> a) & 1)) continue;
    ++free_rows;
    if (jj < 0 || jj >= nvy) continue;
    for (int di = -1; di <= 1; ++di) {
      const int ii = i + di;
      if (ii < 0 || ii >= nvx_f) continue;
      if (ii / cx - I != DI) continue;
      if (dmask_f && ((dmask_f[(size_t)ii + (size_t)nvx_f * jj] >> b) & 1)) continue;
      sum += Af[(vf * 9 + (dj + 1) * 3 + (di + 1)) * AX1_BLK + e];
    }
  }
  if (free_rows == 0 && sc == 4 && a == b) sum = 1.0;
  Ac[t] = sum;
}

__global__ void restrict_x_kernel(int nvx_f, int nvx_c, int nvy, int cx, const double* __restrict__ rf, double* bc,
                                  const uint8_t* __restrict__ dmask_f) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4LL * nvx_c * nvy) return;
  const int vc = (int)(t >> 2), a = (int)(t & 3);
  const int I = vc % nvx_c, j = vc / nvx_c;
  const int i_end = min(cx * I + cx, nvx_f);
  double s = 0.0;
  for (int i = cx * I; i < i_end; ++i) {
    const size_t vf = (size_t)i + (size_t)nvx_f * j;
    if (dmask_f && ((dmask_f[vf] >> a) & 1)) continue;
    s += rf[4 * vf + a];
  }
  bc[t] = s;
}

__global__ void prolong_add_x_kernel(int nvx_f, int nvx_c, int nvy, int cx, const double* __restrict__ xc, double* xf,
                                     const uint8_t* __restrict__ dmask_f) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4LL * nvx_f * nvy) return;
  const int vf = (int)(t >> 2), a = (int)(t & 3);
  if (dmask_f && ((dmask_f[vf] >> a) & 1)) return;
  const int i = vf % nvx_f, j = vf / nvx_f;
  xf[t] += xc[4 * ((size_t)(i / cx) + (size_t)nvx_c * j) + a];
}

__global__ void vec_add_kernel(int n, double* x, const double* __restrict__ t) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) x[k] += t[k];
}

// stand-alone halo push (multi-GPU with peer windows when the last kernel of the preconditioner is not a sweep)
__global__ void __launch_bounds__(1024) halo_push_kernel(Dev g, const double* __restrict__ v, HaloPush hp) {
  const int n = 3 * g.nvy * 4;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int k = t & 3, j = (t >> 2) % g.nvy, c = (t >> 2) / g.nvy;
    if (hp.dst[0]) hp.dst[0][t] = v[4 * (size_t)(c + g.nvx * j) + k];
    if (hp.dst[1]) hp.dst[1][t] = v[4 * (size_t)(g.nvx - 3 + c + g.nvx * j) + k];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (hp.dst[0]) st_release_sys(hp.flag[0], hp.seq);
    if (hp.dst[1]) st_release_sys(hp.flag[1], hp.seq);
  }
}

static unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
void launch_galerkin_x(int nvx_f, int nvx_c, int nvy, int cx, const double* Af, double* Ac, const uint8_t* dmask_f,
                       cudaStream_t st) {
  galerkin_x_kernel<<<nblk((long long)nvx_c * nvy * AX1_ROW, 256), 256, 0, st>>>(nvx_f, nvx_c, nvy, cx, Af, Ac, dmask_f);
}
void launch_restrict_x(int nvx_f, int nvx_c, int nvy, int cx, const double* rf, double* bc, const uint8_t* dmask_f,
                       cudaStream_t st) {
  restrict_x_kernel<<<nblk(4LL * nvx_c * nvy, 256), 256, 0, st>>>(nvx_f, nvx_c, nvy, cx, rf, bc, dmask_f);
}
void launch_prolong_add_x(int nvx_f, int nvx_c, int nvy, int cx, const double* xc, double* xf, const uint8_t* dmask_f,
                          cudaStream_t st) {
  prolong_add_x_kernel<<<nblk(4LL * nvx_f * nvy, 256), 256, 0, st>>>(nvx_f, nvx_c, nvy, cx, xc, xf, dmask_f);
}
void launch_vec_add(int n, double* x, const double* t, cudaStream_t st) {
  const unsigned nb = std::min<unsigned>(nblk(n, 256), 148 * 8);
  vec_add_kernel<<<nb, 256, 0, st>>>(n, x, t);
}
void launch_halo_push(const Dev& g, const double* v, const HaloPush& hp, cudaStream_t st) {
  halo_push_kernel<<<1, 1024, 0, st>>>(g, v, hp);
}

}  // namespace ax1
