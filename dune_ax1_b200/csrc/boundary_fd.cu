// boundary_fd.cu -- finite-difference Jacobian of the membrane-interface boundary term.
//
// The reference has no jacobian_boundary: PDELab's NumericalJacobianBoundary differentiates
// alpha_boundary (acme2_cyl_operator_fully_coupled.hh:29-36, 676-926) with
//   delta = 1e-7 (1 + |u_j|),  J_ij += (alpha_i(u + delta e_j) - alpha_i(u)) / delta
// (formula confirmed by stuff/idefault_patch.dat:14-23), perturbing only the 16 local
// coefficients -- the opposite membrane side is NOT differentiated (quasi-Newton).
// This file reproduces exactly that, row-owner style, and is compiled with --fmad=false so that
// the difference quotient sees the same unfused arithmetic as a CPU build (rounding differences are
// amplified by 1/delta ~ 1e7).
#include "kernels.h"
#include "pnp_device.cuh"

namespace ax1 {

__global__ void jacobian_boundary_fd_kernel(Dev g, const double* __restrict__ u, double* A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * g.nvx) return;
  const int side = t / g.nvx;  // 0: cytosol-side interface row jm, 1: ES-side row jm+1
  const int i = t % g.nvx;
  const int jr = g.jm + side;
  const int cj = side == 0 ? g.jm - 1 : g.jm + 1;
  const int v = i + g.nvx * jr;
  const uint8_t rmask = g.dmask[v];
  const double vs = g.do_volume_scaling ? 1. / g.node_vol[v] : 1.0;
  const double fd_eps = 1e-7;
  for (int a = 0; a < 2; ++a) {
    const int ci = i - 1 + a;
    if (ci < 0 || ci >= g.nx) continue;
    CellCtx c;
    load_cell(g, u, ci, cj, c);
    FaceCtx f;
    make_face(g, u, c, f);
    const int vbase = side == 0 ? 2 : 0;   // local numbers of the two face vertices
    const int k = vbase + (i - ci);        // our test function
    double down[3] = {0.0, 0.0, 0.0};
    elec_boundary_row(g, f, ci, c.xl, k, vs, down);
    for (int comp = 0; comp < 4; ++comp)
      for (int fv = 0; fv < 2; ++fv) {
        const int lc = comp * 4 + vbase + fv;
        const int colv_i = ci + fv;
        const double orig = c.xl[lc];
        const double delta = fd_eps * (1.0 + fabs(orig));
        c.xl[lc] += delta;
        double up[3] = {0.0, 0.0, 0.0};
        elec_boundary_row(g, f, ci, c.xl, k, vs, up);
        c.xl[lc] = orig;
        const int s = 3 + (colv_i - i + 1);
        double* blk = A + ((size_t)v * 9 + s) * AX1_BLK;
        // the flux of species j depends on c_j and phi only: columns other than j and pot get
        // (up - down) == 0 exactly, which keeps the arrow shape of the block
        for (int j = 0; j < 3; ++j) {
          if ((rmask >> j) & 1) continue;   // constrained row
          if (comp == j) blk[2 * j] += (up[j] - down[j]) / delta;
          else if (comp == 3) blk[2 * j + 1] += (up[j] - down[j]) / delta;
        }
      }
  }
}

void launch_jacobian_boundary_fd(const Dev& g, const double* u, double* A, cudaStream_t st) {
  if (!g.membrane_active) return;
  const int bs = 128;
  jacobian_boundary_fd_kernel<<<(2 * g.nvx + bs - 1) / bs, bs, 0, st>>>(g, u, A);
}

}  // namespace ax1
