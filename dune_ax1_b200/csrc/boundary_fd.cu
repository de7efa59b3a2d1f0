// boundary_fd.cu -- finite-difference Jacobians: the membrane-interface boundary term (always) and the
// volume terms of the operators configured without analytic jacobian_volume (the reference's default).
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

// One thread per (interface vertex row, perturbed component, column offset -1 / 0 / +1): the row's test function
// lives in the two cells left and right of the vertex, and a block column receives its difference quotients from
// cell a = 0 first and a = 1 second, which each thread keeps (same sums in the same order as a serial loop over
// a, component and face vertex). 24 threads share a row instead of one thread walking all 18 evaluations.
__global__ void __launch_bounds__(128) jacobian_boundary_fd_kernel(Dev g, const double* __restrict__ u, double* A) {
  const int tg = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = tg / 12, sub = tg % 12, comp = sub / 3, off = sub % 3 - 1;
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
    const int fv = i + off - ci;             // the face vertex of this cell that sits in my block column
    if (fv < 0 || fv > 1) continue;
    CellCtx c;
    load_cell(g, u, ci, cj, c);
    FaceCtx f;
    make_face(g, u, c, f);
    const int vbase = side == 0 ? 2 : 0;   // local numbers of the two face vertices
    const int k = vbase + (i - ci);        // our test function
    double down[3] = {0.0, 0.0, 0.0};
    elec_boundary_row(g, f, ci, c.xl, k, vs, down);
    const int lc = comp * 4 + vbase + fv;
    const double orig = c.xl[lc];
    const double delta = fd_eps * (1.0 + fabs(orig));
    c.xl[lc] += delta;
    double up[3] = {0.0, 0.0, 0.0};
    elec_boundary_row(g, f, ci, c.xl, k, vs, up);
    double* blk = A + ((size_t)v * 9 + 3 + (off + 1)) * AX1_BLK;
    // the flux of species j depends on c_j and phi only: columns other than j and pot get
    // (up - down) == 0 exactly, which keeps the arrow shape of the block
    for (int j = 0; j < 3; ++j) {
      if ((rmask >> j) & 1) continue;   // constrained row
      if (comp == j) blk[2 * j] += (up[j] - down[j]) / delta;
      else if (comp == 3) blk[2 * j + 1] += (up[j] - down[j]) / delta;
    }
  }
}

// elec_volume_row (pnp_device.cuh) with the cell's four quadrature points precomputed and only the rows in `rows`
// (bit r = residual row r of Na, K, Cl, pot) evaluated: same operations in the same order for those rows, so the
// difference quotients below are bit-identical to evaluating all four rows for every perturbation.
__device__ __forceinline__ void elec_volume_rows(const Dev& g, const CellCtx& c, const QP (&qp)[4], double fna, int k,
                                                 double vs, double wm, bool spatial, unsigned rows, double out[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const QP& q = qp[p];
    double uCon[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) s += c.xl[4 * j + m] * q.phi[m];
      uCon[j] = s;
    }
    if (spatial) {
      double gpx = 0, gpy = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) { gpx += c.xl[12 + m] * q.gx[m]; gpy += c.xl[12 + m] * q.gy[m]; }
      if (rows & 8u) {
        const double Apx = (-g.eps_elec) * gpx, Apy = (-g.eps_elec) * gpy;
        double chargeDensity = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) chargeDensity += g.valence[j] * uCon[j];
        const double fPot = -chargeDensity * g.PC;
        double val = (Apx * q.gx[k] + Apy * q.gy[k]) + (-fPot) * q.phi[k];
        out[3] += g.dt * (val * q.factor * vs * g.S_pot);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (!((rows >> j) & 1u)) continue;
        double sx = 0, sy = 0;
#pragma unroll
        for (int m = 0; m < 4; ++m) { sx += c.xl[4 * j + m] * q.gx[m]; sy += c.xl[4 * j + m] * q.gy[m]; }
        const double D = g.D[j];
        const double Acx = D * sx, Acy = D * sy;
        const double bx = -D * g.valence[j] * gpx, by = -D * g.valence[j] * gpy;
        const double f = (j == NA) ? fna : 0.0;
        double val = (Acx * q.gx[k] + Acy * q.gy[k]) - uCon[j] * (bx * q.gx[k] + by * q.gy[k]) + (-f) * q.phi[k];
        out[j] += g.dt * (val * q.factor * vs);
      }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
      if ((rows >> j) & 1u) out[j] += wm * (uCon[j] * q.phi[k] * q.factor * vs);
  }
}

// ------------------------------------------------------------------------------------------
// PDELab NumericalJacobianVolume (formula: stuff/idefault_patch.dat:14-23) for the local operators that run
// without an analytic jacobian_volume -- the reference's DEFAULT (use{Elec,Memb,Time}OperatorJacobianVolume
// = no): per cell, J_ij += (alpha_i(u + delta e_j) - alpha_i(u)) / delta over the 16 local coefficients in
// the order Na0-3, K0-3, Cl0-3, pot0-3, separately for the spatial operator (weight dt inside alpha) and the
// time operator. Row-owner form: one thread per vertex row evaluates its own test function in the <= 4
// cells around it and ADDS to the blocks the analytic kernel left out. Row con_j of alpha does not read the
// other species, so those differences are exactly zero and the arrow shape of the blocks is kept.
__global__ void __launch_bounds__(128) jacobian_volume_fd_kernel(Dev g, const double* __restrict__ u, double* A) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.nv) return;
  const int i = v % g.nvx, j = v / g.nvx;
  const uint8_t rmask = g.dmask[v];
  const double vs = g.do_volume_scaling ? 1. / g.node_vol[v] : 1.0;
  const double fd_eps = 1e-7;
  double* Arow = A + (size_t)v * AX1_ROW;
  for (int b = 0; b < 2; ++b)
    for (int a = 0; a < 2; ++a) {
      const int ci = i - 1 + a, cj = j - 1 + b;
      if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny) continue;
      const int ka = 1 - a, kb = 1 - b, k = ka + 2 * kb;
      const bool memb = (cj == g.jm);
      if (memb ? !(g.fd_mask & 2) : !(g.fd_mask & 5)) continue;
      CellCtx c;
      load_cell(g, u, ci, cj, c);
      if (memb) {
        double down = 0.0;
        memb_volume_row(g, c, k, vs, down);
        for (int m = 0; m < 4; ++m) {
          const int lc = 12 + m;
          const double orig = c.xl[lc];
          const double delta = fd_eps * (1.0 + fabs(orig));
          c.xl[lc] += delta;
          double up = 0.0;
          memb_volume_row(g, c, k, vs, up);
          c.xl[lc] = orig;
          const int s = ((m >> 1) - kb + 1) * 3 + ((m & 1) - ka + 1);
          if (!((rmask >> 3) & 1)) Arow[s * AX1_BLK + 9] += (up - down) / delta;
        }
        continue;
      }
      QP qp[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) make_qp(c, (p >> 1) ? AX1_GP1 : AX1_GP0, (p & 1) ? AX1_GP1 : AX1_GP0, qp[p]);
      const double fna = source_na(g, c);
      for (int part = 0; part < 2; ++part) {     // 0: spatial operator, 1: time operator
        if (!(g.fd_mask & (part ? 4 : 1))) continue;
        const double wm = part ? 1.0 : 0.0;
        const bool spatial = part == 0;
        double down[4] = {0.0, 0.0, 0.0, 0.0};
        elec_volume_rows(g, c, qp, spatial ? fna : 0.0, k, vs, wm, spatial, spatial ? 15u : 7u, down);
        const int ncomp = part ? 3 : 4;           // the time operator lives on the concentrations only
        for (int comp = 0; comp < ncomp; ++comp) {
          // row con_j reads c_j and (spatial part) the potential only, the potential row reads everything: the other
          // difference quotients are exactly zero and are neither evaluated nor added
          const unsigned rows = comp == 3 ? 15u : ((1u << comp) | (spatial ? 8u : 0u));
          for (int m = 0; m < 4; ++m) {
            const int lc = comp * 4 + m;
            const double orig = c.xl[lc];
            const double delta = fd_eps * (1.0 + fabs(orig));
            c.xl[lc] += delta;
            double up[4] = {0.0, 0.0, 0.0, 0.0};
            elec_volume_rows(g, c, qp, spatial ? fna : 0.0, k, vs, wm, spatial, rows, up);
            c.xl[lc] = orig;
            const int s = ((m >> 1) - kb + 1) * 3 + ((m & 1) - ka + 1);
            double* blk = Arow + s * AX1_BLK;
            for (int r = 0; r < 3; ++r) {
              if ((rmask >> r) & 1) continue;
              if (comp == r) blk[2 * r] += (up[r] - down[r]) / delta;
              else if (comp == 3) blk[2 * r + 1] += (up[r] - down[r]) / delta;
            }
            if (spatial && !((rmask >> 3) & 1)) blk[6 + comp] += (up[3] - down[3]) / delta;
          }
        }
      }
    }
}

void launch_jacobian_volume_fd(const Dev& g, const double* u, double* A, cudaStream_t st) {
  if (!g.fd_mask) return;
  const int bs = 128;
  jacobian_volume_fd_kernel<<<(g.nv + bs - 1) / bs, bs, 0, st>>>(g, u, A);
}

void launch_jacobian_boundary_fd(const Dev& g, const double* u, double* A, cudaStream_t st) {
  if (!g.membrane_active) return;
  const int bs = 128;
  const int nthr = 24 * g.nvx;  // 2 rows per column x 4 components x 3 block columns
  jacobian_boundary_fd_kernel<<<(nthr + bs - 1) / bs, bs, 0, st>>>(g, u, A);
}

}  // namespace ax1
