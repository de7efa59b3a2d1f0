// assembly.cu -- residual, mass (r0) and Jacobian assembly kernels + HH channel kernels.
//
// Row-owner ("gather") assembly: thread(s) own one vertex row and visit the <=4 cells around it,
// so every residual entry / Jacobian block is written exactly once, in a fixed order: atomic-free
// and bitwise deterministic. Replaces the PDELab cell loop + scatter and the
// BCRSMatrix backend (acme2_cyl_setup.hh:696-708).
#include "kernels.h"
#include "pnp_device.cuh"

namespace ax1 {

// ------------------------------------------------------------------------------------------
// residual: one thread per vertex. mode 0: full residual r = r0 + M(u) + dt*A(u), constrained
// rows zeroed.  mode 1: r0 = -M(uold) (const part of implicit Euler, no constraints applied).
__global__ void __launch_bounds__(128) residual_kernel(Dev g, const double* __restrict__ u,
                                                       const double* __restrict__ r0, double* __restrict__ r,
                                                       int mode) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.nv) return;
  const int i = v % g.nvx, j = v / g.nvx;
  double out[4] = {0.0, 0.0, 0.0, 0.0};
  const double vs = g.do_volume_scaling ? 1. / g.node_vol[v] : 1.0;
  // cells in index order (x fastest): (i-1,j-1), (i,j-1), (i-1,j), (i,j)
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ci = i - 1 + a, cj = j - 1 + b;
      if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny) continue;
      const int k = (1 - a) + 2 * (1 - b);  // our local vertex number in that cell
      CellCtx c;
      load_cell(g, u, ci, cj, c);
      if (cj == g.jm) {
        if (mode == 0) memb_volume_row(g, c, k, vs, out[3]);
      } else {
        elec_volume_row(g, c, k, vs, mode == 0 ? 1.0 : -1.0, mode == 0, out);
        if (mode == 0 && (g.membrane_active || g.fstim_on) && ((cj == g.jm - 1 && b == 0) || (cj == g.jm + 1 && b == 1))) {
          // our vertex lies on the membrane interface face of this cell
          FaceCtx f;
          make_face(g, u, c, f);
          elec_boundary_row(g, f, ci, c.xl, k, vs, out);
        }
      }
    }
  if (mode == 0) {
    double q0, q1_, q2, q3;
    ld4(r0 + 4 * (size_t)v, q0, q1_, q2, q3);
    out[0] += q0; out[1] += q1_; out[2] += q2; out[3] += q3;
    const uint8_t m = g.dmask[v];
    if (m & 1) out[0] = 0.0;
    if (m & 2) out[1] = 0.0;
    if (m & 4) out[2] = 0.0;
    if (m & 8) out[3] = 0.0;
  }
  double2* dst = reinterpret_cast<double2*>(r + 4 * (size_t)v);
  dst[0] = make_double2(out[0], out[1]);
  dst[1] = make_double2(out[2], out[3]);
}

// ------------------------------------------------------------------------------------------
// Jacobian: 4 lanes per vertex row, lane ri owns scalar row ri of the 9 blocks of the block row.
// Analytic volume blocks (the reference's use*OperatorJacobianVolume = yes mode) + mass blocks.
// Constrained rows -> unit row; constrained columns are kept (PDELab's default non-symmetric etadd).
//
// Per cell and Gauss point the entries factor as (test k, trial j, G_jk = grad psi_j . grad psi_k):
//   con_i <- con_i : dt W [ D G_jk - psi_j (b . grad psi_k) ] + W psi_j psi_k,  b = -D z grad(phi)
//   con_i <- pot   : dt W c_i D z G_jk
//   pot   <- con_i : dt W S PC z_i psi_j psi_k
//   pot   <- pot   : dt W S (-eps) G_jk
// (acme2_cyl_operator_fully_coupled.hh:516-561, acme2_cyl_toperator.hh:251, convectiondiffusionfem.hh:317-318)
// so each lane needs only the potential coefficients and those of its own species.
__global__ void __launch_bounds__(128, 4) jacobian_kernel(Dev g, const double* __restrict__ u, double* __restrict__ A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = t >> 2, ri = t & 3;
  if (v >= g.nv) return;
  const int i = v % g.nvx, j = v / g.nvx;
  // every lane accumulates two scalar columns per slot:  a0 += alpha G + beta psi_j,  a1 += gamma G
  //   con row i: a0 -> column i,  a1 -> column pot;   alpha = dt W D, beta = W psi_k - dt W (b.grad psi_k), gamma = dt W c_i D z
  //   pot row  : a0 -> columns con (times z_kk), a1 -> column pot; alpha = 0, beta = dt W S PC psi_k, gamma = dt W S (-eps)
  double a0[9], a1[9];
#pragma unroll
  for (int s = 0; s < 9; ++s) { a0[s] = 0.0; a1[s] = 0.0; }
  const double vs = g.do_volume_scaling ? 1. / g.node_vol[v] : 1.0;
  const bool potrow = (ri == 3);
  const double Di = potrow ? 0.0 : g.D[ri];
  const double Dz = potrow ? 0.0 : Di * g.valence[ri];
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ci = i - 1 + a, cj = j - 1 + b;
      if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny) continue;
      const int k = (1 - a) + 2 * (1 - b);
      const int ka = 1 - a, kb = 1 - b;
      const bool memb = (cj == g.jm);
      if (memb && !potrow) continue;  // membrane cells carry the potential only
      // operators whose volume Jacobian is taken by forward differences are left to jacobian_volume_fd_kernel
      const bool opA = !(g.fd_mask & (memb ? 2 : 1));   // spatial operator of this cell: analytic?
      const bool timeA = !(g.fd_mask & 4);
      if (potrow ? !opA : (!opA && !timeA)) continue;
      const double x0 = g.x[ci], x1 = g.x[ci + 1], y0 = g.y[cj], y1 = g.y[cj + 1];
      const double hx = x1 - x0, hy = y1 - y0, ihx = 1.0 / hx, ihy = 1.0 / hy;
      const int v0 = ci + g.nvx * cj;
      const int vv[4] = {v0, v0 + 1, v0 + g.nvx, v0 + g.nvx + 1};
      double pot[4], own[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        pot[m] = potrow ? 0.0 : u[4 * (size_t)vv[m] + 3];
        own[m] = potrow ? 0.0 : u[4 * (size_t)vv[m] + ri];
      }
      const double eps = memb ? g.eps_memb[ci] : g.eps_elec;
      const double cS = g.dt * g.S_pot;
#pragma unroll
      for (int qa = 0; qa < 2; ++qa)
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) {
          const double xi = qa ? AX1_GP1 : AX1_GP0, eta = qb ? AX1_GP1 : AX1_GP0;
          double phi[4], dxi[4], deta[4], gx[4], gy[4];
          q1(xi, eta, phi, dxi, deta);
#pragma unroll
          for (int m = 0; m < 4; ++m) { gx[m] = ihx * dxi[m]; gy[m] = ihy * deta[m]; }
          const double yq = y0 + eta * hy;
          const double W = (0.25 * ((2 * con_pi * yq) * (hx * hy))) * vs;
          double alpha, beta, gamma;
          if (potrow) {
            alpha = 0.0;
            beta = memb ? 0.0 : cS * W * g.PC * phi[k];
            gamma = cS * W * (-eps);
          } else {
            double gpx = 0, gpy = 0, uk = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) { gpx += pot[m] * gx[m]; gpy += pot[m] * gy[m]; uk += own[m] * phi[m]; }
            const double Wdt = g.dt * W;
            const double bk = -Dz * (gpx * gx[k] + gpy * gy[k]);  // b . grad psi_k
            alpha = opA ? Wdt * Di : 0.0;
            beta = (timeA ? W * phi[k] : 0.0) - (opA ? Wdt * bk : 0.0);
            gamma = opA ? Wdt * uk * Dz : 0.0;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int s = ((jj >> 1) - kb + 1) * 3 + ((jj & 1) - ka + 1);  // compile-time after unrolling
            const double G = gx[jj] * gx[k] + gy[jj] * gy[k];
            a0[s] += alpha * G + beta * phi[jj];
            a1[s] += gamma * G;
          }
        }
    }
  const bool row_con = (g.dmask[v] >> ri) & 1;
  const double z0 = g.valence[0], z1 = g.valence[1], z2 = g.valence[2];
  double* Arow = A + (size_t)v * AX1_ROW;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    const double unit = (row_con && s == 4) ? 1.0 : 0.0;
    if (potrow) {
      double o0 = z0 * a0[s], o1 = z1 * a0[s], o2 = z2 * a0[s], o3 = a1[s];
      if (row_con) { o0 = 0.0; o1 = 0.0; o2 = 0.0; o3 = unit; }
      double2* dst = reinterpret_cast<double2*>(Arow + s * AX1_BLK + 6);
      dst[0] = make_double2(o0, o1);
      dst[1] = make_double2(o2, o3);
    } else {
      double d = a0[s], c = a1[s];
      if (row_con) { d = unit; c = 0.0; }
      *reinterpret_cast<double2*>(Arow + s * AX1_BLK + 2 * ri) = make_double2(d, c);
    }
  }
}

// ------------------------------------------------------------------------------------------
// HH channels: one thread per membrane column (= ES-side membrane edge).
__device__ __forceinline__ double v_trap(double x, double y) {  // channel.hh:185-194
  if (fabs(x / y) < 1e-8) return y * (1 - x / y / 2);
  return x / (exp(x / y) - 1);
}
__device__ __forceinline__ double alpha_gate(int gt, double v_) {  // channel.hh:606-618, 667-676
  if (gt == 0) return 1.0 * (0.1 * v_trap((25 - v_), 10));
  if (gt == 1) return 1.0 * (0.07 * exp(-v_ / 20));
  return 1.0 * (0.01 * v_trap((10 - v_), 10));
}
__device__ __forceinline__ double beta_gate(int gt, double v_) {  // channel.hh:624-642, 679-687
  if (gt == 0) return 1.0 * (4 * exp(-v_ / 18));
  if (gt == 1) return 1.0 * (1. / (exp((30 - v_) / 10) + 1));
  return 1.0 * (0.125 * exp(-v_ / 80));
}

__device__ __forceinline__ double memb_pot_mV(const Dev& g, const double* u, int i) {  // physics.hh:562-603
  const double in = 0.5 * u[4 * (size_t)(i + g.nvx * g.jm) + 3] + 0.5 * u[4 * (size_t)(i + 1 + g.nvx * g.jm) + 3];
  const double out = 0.5 * u[4 * (size_t)(i + g.nvx * (g.jm + 1)) + 3] + 0.5 * u[4 * (size_t)(i + 1 + g.nvx * (g.jm + 1)) + 3];
  double membPot = in;
  membPot -= out;
  return membPot * con_k * g.temperature / con_e * 1.0e3;
}

__device__ __forceinline__ void write_geff(const ChannelDev& ch, int nx, int i) {
  // getNewEffConductance: channel.hh:356-366 (voltage gated), :497-501 (leak)
  ch.geff[0 * (size_t)nx + i] = ch.gbar[0][i] * ch.ratio[0][i];
  ch.geff[1 * (size_t)nx + i] = ch.gbar[0][i] * ch.ratio[1][i];
  double gn = ch.gbar[1][i];
  gn *= pow(ch.gate_new[0][i], 3.0);
  gn *= pow(ch.gate_new[1][i], 1.0);
  ch.geff[2 * (size_t)nx + i] = gn;
  double gk = ch.gbar[2][i];
  gk *= pow(ch.gate_new[2][i], 4.0);
  ch.geff[3 * (size_t)nx + i] = gk;
}

// mode 0: initChannels (steady state at v - v_rest = 0, leak re-balancing, channelset.hh:186-311)
// mode 1: backward-Euler gating step from the accepted state (channel.hh:394-402)
// mode 2: only refresh geff from the current state
__global__ void channel_kernel(Dev g, ChannelDev ch, const double* u, int mode, double dt_ms, int automatic,
                               double v_fixed, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nx) return;
  if (mode == 0) {
    const double v = automatic ? memb_pot_mV(g, u, i) : v_fixed;
    // v_rest of the shared channel objects ends up as the value of the LAST edge visited
    if (i == g.nx - 1) *ch.v_rest = v;
    for (int gt = 0; gt < 3; ++gt) {
      const double a = alpha_gate(gt, v - v), b = beta_gate(gt, v - v);
      ch.gate[gt][i] = a / (a + b);
      ch.gate_new[gt][i] = ch.gate[gt][i];
    }
    // recalculateLeakConductances
    double gnav = ch.gbar[1][i];
    gnav *= pow(ch.gate_new[0][i], 3.0);
    gnav *= pow(ch.gate_new[1][i], 1.0);
    double gkv = ch.gbar[2][i];
    gkv *= pow(ch.gate_new[2][i], 4.0);
    const double total = ch.gbar[0][i];
    if (total > 0.0) {
      for (int k = 0; k < 2; ++k) {
        const double sumThis = (k == 0) ? gnav : gkv;
        const double sumOther = (k == 0) ? gkv : gnav;
        const double ratio = ch.ratio[k][i];
        double leak = (ratio / (1. - ratio)) * (total + (0.0 + sumOther)) - (0.0 + sumThis);
        leak /= 1. + (ratio / (1. - ratio));
        const double newRatio = leak / total;
        if (newRatio < 0.0 || newRatio > 1.0) atomicExch(err, 1);
        ch.ratio[k][i] = newRatio;
      }
    }
  } else if (mode == 1) {
    const double v_ = memb_pot_mV(g, u, i) - *ch.v_rest;
    for (int gt = 0; gt < 3; ++gt) {
      const double a = alpha_gate(gt, v_), b = beta_gate(gt, v_);
      ch.gate_new[gt][i] = (dt_ms * a + ch.gate[gt][i]) / (1 + dt_ms * (a + b));
    }
  }
  write_geff(ch, g.nx, i);
}

__global__ void channel_accept_kernel(int nx, ChannelDev ch) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nx) return;
  for (int gt = 0; gt < 3; ++gt) ch.gate[gt][i] = ch.gate_new[gt][i];
}

// The five per-ion output terms of MembraneFluxGridFunctionImplicit::evaluate (FluxTerms: total, diffusion term,
// drift term, leak channels, voltage-gated channels; membranefluxgridfunction_implicit.hh:42, 83-361) per membrane
// column, seen from the cytosol side at the face centres of the two interface edges, with the NEW channel state:
// the data behind memb_flux*.dat / the HDF5 datasets of acme2_cyl_output.hh:1215-1410. out[(term * 3 + ion) * nx + i].
__global__ void memb_flux_terms_kernel(Dev g, const double* __restrict__ u, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nx) return;
  double t[15];
#pragma unroll
  for (int k = 0; k < 15; ++k) t[k] = 0.0;
  if (g.membrane_active) {
    const double* ua = u + 4 * (size_t)(i + g.nvx * g.jm);
    const double* uc = u + 4 * (size_t)(i + g.nvx * (g.jm + 1));
    double con1[3], con2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { con1[j] = 0.5 * ua[j] + 0.5 * ua[4 + j]; con2[j] = 0.5 * uc[j] + 0.5 * uc[4 + j]; }
    const double pot1 = 0.5 * ua[3] + 0.5 * ua[7], pot2 = 0.5 * uc[3] + 0.5 * uc[7];
    double potJump = pot2;
    potJump -= pot1;
    potJump *= -1;                       // ySign: this side is the cytosol
    const double C1 = (con_k * g.temperature) / con_e;
#pragma unroll
    for (int j = 0; j < 2; ++j) {        // Na, K; Cl has no channel
      const double conRatio = con1[j] / con2[j];
      const int valence = g.valence[j];
      const double reversal = -log(conRatio) * (C1 / valence);
      double total = 0.0, diff = 0.0, drift = 0.0, leakf = 0.0, vg = 0.0;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int ch = j + 2 * kk;       // 0 Na_leak, 1 K_leak, 2 Nav, 3 Kv
        const double conductance = (g.equil && kk == 1) ? 0.0 : g.geff[(size_t)ch * g.nx + i];
        double C = (10 * conductance) / (con_e * valence * con_mol);
        if (g.equil) C *= g.dummy_value;
        C *= g.time_scale / g.length_scale;
        double diffTerm = C * reversal;
        if (fabs(conRatio - 1) < 1e-8 || isnan(conRatio)) diffTerm = 0.0;
        const double driftTerm = C * C1 * potJump;
        double flux = driftTerm;
        flux -= diffTerm;
        diff += diffTerm; drift += driftTerm; total += flux;
        if (kk == 0) leakf += flux; else vg += flux;
      }
      t[0 + j] = total; t[3 + j] = diff; t[6 + j] = drift; t[9 + j] = leakf; t[12 + j] = vg;
    }
  }
#pragma unroll
  for (int k = 0; k < 15; ++k) out[(size_t)k * g.nx + i] = t[k];
}

__global__ void memb_pot_kernel(Dev g, const double* u, double* vm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nx) return;
  vm[i] = memb_pot_mV(g, u, i);
}

// ------------------------------------------------------------------------------------------
void launch_residual(const Dev& g, const double* u, const double* r0, double* r, int mode, cudaStream_t st) {
  const int bs = 128;
  residual_kernel<<<(g.nv + bs - 1) / bs, bs, 0, st>>>(g, u, r0, r, mode);
}
void launch_jacobian_volume(const Dev& g, const double* u, double* A, cudaStream_t st) {
  const int bs = 128;
  const long long nt = 4LL * g.nv;
  jacobian_kernel<<<(unsigned)((nt + bs - 1) / bs), bs, 0, st>>>(g, u, A);
}
void launch_channels(const Dev& g, const ChannelDev& ch, const double* u, int mode, double dt_ms, int automatic,
                     double v_fixed, int* err, cudaStream_t st) {
  const int bs = 128;
  channel_kernel<<<(g.nx + bs - 1) / bs, bs, 0, st>>>(g, ch, u, mode, dt_ms, automatic, v_fixed, err);
}
void launch_channel_accept(int nx, const ChannelDev& ch, cudaStream_t st) {
  const int bs = 128;
  channel_accept_kernel<<<(nx + bs - 1) / bs, bs, 0, st>>>(nx, ch);
}
void launch_memb_flux_terms(const Dev& g, const double* u, double* out, cudaStream_t st) {
  const int bs = 128;
  memb_flux_terms_kernel<<<(g.nx + bs - 1) / bs, bs, 0, st>>>(g, u, out);
}
void launch_memb_pot(const Dev& g, const double* u, double* vm, cudaStream_t st) {
  const int bs = 128;
  memb_pot_kernel<<<(g.nx + bs - 1) / bs, bs, 0, st>>>(g, u, vm);
}

}  // namespace ax1
