// mori.cu -- the electroneutral ("Mori") operator-split model on the device (BASELINE configs[2]):
// one time step = potential sub-problem, charge-layer update, concentration sub-problem, each a linear problem
// assembled by a row-owner gather kernel and solved with the library's BiCGStab / block-ILU0 on the same 9-slot
// arrow-block matrix as the coupled problem (the other sub-problem's rows are identity), so SpMV, factorisation and
// sweeps are shared. File:line relative to the reference repository:
//   dune/ax1/acme2_cyl/common/acme2_cyl_mori_simulation.hh:562-600   step order, which vector feeds which term
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_potential_operator.hh  alpha_volume / alpha_boundary
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_concentration_operator.hh  alpha_volume / alpha_boundary
//   dune/ax1/acme2_cyl/configurations/mori/mori_poisson_parameters.hh   A :124-158, bGrad :170-186, j (sum z_j j_j)
//   dune/ax1/acme2_cyl/configurations/mori/mori_nernst_planck_parameters.hh  A, b, f, j (channel + capacitive flux)
//   dune/ax1/common/charge_layer_mori_implicit.hh   evaluate :83-250, init / timeStep / acceptTimeStep
//   dune/ax1/common/ax1_linearsubproblem.hh          assemble - solve - update, reduction = max(red, min_defect / d0)
//   dune/ax1/acme2_cyl/common/acme2_cyl_mori_setup.hh:569-575, 726-742, 866-893  wiring (no membrane contributions)
// The opposite side of the membrane enters the interface flux through the stored solution vector, not through the
// unknowns (what the reference's NumericalJacobianBoundary sees); both sub-problems are linear in their unknowns and
// the Jacobians are written analytically.
#include <cmath>
#include <cstring>
#include "ctx.h"
#include "pnp_device.cuh"

namespace ax1 {

struct MoriDev {
  const double* gamma;   // [side 0 cytosol / 1 ES][0 old / 1 new][ion][nx]
  int active;            // charge layer initialised
  int membrane_neumann;  // potential Neumann condition on the membrane interface
  double dmemb;          // membrane thickness [length units]
};

__device__ __forceinline__ double face_centre(const Dev& g, const double* u, int ci, int jrow, int comp) {
  const double* a = u + 4 * (size_t)(ci + g.nvx * jrow);
  return 0.5 * a[comp] + 0.5 * a[4 + comp];
}

// channel + capacitive flux of species j at one quadrature point (before the surface scaling), and its derivative
// with respect to this side's potential (oracle/mori.cpp: interface_flux)
__device__ __forceinline__ double mori_interface_flux(const Dev& g, const MoriDev& md, int ci, bool from_cytosol, int j,
                                                      double uPot, double uConj, double pot2, double con2j,
                                                      double potJumpOld, double& dflux) {
  const int ySign = from_cytosol ? -1 : 1;
  const double n_y = from_cytosol ? 1.0 : -1.0;
  const double C1 = (con_k * g.temperature) / con_e;
  const int valence = g.valence[j];
  double potJump = pot2;
  potJump -= uPot;
  potJump *= ySign;
  const double dJump = -(double)ySign;
  const double conRatio = ySign > 0 ? con2j / uConj : uConj / con2j;
  const double reversal = -log(conRatio) * (C1 / valence);
  double total = 0.0, dtotal = 0.0;
  if (j != CL) {
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int ch = j + 2 * kk;
      double conductance = (g.equil && kk == 1) ? 0.0 : g.geff[(size_t)ch * g.nx + ci];
      if (!g.membrane_active) conductance = 0.0;
      double C = (10 * conductance) / (con_e * valence * con_mol);
      if (g.equil) C *= g.dummy_value;
      C *= g.time_scale / g.length_scale;
      double diffTerm = C * reversal;
      if (fabs(conRatio - 1) < 1e-8 || isnan(conRatio)) diffTerm = 0.0;
      const double driftTerm = C * C1 * potJump;
      double flux = driftTerm;
      flux -= diffTerm;
      total += flux;
      dtotal += C * C1 * dJump;
    }
  }
  double y = total * n_y, dy = dtotal * n_y;
  if (md.active) {
    const double real_dt = g.dt * g.time_scale;
    const double C_m = con_eps0 * g.eps_memb[ci] / (md.dmemb * g.length_scale);
    const double mV = C1 * 1.0e3;
    const double membPotOld = 1e-3 * (potJumpOld * mV);
    const double membPotNew = 1e-3 * (potJump * mV);
    const size_t side = from_cytosol ? 0 : 1;
    const double go = md.gamma[((side * 2 + 0) * 3 + j) * g.nx + ci];
    const double gn = md.gamma[((side * 2 + 1) * 3 + j) * g.nx + ci];
    const double sigma_old = go * C_m * membPotOld;
    const double sigma_new = gn * C_m * membPotNew;
    const double scs = (g.time_scale / g.length_scale) / (con_e * con_mol);
    const double moriFlux = (scs / valence) * (sigma_new - sigma_old) / real_dt;
    const double dmori = (scs / valence) * (gn * C_m * (1e-3 * (dJump * mV))) / real_dt;
    y += moriFlux * n_y;
    dy += dmori * n_y;
  }
  dflux = dy;
  return y;
}

// PHASE 0: potential rows, PHASE 1: concentration rows. One thread per vertex row: residual entry(ies) and the 9 slots
// of the Jacobian row, written once in a fixed order.
//   u     current iterate (unknowns of the phase are read from it)
//   ulast solution of the last iteration (concentrations of the potential problem; drift / membrane terms of phase 1)
//   uold  previous time step (mass term, old potential jump, this side's concentration in the phase-1 membrane flux)
template <int PHASE>
__global__ void __launch_bounds__(128) mori_assemble_kernel(Dev g, MoriDev md, const double* __restrict__ u,
                                                            const double* __restrict__ ulast,
                                                            const double* __restrict__ uold, double* __restrict__ r,
                                                            double* __restrict__ A) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.nv) return;
  const int i = v % g.nvx, j = v / g.nvx;
  constexpr int NR = PHASE == 0 ? 1 : 3;
  double out[NR], J[NR][9];
#pragma unroll
  for (int a = 0; a < NR; ++a) {
    out[a] = 0.0;
#pragma unroll
    for (int s = 0; s < 9; ++s) J[a][s] = 0.0;
  }
  const double vs = g.do_volume_scaling ? 1. / g.node_vol[v] : 1.0;
  const double cond_unit = g.length_scale * con_e * con_mol;
  const double jscale = g.PC * g.time_scale / (g.length_scale * g.length_scale);
#pragma unroll
  for (int b = 0; b < 2; ++b)
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ci = i - 1 + a, cj = j - 1 + b;
      if (ci < 0 || ci >= g.nx || cj < 0 || cj >= g.ny) continue;
      if (cj == g.jm) continue;                        // no membrane contributions
      const int k = (1 - a) + 2 * (1 - b);             // my local vertex in that cell
      CellCtx c, cl;
      load_cell(g, u, ci, cj, c);
      load_cell(g, ulast, ci, cj, cl);
      double rl[NR], ml[NR][4];
#pragma unroll
      for (int q = 0; q < NR; ++q) { rl[q] = 0.0; ml[q][0] = ml[q][1] = ml[q][2] = ml[q][3] = 0.0; }
      CellCtx co;      // previous time step: mass term (phase 1), this side's concentration in the interface flux
      if (PHASE == 1 || cj == g.jm - 1 || cj == g.jm + 1) load_cell(g, uold, ci, cj, co);
      const double fna = PHASE == 1 ? source_na(g, c) : 0.0;
#pragma unroll
      for (int qa = 0; qa < 2; ++qa)
#pragma unroll
        for (int qb = 0; qb < 2; ++qb) {
          QP q;
          make_qp(c, qa ? AX1_GP1 : AX1_GP0, qb ? AX1_GP1 : AX1_GP0, q);
          if (PHASE == 0) {
            double cond = 0.0, bgx = 0.0, bgy = 0.0;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              double val = 0, sx = 0, sy = 0;
#pragma unroll
              for (int m = 0; m < 4; ++m) { val += cl.xl[4 * s + m] * q.phi[m]; sx += cl.xl[4 * s + m] * q.gx[m]; sy += cl.xl[4 * s + m] * q.gy[m]; }
              cond += cond_unit * g.valence[s] * g.valence[s] * g.D[s] * val;
              double tx = g.D[s] * sx, ty = g.D[s] * sy;
              tx *= cond_unit * g.valence[s]; ty *= cond_unit * g.valence[s];
              bgx += tx; bgy += ty;
            }
            double gpx = 0, gpy = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) { gpx += c.xl[12 + m] * q.gx[m]; gpy += c.xl[12 + m] * q.gy[m]; }
            const double Ax = cond * gpx, Ay = cond * gpy;
            rl[0] += ((Ax * q.gx[k] + Ay * q.gy[k]) + (bgx * q.gx[k] + bgy * q.gy[k])) * q.factor * vs * g.S_pot;
#pragma unroll
            for (int m = 0; m < 4; ++m)
              ml[0][m] += ((cond * q.gx[m]) * q.gx[k] + (cond * q.gy[m]) * q.gy[k]) * q.factor * vs * g.S_pot;
          } else {
            double gpx = 0, gpy = 0;
#pragma unroll
            for (int m = 0; m < 4; ++m) { gpx += cl.xl[12 + m] * q.gx[m]; gpy += cl.xl[12 + m] * q.gy[m]; }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              double uC = 0, uCl = 0, uCo = 0, sx = 0, sy = 0;
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                uC += c.xl[4 * s + m] * q.phi[m]; uCl += cl.xl[4 * s + m] * q.phi[m]; uCo += co.xl[4 * s + m] * q.phi[m];
                sx += c.xl[4 * s + m] * q.gx[m]; sy += c.xl[4 * s + m] * q.gy[m];
              }
              const double D = g.D[s];
              const double Acx = D * sx, Acy = D * sy;
              const double bx = -D * g.valence[s] * gpx, by = -D * g.valence[s] * gpy;
              const double f = (s == NA) ? fna : 0.0;
              const double spatial = (Acx * q.gx[k] + Acy * q.gy[k]) - uCl * (bx * q.gx[k] + by * q.gy[k]) + (0.0 * uC - f) * q.phi[k];
              rl[PHASE == 1 ? s : 0] += g.dt * (spatial * q.factor * vs);
              rl[PHASE == 1 ? s : 0] += (uC * q.phi[k] * q.factor * vs);
              rl[PHASE == 1 ? s : 0] += -(uCo * q.phi[k] * q.factor * vs);
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                ml[PHASE == 1 ? s : 0][m] += g.dt * (((D * q.gx[m]) * q.gx[k] + (D * q.gy[m]) * q.gy[k]) * q.factor * vs);
                ml[PHASE == 1 ? s : 0][m] += (q.phi[m] * q.phi[k] * q.factor * vs);
              }
            }
          }
        }
      // membrane-interface face of this cell, if my vertex lies on it
      if (((cj == g.jm - 1 && b == 0) || (cj == g.jm + 1 && b == 1)) && (PHASE == 1 || md.membrane_neumann)) {
        const bool from_cytosol = (cj == g.jm - 1);
        const double eta = from_cytosol ? 1.0 : 0.0;
        const double y_this = from_cytosol ? g.y[g.jm] : g.y[g.jm + 1];
        const double y_es = g.y[g.jm + 1];
        const double my_surface = (con_pi * (y_this + y_this)) * c.hx;
        const double ext_surface = (con_pi * (y_es + y_es)) * c.hx;
        const double surfaceScale = ext_surface / my_surface;
        const double ffac = 0.5 * ((2 * con_pi * y_this) * c.hx);
        const int jthis = from_cytosol ? g.jm : g.jm + 1, jopp = from_cytosol ? g.jm + 1 : g.jm;
        const int ySign = from_cytosol ? -1 : 1;
        const double* upn = PHASE == 0 ? u : ulast;       // vector holding the "new" potential of this phase
        const double pot2 = face_centre(g, upn, ci, jopp, 3);
        const double potJumpOld = (face_centre(g, uold, ci, jopp, 3) - face_centre(g, uold, ci, jthis, 3)) * ySign;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          double phi[4], dxi[4], deta[4];
          q1(s ? AX1_GP1 : AX1_GP0, eta, phi, dxi, deta);
          double uPot = 0.0;
#pragma unroll
          for (int m = 0; m < 4; ++m) uPot += (PHASE == 0 ? c.xl[12 + m] : cl.xl[12 + m]) * phi[m];
          double jPot = 0.0, djPot = 0.0;
#pragma unroll
          for (int sp = 0; sp < 3; ++sp) {
            double uc = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) uc += co.xl[4 * sp + m] * phi[m];   // getSolutionConOld in both operators
            double d;
            double jc = mori_interface_flux(g, md, ci, from_cytosol, sp, uPot, uc, pot2, face_centre(g, ulast, ci, jopp, sp),
                                            potJumpOld, d);
            jc *= surfaceScale; d *= surfaceScale;
            if (PHASE == 0) { jPot += g.valence[sp] * jc; djPot += g.valence[sp] * d; }
            else rl[PHASE == 1 ? sp : 0] += g.dt * (jc * phi[k] * ffac * vs);
          }
          if (PHASE == 0) {
            jPot *= jscale; djPot *= jscale;
            rl[0] += jPot * phi[k] * ffac * vs * g.S_pot;
#pragma unroll
            for (int m = 0; m < 4; ++m) ml[0][m] += (djPot * phi[m]) * phi[k] * ffac * vs * g.S_pot;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NR; ++q) {
        out[q] += rl[q];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int di = ci + (m & 1) - i, dj = cj + (m >> 1) - j;
          J[q][(dj + 1) * 3 + (di + 1)] += ml[q][m];
        }
      }
    }
  // constraints + write-out (arrow block: [d0 c0 | d1 c1 | d2 c2 | p0 p1 p2 p3])
  const uint8_t dm = g.dmask[v];
  double rv[4] = {0.0, 0.0, 0.0, 0.0};
  double* arow = A + (size_t)v * AX1_ROW;
#pragma unroll
  for (int s = 0; s < 9; ++s) {
    double blk[AX1_BLK];
#pragma unroll
    for (int e = 0; e < AX1_BLK; ++e) blk[e] = 0.0;
    if (PHASE == 0) {
      const bool con = (dm >> 3) & 1;
      blk[9] = con ? (s == 4 ? 1.0 : 0.0) : J[0][s];
      if (s == 4) { blk[0] = 1.0; blk[2] = 1.0; blk[4] = 1.0; }
    } else {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const bool con = (dm >> q) & 1;
        blk[2 * q] = con ? (s == 4 ? 1.0 : 0.0) : J[PHASE == 1 ? q : 0][s];
      }
      if (s == 4) blk[9] = 1.0;
    }
#pragma unroll
    for (int e = 0; e < AX1_BLK; e += 2) *reinterpret_cast<double2*>(arow + s * AX1_BLK + e) = make_double2(blk[e], blk[e + 1]);
  }
  if (PHASE == 0) rv[3] = ((dm >> 3) & 1) ? 0.0 : out[0];
  else {
#pragma unroll
    for (int q = 0; q < 3; ++q) rv[q] = ((dm >> q) & 1) ? 0.0 : out[PHASE == 1 ? q : 0];
  }
  double2* dst = reinterpret_cast<double2*>(r + 4 * (size_t)v);
  dst[0] = make_double2(rv[0], rv[1]);
  dst[1] = make_double2(rv[2], rv[3]);
}

// charge layer: mode 0 init (gamma = z^2 c / sum z^2 c on both sides), 1 time step (relaxation with tau = 1 ns),
// 2 accept (old = new)   (charge_layer_mori_implicit.hh init / timeStep / acceptTimeStep)
__global__ void mori_gamma_kernel(Dev g, double* gamma, const double* __restrict__ u, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nx) return;
  auto G = [&](int side, int nw, int j) -> double& { return gamma[(((size_t)side * 2 + nw) * 3 + j) * g.nx + i]; };
  if (mode == 2) {
#pragma unroll
    for (int side = 0; side < 2; ++side)
#pragma unroll
      for (int j = 0; j < 3; ++j) G(side, 0, j) = G(side, 1, j);
    return;
  }
  const double tau = 1e-9, dt_ = g.dt * g.time_scale;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    double con[3], sum = 0.0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      con[j] = face_centre(g, u, i, side == 0 ? g.jm : g.jm + 1, j);
      sum += (double)(g.valence[j] * g.valence[j]) * con[j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double snake = (double)(g.valence[j] * g.valence[j]) * con[j] / sum;
      if (mode == 0) { G(side, 0, j) = snake; G(side, 1, j) = snake; }
      else G(side, 1, j) = (tau * G(side, 0, j) + dt_ * snake) / (tau + dt_);
    }
  }
}

__global__ void __launch_bounds__(256) mori_axpy_kernel(size_t n, double* u, const double* __restrict__ z) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) u[k] -= z[k];
}

#define AX1_MORI_CHECK(c)                                                    \
  do {                                                                       \
    (c)->launches++;                                                         \
    cudaError_t e_ = cudaGetLastError();                                     \
    if (e_ != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e_)); return AX1GPU_ECUDA; } \
  } while (0)

static MoriDev mori_dev(const ax1gpu_ctx* c) {
  MoriDev md;
  md.gamma = c->d_mori_gamma;
  md.active = (c->mori_initialized || c->mori_partially_initialized) ? 1 : 0;
  md.membrane_neumann = c->prm.mori_membrane_neumann;
  md.dmemb = c->hy[c->g.jm + 1] - c->hy[c->g.jm];
  return md;
}

static int mori_ensure(ax1gpu_ctx* c) {
  if (c->nranks > 1) { set_error("the electroneutral (Mori) model runs on a single rank"); return AX1GPU_EINVAL; }
  if (!c->d_mori_gamma) {
    AX1_CUDA(cudaMalloc((void**)&c->d_mori_gamma, 12 * (size_t)c->g.nx * sizeof(double)));
    AX1_CUDA(cudaMemsetAsync(c->d_mori_gamma, 0, 12 * (size_t)c->g.nx * sizeof(double), c->stream));
  }
  return 0;
}

int drv_mori_update_state(ax1gpu_ctx* c, const double* d_u) {   // ChargeLayerMoriImplicit::updateState
  int rc = mori_ensure(c);
  if (rc) return rc;
  if (!c->mori_initialized && c->mori_partially_initialized) return 0;
  const ax1gpu_params& p = c->prm;
  drv_refresh_dev(c);
  const bool do_init = !c->mori_initialized;
  if (do_init && p.do_equilibration && !(std::fabs(c->time - p.t_equilibrium) < 1e-6 || c->time > p.t_equilibrium)) return 0;
  mori_gamma_kernel<<<(c->g.nx + 127) / 128, 128, 0, c->stream>>>(c->g, c->d_mori_gamma, d_u, do_init ? 0 : 1);
  AX1_MORI_CHECK(c);
  if (do_init) c->mori_partially_initialized = true;
  return 0;
}

int drv_mori_accept_state(ax1gpu_ctx* c) {
  if (!c->d_mori_gamma) return 0;
  if (c->mori_partially_initialized) { c->mori_initialized = true; c->mori_partially_initialized = false; }
  mori_gamma_kernel<<<(c->g.nx + 127) / 128, 128, 0, c->stream>>>(c->g, c->d_mori_gamma, c->d_u, 2);
  AX1_MORI_CHECK(c);
  return 0;
}

int drv_mori_assemble(ax1gpu_ctx* c, int which, const double* d_u, const double* d_ulast) {
  int rc = mori_ensure(c);
  if (rc) return rc;
  drv_refresh_dev(c);
  const MoriDev md = mori_dev(c);
  const int bs = 128, nb = (c->g.nv + bs - 1) / bs;
  if (which == 0) mori_assemble_kernel<0><<<nb, bs, 0, c->stream>>>(c->g, md, d_u, d_ulast, c->d_uold, c->d_r, c->d_A);
  else mori_assemble_kernel<1><<<nb, bs, 0, c->stream>>>(c->g, md, d_u, d_ulast, c->d_uold, c->d_r, c->d_A);
  AX1_MORI_CHECK(c);
  c->factored = false;
  return 0;
}

// one operator-split time step on the device iterate d_u (= uold = ulast on entry). d_uprev keeps ulast.
int drv_mori_step(ax1gpu_ctx* c, const ax1gpu_newton_opts* o, ax1gpu_mori_result* res) {
  int rc = mori_ensure(c);
  if (rc) return rc;
  std::memset(res, 0, sizeof(*res));
  const size_t n = 4 * (size_t)c->g.nv, bytes = n * sizeof(double);
  if (!c->mori_initialized && !c->mori_partially_initialized) {   // [init] block of the simulation
    if ((rc = drv_mori_update_state(c, c->d_u))) return rc;
    if ((rc = drv_mori_accept_state(c))) return rc;
  }
  AX1_CUDA(cudaMemcpyAsync(c->d_uprev, c->d_u, bytes, cudaMemcpyDeviceToDevice, c->stream));   // ulast
  auto solve = [&](double* first_defect, double* defect, int* its) -> int {
    double d0 = 0.0;
    int r_ = drv_norm(c, c->d_r, &d0);
    if (r_) return r_;
    *first_defect = d0;
    if (!std::isfinite(d0)) { set_error("Initial residual is NaN!"); return AX1GPU_ESTATE; }
    ax1gpu_lin_result lr{};
    lr.converged = 1;
    if (d0 > 0.0) {
      const double red = std::max(o->reduction, o->abs_limit / d0);
      if ((r_ = drv_factor(c, o->slab_w, o->ilu_fill, 0))) return r_;
      if ((r_ = drv_solve(c, c->d_r, c->d_z, red, o->lin_maxit, &lr))) return r_;
      if (!lr.converged) { set_error("linear sub-problem solver did not converge"); return AX1GPU_ESTATE; }
      mori_axpy_kernel<<<c->nred_blocks, 256, 0, c->stream>>>(n, c->d_u, c->d_z);
      AX1_MORI_CHECK(c);
    }
    *its = lr.iterations;
    *defect = d0 * lr.reduction;
    return 0;
  };
  if ((rc = drv_mori_assemble(c, 0, c->d_u, c->d_uprev))) return rc;
  if ((rc = solve(&res->pot_first_defect, &res->pot_defect, &res->pot_iterations))) return rc;
  if ((rc = drv_mori_update_state(c, c->d_uprev))) return rc;
  if ((rc = drv_mori_assemble(c, 1, c->d_u, c->d_uprev))) return rc;
  if ((rc = solve(&res->con_first_defect, &res->con_defect, &res->con_iterations))) return rc;
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  res->status = 0;
  return 0;
}

}  // namespace ax1
