// pnp_device.cuh -- per-cell device functions of the PNP element kernels, written for the
// row-owner (gather) assembly: a thread evaluates, for ONE test function k of a cell, the
// contributions that cell makes to the rows of "its" vertex. No scatter, no atomics.
//
// Math follows the reference local operators (file:line relative to the reference repo):
//   dune/ax1/acme2_cyl/operator/acme2_cyl_operator_fully_coupled.hh:318-348 (elec residual)
//   ...:516-561 (analytic Jacobian), ...:790-830 (membrane-interface boundary term)
//   dune/ax1/acme2_cyl/operator/convectiondiffusionfem.hh:211-212, 317-318 (membrane Poisson)
//   dune/ax1/acme2_cyl/operator/acme2_cyl_toperator.hh:150, 251 (mass term)
//   dune/ax1/common/membranefluxgridfunction_implicit.hh:198-311, 353 (channel flux)
//   dune/ax1/acme2_cyl/common/acme2_cyl_geometrywrapper.hh:56-72 (2 pi r weight)
#pragma once
#include "common.cuh"

namespace ax1 {

struct CellCtx {
  int ci, cj;
  double hx, hy, ihx, ihy, y0;
  double xl[16];  // comp*4 + local vertex, comp = Na,K,Cl,pot
};

__device__ __forceinline__ void ld4(const double* p, double& a, double& b, double& c, double& d) {
  const double2 lo = *reinterpret_cast<const double2*>(p);
  const double2 hi = *reinterpret_cast<const double2*>(p + 2);
  a = lo.x; b = lo.y; c = hi.x; d = hi.y;
}

__device__ __forceinline__ void load_cell(const Dev& g, const double* u, int ci, int cj, CellCtx& c) {
  c.ci = ci; c.cj = cj;
  const double x0 = g.x[ci], x1 = g.x[ci + 1], y0 = g.y[cj], y1 = g.y[cj + 1];
  c.hx = x1 - x0; c.hy = y1 - y0; c.ihx = 1.0 / c.hx; c.ihy = 1.0 / c.hy; c.y0 = y0;
  const int v0 = ci + g.nvx * cj;
  const int vv[4] = {v0, v0 + 1, v0 + g.nvx, v0 + g.nvx + 1};
#pragma unroll
  for (int k = 0; k < 4; ++k)
    ld4(u + 4 * (size_t)vv[k], c.xl[k], c.xl[4 + k], c.xl[8 + k], c.xl[12 + k]);
}

__device__ __forceinline__ void q1(double xi, double eta, double* phi, double* dxi, double* deta) {
  phi[0] = (1 - xi) * (1 - eta); phi[1] = xi * (1 - eta); phi[2] = (1 - xi) * eta; phi[3] = xi * eta;
  dxi[0] = -(1 - eta); dxi[1] = (1 - eta); dxi[2] = -eta; dxi[3] = eta;
  deta[0] = -(1 - xi); deta[1] = -xi; deta[2] = (1 - xi); deta[3] = xi;
}

// Na injection source of the stimulated cell (default_nernst_planck_parameters.hh:168-216)
__device__ __forceinline__ double source_na(const Dev& g, const CellCtx& c) {
  if (!g.stim_on) return 0.0;
  const double x0 = g.x[c.ci], x1 = g.x[c.ci + 1], y0 = g.y[c.cj], y1 = g.y[c.cj + 1];
  if (x0 <= g.stim_x && y0 <= g.stim_y && x1 >= g.stim_x && y1 >= g.stim_y) {
    double vol = (con_pi * (y1 + y0)) * (c.hx * c.hy);
    double con_inj = g.I_inj;
    con_inj /= vol;
    con_inj *= g.time_scale;
    return con_inj;
  }
  return 0.0;
}

struct QP {
  double phi[4], gx[4], gy[4], factor;
};
__device__ __forceinline__ void make_qp(const CellCtx& c, double xi, double eta, QP& q) {
  double dxi[4], deta[4];
  q1(xi, eta, q.phi, dxi, deta);
#pragma unroll
  for (int k = 0; k < 4; ++k) { q.gx[k] = c.ihx * dxi[k]; q.gy[k] = c.ihy * deta[k]; }
  const double yq = c.y0 + eta * c.hy;
  q.factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
}

// residual rows of test function k of an electrolyte cell: spatial part (weight dt) + mass (weight wm)
__device__ __forceinline__ void elec_volume_row(const Dev& g, const CellCtx& c, int k, double vs, double wm,
                                                bool spatial, double out[4]) {
  const double fna = spatial ? source_na(g, c) : 0.0;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      QP q;
      make_qp(c, a ? AX1_GP1 : AX1_GP0, b ? AX1_GP1 : AX1_GP0, q);
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
        const double Apx = (-g.eps_elec) * gpx, Apy = (-g.eps_elec) * gpy;
        double chargeDensity = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) chargeDensity += g.valence[j] * uCon[j];
        const double fPot = -chargeDensity * g.PC;
        {
          double val = (Apx * q.gx[k] + Apy * q.gy[k]) + (-fPot) * q.phi[k];
          out[3] += g.dt * (val * q.factor * vs * g.S_pot);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
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
      for (int j = 0; j < 3; ++j) out[j] += wm * (uCon[j] * q.phi[k] * q.factor * vs);
    }
}

// potential row of test function k of a membrane cell (Poisson with eps_memb, f = 0)
__device__ __forceinline__ void memb_volume_row(const Dev& g, const CellCtx& c, int k, double vs, double& out) {
  const double eps = g.eps_memb[c.ci];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      QP q;
      make_qp(c, a ? AX1_GP1 : AX1_GP0, b ? AX1_GP1 : AX1_GP0, q);
      double gpx = 0, gpy = 0;
#pragma unroll
      for (int m = 0; m < 4; ++m) { gpx += c.xl[12 + m] * q.gx[m]; gpy += c.xl[12 + m] * q.gy[m]; }
      const double Ax = (-eps) * gpx, Ay = (-eps) * gpy;
      double val = (Ax * q.gx[k] + Ay * q.gy[k]);
      out += g.dt * (val * q.factor * vs * g.S_pot);
    }
}

// ------------------------------------------------------------------------------------------
// membrane flux at one quadrature point of an interface face of electrolyte cell column ci.
// this-side values uPot,uCon are local; con2/pot2 are the face-centre values of the current
// iterate on the opposite interface edge (acme2_cyl_physics.hh:687-696, 950-960).
// Returns j_species * (A_ext / A_this) for the three species (Cl has no channel -> 0).
__device__ __forceinline__ void membrane_flux(const Dev& g, int ci, bool from_cytosol, double uPot,
                                              const double uCon[3], double pot2, const double con2[3],
                                              double surfaceScale, double stimFlux, double jout[3]) {
  const int ySign = from_cytosol ? -1 : 1;
  const double n_y = from_cytosol ? 1.0 : -1.0;
  double potJump = pot2;
  potJump -= uPot;
  potJump *= ySign;
  const double C1 = (con_k * g.temperature) / con_e;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double conRatio = ySign > 0 ? con2[j] / uCon[j] : uCon[j] / con2[j];
    const int valence = g.valence[j];
    const double reversal = -log(conRatio) * (C1 / valence);
    double total = 0.0;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {  // channels of species j: leak (kk=0) then voltage gated (kk=1)
      if (j == CL) continue;
      const int ch = j + 2 * kk;      // 0 Na_leak, 1 K_leak, 2 Nav, 3 Kv
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
    }
    double yj = total * n_y;
    if (j == g.fstim_ion) yj += stimFlux;   // default_nernst_planck_parameters.hh:439-475
    yj *= surfaceScale;
    jout[j] = yj;
  }
}

struct FaceCtx {
  bool from_cytosol;
  double eta, y_this, surfaceScale, factor;  // factor = 0.5 * 2 pi y L
  double stimFlux;                           // flux stimulation of this face (0 if none)
  double pot2, con2[3];
};

__device__ __forceinline__ void make_face(const Dev& g, const double* u, const CellCtx& c, FaceCtx& f) {
  f.from_cytosol = (c.cj == g.jm - 1);
  f.eta = f.from_cytosol ? 1.0 : 0.0;
  f.y_this = f.from_cytosol ? g.y[g.jm] : g.y[g.jm + 1];
  const double y_es = g.y[g.jm + 1];
  const double my_surface = (con_pi * (f.y_this + f.y_this)) * c.hx;
  const double ext_surface = (con_pi * (y_es + y_es)) * c.hx;
  f.surfaceScale = ext_surface / my_surface;
  f.factor = 0.5 * ((2 * con_pi * f.y_this) * c.hx);
  f.stimFlux = 0.0;
  if (g.fstim_on && (g.x[c.ci] - 1e-6 < g.stim_x && g.x[c.ci + 1] + 1e-6 > g.stim_x)) {
    double sf = g.fstim_inj / my_surface;
    sf *= g.fstim_amp;
    sf *= f.from_cytosol ? 1.0 : -1.0;
    f.stimFlux = sf;
  }
  const int jopp = f.from_cytosol ? g.jm + 1 : g.jm;
  const double* ua = u + 4 * (size_t)(c.ci + g.nvx * jopp);
  double a0, a1, a2, a3, b0, b1, b2, b3;
  ld4(ua, a0, a1, a2, a3);
  ld4(ua + 4, b0, b1, b2, b3);
  f.con2[0] = 0.5 * a0 + 0.5 * b0; f.con2[1] = 0.5 * a1 + 0.5 * b1; f.con2[2] = 0.5 * a2 + 0.5 * b2;
  f.pot2 = 0.5 * a3 + 0.5 * b3;
}

// boundary contribution of the interface face to the concentration rows of test function k,
// evaluated with local coefficients xl (may be FD-perturbed)
__device__ __forceinline__ void elec_boundary_row(const Dev& g, const FaceCtx& f, int ci, const double* xl, int k,
                                                  double vs, double out[3]) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    double phi[4], dxi[4], deta[4];
    q1(s ? AX1_GP1 : AX1_GP0, f.eta, phi, dxi, deta);
    double uPot = 0.0, uCon[3];
#pragma unroll
    for (int m = 0; m < 4; ++m) uPot += xl[12 + m] * phi[m];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double t = 0.0;
#pragma unroll
      for (int m = 0; m < 4; ++m) t += xl[4 * j + m] * phi[m];
      uCon[j] = t;
    }
    double jf[3];
    membrane_flux(g, ci, f.from_cytosol, uPot, uCon, f.pot2, f.con2, f.surfaceScale, f.stimFlux, jf);
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] += g.dt * (jf[j] * phi[k] * f.factor * vs);
  }
}

}  // namespace ax1
