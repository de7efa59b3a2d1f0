// oracle/assembly.cpp -- TEST INFRASTRUCTURE ONLY.
// Restatement of the element kernels and of the PDELab assembly semantics of dune-ax1:
//   elec residual     dune/ax1/acme2_cyl/operator/acme2_cyl_operator_fully_coupled.hh:91-351
//   elec boundary     ...:676-926  (+ default_nernst_planck_parameters.hh:417-509,
//                      membranefluxgridfunction_implicit.hh:83-361, acme2_cyl_physics.hh:666-721,927-1013)
//   analytic Jacobian ...:356-563
//   membrane Poisson  dune/ax1/acme2_cyl/operator/convectiondiffusionfem.hh:86-215,230-322
//   mass operator     dune/ax1/acme2_cyl/operator/acme2_cyl_toperator.hh:56-154,168-256
//   FD Jacobian       PDELab NumericalJacobian{Volume,Boundary}; formula stuff/idefault_patch.dat:14-23
//   one-step method   implicit Euler r = M(u) - M(uold) + dt*A(u)  (acme2_cyl_setup.hh:704-708,861-868)
//   constraints       rows of constrained DOFs -> unit rows; their COLUMNS are kept (PDELab's default,
//                     non-symmetric etadd) -- the overlapping operator needs the front-vertex columns [ext]
#include "oracle.hpp"
#include <stdexcept>
#include <algorithm>
#include <cstring>

namespace ax1o {

double eff_conductance_new(const Problem& p, int k, int i);  // model.cpp

namespace {

const double GP[2] = {0.2113248654051871177454256097490212721762,
                      0.7886751345948128822545743902509787278238};  // dune-geometry Gauss-Legendre, 2 pts

inline void q1(double xi, double eta, double phi[4], double dxi[4], double deta[4]) {
  phi[0] = (1 - xi) * (1 - eta); phi[1] = xi * (1 - eta); phi[2] = (1 - xi) * eta; phi[3] = xi * eta;
  dxi[0] = -(1 - eta); dxi[1] = (1 - eta); dxi[2] = -eta; dxi[3] = eta;
  deta[0] = -(1 - xi); deta[1] = -xi; deta[2] = (1 - xi); deta[3] = xi;
}

struct Cell {
  int i, j, v[4];
  double hx, hy, ihx, ihy, y0, vs[4];
};

Cell make_cell(const Problem& p, int i, int j) {
  Cell c;
  c.i = i; c.j = j;
  c.v[0] = p.vid(i, j); c.v[1] = p.vid(i + 1, j); c.v[2] = p.vid(i, j + 1); c.v[3] = p.vid(i + 1, j + 1);
  c.hx = p.x[i + 1] - p.x[i]; c.hy = p.y[j + 1] - p.y[j];
  c.ihx = 1.0 / c.hx; c.ihy = 1.0 / c.hy; c.y0 = p.y[j];
  for (int k = 0; k < 4; ++k) c.vs[k] = p.cfg.do_volume_scaling ? 1. / p.node_volume[c.v[k]] : 1.0;
  return c;
}

// Na injection source, default_nernst_planck_parameters.hh:168-216 (time = LOP time t+dt)
double source_f(const Problem& p, const Cell& c, int species) {
  const Config& g = p.cfg;
  double t = p.time + p.dt;
  if (g.do_stimulation && (t > g.tinj_start && t < g.tinj_end)) {
    bool contains = p.x[c.i] <= g.stim_x && p.y[c.j] <= g.stim_y && p.x[c.i + 1] >= g.stim_x &&
                    p.y[c.j + 1] >= g.stim_y;
    if (contains && species == Na) {
      double vol = (con_pi * (p.y[c.j + 1] + p.y[c.j])) * (c.hx * c.hy);
      double con_inj = g.I_inj;
      con_inj /= vol;
      con_inj *= g.time_scale;
      return con_inj;
    }
  }
  return 0.0;
}

// local coefficient order: comp*4 + vertex, comp = Na,K,Cl,pot
void elec_alpha_volume(const Problem& p, const Cell& c, const double* xl, double weight, double* r,
                       double* ra = nullptr) {
  const double eps = p.cfg.eps_elec;
  const double S = p.cfg.pot_residual_scale;
  double fsrc[NSPEC];
  for (int j = 0; j < NSPEC; ++j) fsrc[j] = source_f(p, c, j);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4], gx[4], gy[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
      double uPot = 0.0, uCon[NSPEC];
      for (int k = 0; k < 4; ++k) uPot += xl[12 + k] * phi[k];
      for (int j = 0; j < NSPEC; ++j) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += xl[4 * j + k] * phi[k];
        uCon[j] = s;
      }
      double gpx = 0, gpy = 0, gcx[NSPEC], gcy[NSPEC];
      for (int k = 0; k < 4; ++k) { gpx += xl[12 + k] * gx[k]; gpy += xl[12 + k] * gy[k]; }
      for (int j = 0; j < NSPEC; ++j) {
        double sx = 0, sy = 0;
        for (int k = 0; k < 4; ++k) { sx += xl[4 * j + k] * gx[k]; sy += xl[4 * j + k] * gy[k]; }
        gcx[j] = sx; gcy[j] = sy;
      }
      // magnitudes before cancellation (forward-error scale for the parity tolerance)
      double agpx = 0, agpy = 0, achg = 0;
      if (ra) {
        for (int k = 0; k < 4; ++k) { agpx += std::fabs(xl[12 + k] * gx[k]); agpy += std::fabs(xl[12 + k] * gy[k]); }
        for (int j = 0; j < NSPEC; ++j) achg += std::fabs(uCon[j]);
      }
      // A grad u; Poisson tensor = -eps I (default_poisson_parameters.hh:86-105)
      double Apx = (-eps) * gpx, Apy = (-eps) * gpy;
      double chargeDensity = 0.0;
      for (int j = 0; j < NSPEC; ++j) chargeDensity += p.cfg.valence[j] * uCon[j];
      double fPot = -chargeDensity * p.PC;
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int k = 0; k < 4; ++k) {
        double val = (Apx * gx[k] + Apy * gy[k]) - uPot * (0.0 * gx[k] + 0.0 * gy[k]) + (0.0 * uPot - fPot) * phi[k];
        r[12 + k] += weight * (val * factor * c.vs[k] * S);
        if (ra) ra[12 + k] += std::fabs(weight) * ((eps * agpx * std::fabs(gx[k]) + eps * agpy * std::fabs(gy[k]) + p.PC * achg * phi[k]) * factor * c.vs[k] * S);
      }
      for (int j = 0; j < NSPEC; ++j) {
        double D = p.D[j];
        double Acx = D * gcx[j], Acy = D * gcy[j];
        double bx = -D * p.cfg.valence[j] * gpx, by = -D * p.cfg.valence[j] * gpy;
        for (int k = 0; k < 4; ++k) {
          double val = (Acx * gx[k] + Acy * gy[k]) - uCon[j] * (bx * gx[k] + by * gy[k]) +
                       (0.0 * uCon[j] - fsrc[j]) * phi[k];
          r[4 * j + k] += weight * (val * factor * c.vs[k]);
          if (ra) {
            double agcx = 0, agcy = 0;
            for (int m = 0; m < 4; ++m) { agcx += std::fabs(xl[4 * j + m] * gx[m]); agcy += std::fabs(xl[4 * j + m] * gy[m]); }
            ra[4 * j + k] += std::fabs(weight) * ((D * agcx * std::fabs(gx[k]) + D * agcy * std::fabs(gy[k]) +
                                                   std::fabs(uCon[j]) * D * (agpx * std::fabs(gx[k]) + agpy * std::fabs(gy[k])) +
                                                   std::fabs(fsrc[j] * phi[k])) * factor * c.vs[k]);
          }
        }
      }
    }
}

void memb_alpha_volume(const Problem& p, const Cell& c, const double* xpot, double weight, double* r,
                       double* ra = nullptr) {
  const double eps = p.eps_memb[c.i];
  const double S = p.cfg.pot_residual_scale;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4], gx[4], gy[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
      double u = 0, gpx = 0, gpy = 0;
      for (int k = 0; k < 4; ++k) u += xpot[k] * phi[k];
      for (int k = 0; k < 4; ++k) { gpx += xpot[k] * gx[k]; gpy += xpot[k] * gy[k]; }
      double Ax = (-eps) * gpx, Ay = (-eps) * gpy;
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int k = 0; k < 4; ++k) {
        double val = (Ax * gx[k] + Ay * gy[k]) - u * (0.0) + (0.0 * u - 0.0) * phi[k];
        r[k] += weight * (val * factor * c.vs[k] * S);
        if (ra) {
          double agx = 0, agy = 0;
          for (int m = 0; m < 4; ++m) { agx += std::fabs(xpot[m] * gx[m]); agy += std::fabs(xpot[m] * gy[m]); }
          ra[k] += std::fabs(weight) * ((eps * agx * std::fabs(gx[k]) + eps * agy * std::fabs(gy[k])) * factor * c.vs[k] * S);
        }
      }
    }
}

void time_alpha_volume(const Cell& c, const double* xcon, double weight, double* r, double* ra = nullptr) {
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int j = 0; j < NSPEC; ++j) {
        double con = 0.0;
        for (int k = 0; k < 4; ++k) con += xcon[4 * j + k] * phi[k];
        for (int k = 0; k < 4; ++k) r[4 * j + k] += weight * (con * phi[k] * factor * c.vs[k]);
        if (ra) for (int k = 0; k < 4; ++k) ra[4 * j + k] += std::fabs(weight * (con * phi[k] * factor * c.vs[k]));
      }
    }
}

// membrane-interface face of an elec cell: face 3 (top) of the cytosol row jm-1, face 2 (bottom)
// of the ES row jm+1. uglob = current global iterate (opposite side is read from it).
void elec_alpha_boundary(const Problem& p, const Cell& c, const double* xl, const double* uglob,
                         double weight, double* r, double* ra = nullptr) {
  const Config& g = p.cfg;
  // [stimulation] memb_flux_stimulation (default_nernst_planck_parameters.hh:439-475): time = LOP time t+dt
  double stimFlux = 0.0;
  {
    const double t = p.time + p.dt;
    if (g.do_memb_flux_stimulation && (t > g.tinj_start && t < g.tinj_end) &&
        (p.x[c.i] - 1e-6 < g.stim_x && p.x[c.i + 1] + 1e-6 > g.stim_x)) {
      const double y_face = (c.j == p.jm - 1) ? p.y[p.jm] : p.y[p.jm + 1];
      stimFlux = g.memb_flux_inj / ((con_pi * (y_face + y_face)) * c.hx);   // / igeo.volume()
      const double t_rel = (t - g.tinj_start) / g.tinj_end;
      stimFlux *= 0.5 * (1 - std::cos(2 * con_pi * t_rel));
      stimFlux *= (c.j == p.jm - 1) ? 1.0 : -1.0;                            // centerUnitOuterNormal()[1]
    }
  }
  // passive membrane: channel flux y = 0 (flux*normal*scale accumulates zeros)
  if (!g.membrane_active && stimFlux == 0.0) return;
  const bool from_cytosol = (c.j == p.jm - 1);
  const double eta = from_cytosol ? 1.0 : 0.0;
  const double y_this = from_cytosol ? p.y[p.jm] : p.y[p.jm + 1];
  const double y_es = p.y[p.jm + 1];
  const int jopp = from_cytosol ? p.jm + 1 : p.jm;  // vertex row of the opposite interface edge
  const double n_y = from_cytosol ? 1.0 : -1.0;
  // opposite-side face-centre values of the current iterate (physics.hh:687-696, 950-960)
  double con2[NSPEC], pot2;
  {
    const double* ua = &uglob[4 * p.vid(c.i, jopp)];
    const double* ub = &uglob[4 * p.vid(c.i + 1, jopp)];
    for (int j = 0; j < NSPEC; ++j) con2[j] = 0.5 * ua[j] + 0.5 * ub[j];
    pot2 = 0.5 * ua[POT] + 0.5 * ub[POT];
  }
  const int ySign = from_cytosol ? -1 : 1;  // sign(y1 - y2), x1 = this side
  const bool equil = g.do_equilibration && p.time < g.t_equilibrium;
  static const int species_of[NCHAN] = {Na, K, Na, K};
  const double my_surface = (con_pi * (y_this + y_this)) * c.hx;   // igeo.volume()
  const double ext_surface = (con_pi * (y_es + y_es)) * c.hx;      // physics.getMembraneSurfaceArea
  const double surfaceScaleFactor = ext_surface / my_surface;
  for (int s = 0; s < 2; ++s) {
    double phi[4], dxi[4], deta[4];
    q1(GP[s], eta, phi, dxi, deta);
    double uPot = 0.0, uCon[NSPEC];
    for (int k = 0; k < 4; ++k) uPot += xl[12 + k] * phi[k];
    for (int j = 0; j < NSPEC; ++j) {
      double t = 0.0;
      for (int k = 0; k < 4; ++k) t += xl[4 * j + k] * phi[k];
      uCon[j] = t;
    }
    double potJump = pot2;
    potJump -= uPot;
    potJump *= ySign;
    double factor = 0.5 * ((2 * con_pi * y_this) * c.hx);
    for (int j = 0; j < NSPEC; ++j) {
      double conRatio = ySign > 0 ? con2[j] / uCon[j] : uCon[j] / con2[j];
      int valence = g.valence[j];
      double C1 = (con_k * g.temperature) / con_e;
      double reversal_potential = -std::log(conRatio) * (C1 / valence);
      double totalFlux = 0.0, absFlux = 0.0;
      for (int k = 0; k < NCHAN; ++k) {
        if (species_of[k] != j) continue;
        bool leak = k < 2;
        double conductance = (equil && !leak) ? 0.0 : eff_conductance_new(p, k, c.i);
        if (!g.membrane_active) conductance = 0.0;
        double C = (10 * conductance) / (con_e * valence * con_mol);
        if (equil) C *= g.dummy_value;
        C *= g.time_scale / g.length_scale;
        double diffTerm = C * reversal_potential;
        if (std::fabs(conRatio - 1) < 1e-8 || std::isnan(conRatio)) diffTerm = 0.0;
        double driftTerm = C * C1 * potJump;
        double flux = driftTerm;
        flux -= diffTerm;
        totalFlux += flux;
        absFlux += std::fabs(C * C1) * (std::fabs(pot2) + std::fabs(uPot)) + std::fabs(diffTerm);
      }
      double yj = totalFlux * n_y;
      if (j == g.memb_flux_ion) yj += stimFlux;
      yj *= surfaceScaleFactor;
      for (int k = 0; k < 4; ++k) r[4 * j + k] += weight * (yj * phi[k] * factor * c.vs[k]);
      if (j == g.memb_flux_ion) absFlux += std::fabs(stimFlux);
      if (ra) for (int k = 0; k < 4; ++k) ra[4 * j + k] += std::fabs(weight * (absFlux * surfaceScaleFactor * phi[k] * factor * c.vs[k]));
    }
  }
}

void elec_jacobian_volume(const Problem& p, const Cell& c, const double* xl, double weight, double* m /*16x16*/) {
  const double eps = p.cfg.eps_elec;
  const double S = p.cfg.pot_residual_scale;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4], gx[4], gy[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
      double uCon[NSPEC];
      for (int j = 0; j < NSPEC; ++j) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += xl[4 * j + k] * phi[k];
        uCon[j] = s;
      }
      double gpx = 0, gpy = 0;
      for (int k = 0; k < 4; ++k) { gpx += xl[12 + k] * gx[k]; gpy += xl[12 + k] * gy[k]; }
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int k = 0; k < NSPEC; ++k) {
        double D = p.D[k];
        int z = p.cfg.valence[k];
        double bx = -D * z * gpx, by = -D * z * gpy;
        for (int jj = 0; jj < 4; ++jj) {
          for (int ii = 0; ii < 4; ++ii) {  // con-con
            double val = ((D * gx[jj]) * gx[ii] + (D * gy[jj]) * gy[ii]) - phi[jj] * (bx * gx[ii] + by * gy[ii]) +
                         0.0 * phi[jj] * phi[ii];
            m[(4 * k + ii) * 16 + (4 * k + jj)] += weight * (val * factor * c.vs[ii]);
          }
          for (int ii = 0; ii < 4; ++ii) {  // pot row <- con column
            double dfPot_du = -z * phi[jj] * p.PC;
            m[(12 + ii) * 16 + (4 * k + jj)] += weight * ((-dfPot_du * phi[ii]) * factor * c.vs[ii] * S);
          }
        }
      }
      for (int jj = 0; jj < 4; ++jj) {
        for (int ii = 0; ii < 4; ++ii) {  // pot-pot
          double val = (((-eps) * gx[jj]) * gx[ii] + ((-eps) * gy[jj]) * gy[ii]) - phi[jj] * 0.0 + 0.0 * phi[jj] * phi[ii];
          m[(12 + ii) * 16 + (12 + jj)] += weight * (val * factor * c.vs[ii] * S);
        }
        for (int k = 0; k < NSPEC; ++k) {  // con row <- pot column
          double D = p.D[k];
          int z = p.cfg.valence[k];
          double dbx = gx[jj] * (-D * z), dby = gy[jj] * (-D * z);
          for (int ii = 0; ii < 4; ++ii)
            m[(4 * k + ii) * 16 + (12 + jj)] += weight * ((-uCon[k] * (dbx * gx[ii] + dby * gy[ii])) * factor * c.vs[ii]);
        }
      }
    }
}

void memb_jacobian_volume(const Problem& p, const Cell& c, double weight, double* m /*4x4*/) {
  const double eps = p.eps_memb[c.i];
  const double S = p.cfg.pot_residual_scale;
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4], gx[4], gy[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int jj = 0; jj < 4; ++jj)
        for (int ii = 0; ii < 4; ++ii) {
          double val = ((-eps) * gx[jj]) * gx[ii] + ((-eps) * gy[jj]) * gy[ii];
          m[ii * 4 + jj] += weight * (val * factor * c.vs[ii] * S);
        }
    }
}

void time_jacobian_volume(const Cell& c, double weight, double* m /*12x12*/) {
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      double phi[4], dxi[4], deta[4];
      q1(GP[a], GP[b], phi, dxi, deta);
      double yq = c.y0 + GP[b] * c.hy;
      double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
      for (int k = 0; k < NSPEC; ++k)
        for (int jj = 0; jj < 4; ++jj)
          for (int ii = 0; ii < 4; ++ii)
            m[(4 * k + ii) * 12 + (4 * k + jj)] += weight * (phi[jj] * phi[ii] * factor * c.vs[ii]);
    }
}

inline void gather(const Cell& c, const double* u, double* xl) {
  for (int comp = 0; comp < 4; ++comp)
    for (int k = 0; k < 4; ++k) xl[4 * comp + k] = u[4 * c.v[k] + comp];
}

inline bool is_interface_cell(const Problem& p, int j) { return j == p.jm - 1 || j == p.jm + 1; }

}  // namespace

// The five output terms of MembraneFluxGridFunctionImplicit::evaluate (membranefluxgridfunction_implicit.hh:83-361;
// FluxTerms 0 total, 1 diffusion term, 2 drift term, 3 leak channels, 4 voltage-gated channels) for one membrane
// column i, seen from the cytosol side (outer normal +y) at the face centres of the two interface edges -- what
// acme2_cyl_output.hh:1215-1410 writes per ion. uglob = current iterate; the channel state is the NEW one.
// out15[term * 3 + ion].
void membrane_flux_terms(const Problem& p, int i, const double* uglob, double* out15) {
  const Config& g = p.cfg;
  for (int k = 0; k < 15; ++k) out15[k] = 0.0;
  if (!g.membrane_active) return;
  double con1[NSPEC], con2[NSPEC], pot1, pot2;
  {
    const double* ua = &uglob[4 * p.vid(i, p.jm)];
    const double* ub = &uglob[4 * p.vid(i + 1, p.jm)];
    for (int j = 0; j < NSPEC; ++j) con1[j] = 0.5 * ua[j] + 0.5 * ub[j];
    pot1 = 0.5 * ua[POT] + 0.5 * ub[POT];
    const double* uc = &uglob[4 * p.vid(i, p.jm + 1)];
    const double* ud = &uglob[4 * p.vid(i + 1, p.jm + 1)];
    for (int j = 0; j < NSPEC; ++j) con2[j] = 0.5 * uc[j] + 0.5 * ud[j];
    pot2 = 0.5 * uc[POT] + 0.5 * ud[POT];
  }
  const int ySign = -1;        // this side = cytosol
  const double n_y = 1.0;
  const bool equil = g.do_equilibration && p.time < g.t_equilibrium;
  static const int species_of[NCHAN] = {Na, K, Na, K};
  double potJump = pot2;
  potJump -= pot1;
  potJump *= ySign;
  for (int j = 0; j < NSPEC; ++j) {
    double conRatio = ySign > 0 ? con2[j] / con1[j] : con1[j] / con2[j];
    int valence = g.valence[j];
    double C1 = (con_k * g.temperature) / con_e;
    double reversal_potential = -std::log(conRatio) * (C1 / valence);
    double total = 0.0, diff = 0.0, drift = 0.0, leakf = 0.0, vg = 0.0;
    for (int k = 0; k < NCHAN; ++k) {
      if (species_of[k] != j) continue;
      bool leak = k < 2;
      double conductance = (equil && !leak) ? 0.0 : eff_conductance_new(p, k, i);
      double C = (10 * conductance) / (con_e * valence * con_mol);
      if (equil) C *= g.dummy_value;
      C *= g.time_scale / g.length_scale;
      double diffTerm = C * reversal_potential;
      if (std::fabs(conRatio - 1) < 1e-8 || std::isnan(conRatio)) diffTerm = 0.0;
      double driftTerm = C * C1 * potJump;
      double flux = driftTerm;
      flux -= diffTerm;
      diff += diffTerm; drift += driftTerm; total += flux;
      if (leak) leakf += flux; else vg += flux;
    }
    out15[0 + j] = total * n_y; out15[3 + j] = diff * n_y; out15[6 + j] = drift * n_y;
    out15[9 + j] = leakf * n_y; out15[12 + j] = vg * n_y;
  }
}


// one local kernel on cell (i, j) with given local coefficients xl[16] (comp * 4 + vertex, comp = Na, K, Cl, pot):
// which = 0 elec spatial alpha_volume, 1 time-operator alpha_volume (mass), 2 membrane Poisson alpha_volume,
// 3 elec alpha_boundary on the membrane-interface face (uglob supplies the opposite side). weight as the
// one-step method passes it (dt for the spatial operator, 1 for the time operator). out[16] is overwritten.
// Used by tests/test_reference_pins.py to compare with the reference's own local operators (oracle/_ref).
void cell_kernel(const Problem& p, int i, int j, int which, const double* xl, const double* uglob, double weight,
                 double* out) {
  Cell c = make_cell(p, i, j);
  for (int k = 0; k < 16; ++k) out[k] = 0.0;
  switch (which) {
    case 0: elec_alpha_volume(p, c, xl, weight, out); break;
    case 1: time_alpha_volume(c, xl, weight, out); break;
    case 2: memb_alpha_volume(p, c, xl + 12, weight, out + 12); break;
    case 3: elec_alpha_boundary(p, c, xl, uglob, weight, out); break;
    default: throw std::runtime_error("cell_kernel: unknown kernel");
  }
}

// local Jacobian of one kernel as a 16 x 16 matrix (row = comp_i * 4 + vertex_i, column = comp_j * 4 + vertex_j):
// which = 0 elec spatial (analytic), 1 time operator (analytic), 2 membrane Poisson (analytic)
void cell_jacobian(const Problem& p, int i, int j, int which, const double* xl, double weight, double* out256) {
  Cell c = make_cell(p, i, j);
  for (int k = 0; k < 256; ++k) out256[k] = 0.0;
  if (which == 0) {
    elec_jacobian_volume(p, c, xl, weight, out256);
  } else if (which == 1) {
    double m[144] = {0};
    time_jacobian_volume(c, weight, m);
    for (int a = 0; a < 12; ++a) for (int b = 0; b < 12; ++b) out256[a * 16 + b] = m[a * 12 + b];
  } else if (which == 2) {
    double m[16] = {0};
    memb_jacobian_volume(p, c, weight, m);
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) out256[(12 + a) * 16 + 12 + b] = m[a * 4 + b];
  } else {
    throw std::runtime_error("cell_jacobian: unknown kernel");
  }
}

void set_uold(Problem& p, const double* uold) {
  p.uold.assign(uold, uold + 4 * (size_t)p.nv);
  p.r0.assign(4 * (size_t)p.nv, 0.0);
  // const residual of implicit Euler: a = {-1, 1}, b = {0, 1}  ->  r0 = -M(uold)   [ext]
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      if (p.cell_sub[p.cid(i, j)] == MEMBRANE) continue;
      Cell c = make_cell(p, i, j);
      double xl[16], rl[12] = {0};
      gather(c, uold, xl);
      time_alpha_volume(c, xl, -1.0, rl);
      for (int comp = 0; comp < 3; ++comp)
        for (int k = 0; k < 4; ++k) p.r0[4 * c.v[k] + comp] += rl[4 * comp + k];
    }
}

void residual(const Problem& p, const double* u, double* r, double* abs_terms) {
  const size_t n = 4 * (size_t)p.nv;
  std::memset(r, 0, n * sizeof(double));
  if (abs_terms) std::memset(abs_terms, 0, n * sizeof(double));
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      Cell c = make_cell(p, i, j);
      double xl[16], rl[16] = {0}, al[16] = {0};
      double* ra = abs_terms ? al : nullptr;
      gather(c, u, xl);
      if (p.cell_sub[p.cid(i, j)] == MEMBRANE) {
        memb_alpha_volume(p, c, xl + 12, p.dt, rl + 12, ra ? ra + 12 : nullptr);
      } else {
        elec_alpha_volume(p, c, xl, p.dt, rl, ra);
        time_alpha_volume(c, xl, 1.0, rl, ra);
        if (is_interface_cell(p, j)) elec_alpha_boundary(p, c, xl, u, p.dt, rl, ra);
      }
      for (int comp = 0; comp < 4; ++comp)
        for (int k = 0; k < 4; ++k) {
          r[4 * c.v[k] + comp] += rl[4 * comp + k];
          if (abs_terms) abs_terms[4 * c.v[k] + comp] += al[4 * comp + k];
        }
    }
  for (size_t k = 0; k < n; ++k) {
    r[k] += p.r0[k];
    if (abs_terms) abs_terms[k] += std::fabs(p.r0[k]);
  }
  for (int v = 0; v < p.nv; ++v)
    for (int comp = 0; comp < 4; ++comp)
      if (p.dirichlet[v] & (1 << comp)) r[4 * v + comp] = 0.0;
}

void jacobian(const Problem& p, const double* u, BCRS& A) {
  std::fill(A.val.begin(), A.val.end(), 0.0);
  const double fd_eps = 1e-7;
  auto scatter = [&](const Cell& c, const double* m, const int* comps, int ncomp) {
    // m is (ncomp*4)x(ncomp*4), local index = ci*4 + vertex
    int nl = ncomp * 4;
    for (int ci = 0; ci < ncomp; ++ci)
      for (int vi = 0; vi < 4; ++vi)
        for (int cj = 0; cj < ncomp; ++cj)
          for (int vj = 0; vj < 4; ++vj) {
            double val = m[(ci * 4 + vi) * nl + (cj * 4 + vj)];
            if (val == 0.0) continue;
            int row = c.v[vi], col = c.v[vj];
            int ri = comps[ci], cjj = comps[cj];
            if ((p.dirichlet[row] >> ri) & 1) continue;   // constrained row dropped
            int k = A.find(row, col);
            A.block(k)[ri * 4 + cjj] += val;
          }
  };
  static const int comps_all[4] = {0, 1, 2, 3}, comps_con[3] = {0, 1, 2}, comps_pot[1] = {3};
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      Cell c = make_cell(p, i, j);
      double xl[16];
      gather(c, u, xl);
      bool memb = p.cell_sub[p.cid(i, j)] == MEMBRANE;
      if (memb) {
        double m[16] = {0};
        if (p.cfg.analytic_volume_jacobian & 2) {
          memb_jacobian_volume(p, c, p.dt, m);
        } else {
          double down[4] = {0}, xp[4];
          std::memcpy(xp, xl + 12, sizeof xp);
          memb_alpha_volume(p, c, xp, p.dt, down);
          for (int jj = 0; jj < 4; ++jj) {
            double up[4] = {0};
            double delta = fd_eps * (1.0 + std::fabs(xp[jj]));
            xp[jj] += delta;
            memb_alpha_volume(p, c, xp, p.dt, up);
            for (int ii = 0; ii < 4; ++ii) m[ii * 4 + jj] += (up[ii] - down[ii]) / delta;
            xp[jj] = xl[12 + jj];
          }
        }
        scatter(c, m, comps_pot, 1);
        continue;
      }
      {  // spatial volume term, weight dt
        double m[256] = {0};
        if (p.cfg.analytic_volume_jacobian & 1) {
          elec_jacobian_volume(p, c, xl, p.dt, m);
        } else {
          double down[16] = {0}, xp[16];
          std::memcpy(xp, xl, sizeof xp);
          elec_alpha_volume(p, c, xp, p.dt, down);
          for (int jj = 0; jj < 16; ++jj) {
            double up[16] = {0};
            double delta = fd_eps * (1.0 + std::fabs(xp[jj]));
            xp[jj] += delta;
            elec_alpha_volume(p, c, xp, p.dt, up);
            for (int ii = 0; ii < 16; ++ii) m[ii * 16 + jj] += (up[ii] - down[ii]) / delta;
            xp[jj] = xl[jj];
          }
        }
        scatter(c, m, comps_all, 4);
      }
      {  // mass term, weight 1
        double m[144] = {0};
        if (p.cfg.analytic_volume_jacobian & 4) {
          time_jacobian_volume(c, 1.0, m);
        } else {
          double down[12] = {0}, xp[12];
          std::memcpy(xp, xl, sizeof xp);
          time_alpha_volume(c, xp, 1.0, down);
          for (int jj = 0; jj < 12; ++jj) {
            double up[12] = {0};
            double delta = fd_eps * (1.0 + std::fabs(xp[jj]));
            xp[jj] += delta;
            time_alpha_volume(c, xp, 1.0, up);
            for (int ii = 0; ii < 12; ++ii) m[ii * 12 + jj] += (up[ii] - down[ii]) / delta;
            xp[jj] = xl[jj];
          }
        }
        scatter(c, m, comps_con, 3);
      }
      if (is_interface_cell(p, j)) {  // boundary term: always finite differences (no jacobian_boundary)
        double m[256] = {0};
        double down[16] = {0}, xp[16];
        std::memcpy(xp, xl, sizeof xp);
        elec_alpha_boundary(p, c, xp, u, p.dt, down);
        for (int jj = 0; jj < 16; ++jj) {
          double up[16] = {0};
          double delta = fd_eps * (1.0 + std::fabs(xp[jj]));
          xp[jj] += delta;
          elec_alpha_boundary(p, c, xp, u, p.dt, up);
          for (int ii = 0; ii < 16; ++ii) m[ii * 16 + jj] += (up[ii] - down[ii]) / delta;
          xp[jj] = xl[jj];
        }
        scatter(c, m, comps_all, 4);
      }
    }
  // trivial rows for constrained DOFs
  for (int v = 0; v < p.nv; ++v) {
    if (!p.dirichlet[v]) continue;
    int k = A.find(v, v);
    for (int comp = 0; comp < 4; ++comp)
      if ((p.dirichlet[v] >> comp) & 1) A.block(k)[comp * 4 + comp] = 1.0;
  }
}

// rownorm_preconditioner.hh:65-155: scale every scalar row of block row i (and r_i) by the
// inverse infinity norm of the matching row of the DIAGONAL block.
void rownorm_scale(BCRS& A, double* r) {
  for (int i = 0; i < A.n; ++i) {
    int kd = A.find(i, i);
    double s[4];
    for (int a = 0; a < 4; ++a) {
      double mx = 0.0;
      for (int b = 0; b < 4; ++b) mx = std::max(mx, std::fabs(A.block(kd)[4 * a + b]));
      s[a] = mx;
    }
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      for (int a = 0; a < 4; ++a) {
        if (s[a]) for (int b = 0; b < 4; ++b) A.block(k)[4 * a + b] /= s[a];
        else if (k == kd) A.block(k)[4 * a + a] = 1.0;
      }
    for (int a = 0; a < 4; ++a) if (s[a]) r[4 * i + a] /= s[a];
  }
}

}  // namespace ax1o
