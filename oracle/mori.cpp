// oracle/mori.cpp -- TEST INFRASTRUCTURE ONLY.
// Restatement of the electroneutral ("Mori") operator-split model of dune-ax1 (BASELINE configs[2]),
// one time step as Acme2CylMoriSimulation::run drives it (paths relative to the reference repository):
//   dune/ax1/acme2_cyl/common/acme2_cyl_mori_simulation.hh:562-600   ulast = unew; (1) potential solve with the last
//        concentrations, (2) charge-layer (gamma) update, (3) concentration solve with the last potential and
//        concentrations in the drift and membrane terms
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_potential_operator.hh  alpha_volume (conductivity tensor a(c) grad phi
//        + diffusion current), alpha_boundary (membrane Neumann current = ionic + capacitive)
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_concentration_operator.hh  alpha_volume (diffusion implicit, drift with
//        the last concentration / potential), alpha_boundary (channel flux + capacitive flux)
//   dune/ax1/acme2_cyl/configurations/mori/mori_poisson_parameters.hh   A :124-158, bGrad :170-186, f, j (sum z_j j_j)
//   dune/ax1/acme2_cyl/configurations/mori/mori_nernst_planck_parameters.hh  A, b, f, j = channel flux + Mori flux,
//        scaled by A_ext / A_this
//   dune/ax1/common/charge_layer_mori_implicit.hh   evaluate :83-250, updateState / init / timeStep / acceptTimeStep
//   dune/ax1/common/ax1_linearsubproblem.hh          one assemble-solve-update pass per sub-problem
//   dune/ax1/acme2_cyl/common/acme2_cyl_mori_setup.hh:569-575, 726-742, 866-893  wiring (no membrane contributions,
//        concentration block through OneStepMethod / implicit Euler, BiCGStab + ILU(n) per sub-block)
// Membrane elements contribute nothing (MoriConfiguration::Traits::useMembraneContributions = false); the two sides
// couple only through the interface flux, in which the OPPOSITE side enters through the stored solution vector
// (dgfPotNew), not through the unknowns of the linear solve -- exactly what the reference's NumericalJacobianBoundary
// sees, since it perturbs the inside element's coefficients only. Both sub-problems are linear in their unknowns, so
// the Jacobians are written analytically (the reference differentiates numerically; see DESIGN on the tolerance).
// The 4 x 4 vertex blocks of the coupled problem are reused with the other sub-problem's rows set to identity, which
// makes the sub-block solve the same BiCGStab / block-ILU as everywhere else in the oracle.
#include "oracle.hpp"
#include <cstring>
#include <stdexcept>

namespace ax1o {

double eff_conductance_new(const Problem& p, int k, int i);  // model.cpp

namespace {

const double GP[2] = {0.2113248654051871177454256097490212721762, 0.7886751345948128822545743902509787278238};

inline void q1(double xi, double eta, double phi[4], double dxi[4], double deta[4]) {
  phi[0] = (1 - xi) * (1 - eta); phi[1] = xi * (1 - eta); phi[2] = (1 - xi) * eta; phi[3] = xi * eta;
  dxi[0] = -(1 - eta); dxi[1] = (1 - eta); dxi[2] = -eta; dxi[3] = eta;
  deta[0] = -(1 - xi); deta[1] = -xi; deta[2] = (1 - xi); deta[3] = xi;
}

struct MCell {
  int i, j, v[4];
  double hx, hy, ihx, ihy, y0, vs[4];
};
MCell make_cell(const Problem& p, int i, int j) {
  MCell c;
  c.i = i; c.j = j;
  c.v[0] = p.vid(i, j); c.v[1] = p.vid(i + 1, j); c.v[2] = p.vid(i, j + 1); c.v[3] = p.vid(i + 1, j + 1);
  c.hx = p.x[i + 1] - p.x[i]; c.hy = p.y[j + 1] - p.y[j];
  c.ihx = 1.0 / c.hx; c.ihy = 1.0 / c.hy; c.y0 = p.y[j];
  for (int k = 0; k < 4; ++k) c.vs[k] = p.cfg.do_volume_scaling ? 1. / p.node_volume[c.v[k]] : 1.0;
  return c;
}
inline void gather(const MCell& c, const double* u, double* xl) {  // xl[comp * 4 + vertex]
  for (int k = 0; k < 4; ++k)
    for (int comp = 0; comp < 4; ++comp) xl[4 * comp + k] = u[4 * c.v[k] + comp];
}
inline bool is_interface_cell(const Problem& p, int j) { return j == p.jm - 1 || j == p.jm + 1; }

// Na injection source (mori_nernst_planck_parameters.hh f(); identical to the default parameter class)
double source_f(const Problem& p, const MCell& c, int species) {
  const Config& g = p.cfg;
  const double t = p.time + p.dt;
  if (g.do_stimulation && (t > g.tinj_start && t < g.tinj_end)) {
    const bool contains = p.x[c.i] <= g.stim_x && p.y[c.j] <= g.stim_y && p.x[c.i + 1] >= g.stim_x && p.y[c.j + 1] >= g.stim_y;
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

// everything about a membrane-interface face of an electrolyte cell that does not depend on the quadrature point
struct Face {
  bool from_cytosol;
  double eta, n_y, y_this, surfaceScale, factor;
  int ySign, jopp, jthis;
};
Face make_face(const Problem& p, const MCell& c) {
  Face f;
  f.from_cytosol = (c.j == p.jm - 1);
  f.eta = f.from_cytosol ? 1.0 : 0.0;
  f.n_y = f.from_cytosol ? 1.0 : -1.0;
  f.ySign = f.from_cytosol ? -1 : 1;
  f.y_this = f.from_cytosol ? p.y[p.jm] : p.y[p.jm + 1];
  f.jthis = f.from_cytosol ? p.jm : p.jm + 1;
  f.jopp = f.from_cytosol ? p.jm + 1 : p.jm;
  const double y_es = p.y[p.jm + 1];
  const double my_surface = (con_pi * (f.y_this + f.y_this)) * c.hx;
  const double ext_surface = (con_pi * (y_es + y_es)) * c.hx;
  f.surfaceScale = ext_surface / my_surface;
  f.factor = 0.5 * ((2 * con_pi * f.y_this) * c.hx);
  return f;
}
inline double face_centre(const Problem& p, const double* u, int i, int jrow, int comp) {
  return 0.5 * u[4 * p.vid(i, jrow) + comp] + 0.5 * u[4 * p.vid(i + 1, jrow) + comp];
}

// Trans-membrane flux of species j at one quadrature point as MoriNernstPlanckParameters::j assembles it (before the
// surface scaling): channel flux (membranefluxgridfunction_implicit.hh:205-282, x n_y) + capacitive "Mori" flux
// (charge_layer_mori_implicit.hh:190-245, x n_y). uPot / uCon: this side at the point; pot2 / con2: face-centre values
// of the opposite side; potJumpOld: oriented jump of the previous time step. *d_duPot receives d(flux)/d(uPot).
double interface_flux(const Problem& p, const MCell& c, const Face& f, int j, double uPot, double uConj, double pot2,
                      double con2j, double potJumpOld, double* d_duPot) {
  const Config& g = p.cfg;
  static const int species_of[NCHAN] = {Na, K, Na, K};
  const double C1 = (con_k * g.temperature) / con_e;
  const int valence = g.valence[j];
  double potJump = pot2;
  potJump -= uPot;
  potJump *= f.ySign;
  const double dJump = -(double)f.ySign;                  // d potJump / d uPot
  const bool equil = g.do_equilibration && p.time < g.t_equilibrium;
  const double conRatio = f.ySign > 0 ? con2j / uConj : uConj / con2j;
  const double reversal_potential = -std::log(conRatio) * (C1 / valence);
  double total = 0.0, dtotal = 0.0;
  for (int k = 0; k < NCHAN; ++k) {
    if (species_of[k] != j) continue;
    const bool leak = k < 2;
    double conductance = (equil && !leak) ? 0.0 : eff_conductance_new(p, k, c.i);
    if (!g.membrane_active) conductance = 0.0;
    double C = (10 * conductance) / (con_e * valence * con_mol);
    if (equil) C *= g.dummy_value;
    C *= g.time_scale / g.length_scale;
    double diffTerm = C * reversal_potential;
    if (std::fabs(conRatio - 1) < 1e-8 || std::isnan(conRatio)) diffTerm = 0.0;
    const double driftTerm = C * C1 * potJump;
    double flux = driftTerm;
    flux -= diffTerm;
    total += flux;
    dtotal += C * C1 * dJump;
  }
  double y = total * f.n_y, dy = dtotal * f.n_y;
  // capacitive flux of the charge layer
  if (p.mori_initialized || p.mori_partially_initialized) {
    const double real_dt = p.dt * g.time_scale;
    const double dMemb = g.dmemb * g.length_scale;
    const double C_m = con_eps0 * p.eps_memb[c.i] / dMemb;
    const double mV = C1 * 1.0e3;                                        // convertTo_mV
    const double membPotOld = 1e-3 * (potJumpOld * mV);
    const double membPotNew = 1e-3 * (potJump * mV);
    const std::vector<double>* go = f.from_cytosol ? p.mori_gamma_int_old : p.mori_gamma_ext_old;
    const std::vector<double>* gn = f.from_cytosol ? p.mori_gamma_int_new : p.mori_gamma_ext_new;
    const double sigma_old = go[j][c.i] * C_m * membPotOld;
    const double sigma_new = gn[j][c.i] * C_m * membPotNew;
    const double surfaceChargeScale = (g.time_scale / g.length_scale) / (con_e * con_mol);
    const double moriFlux = (surfaceChargeScale / valence) * (sigma_new - sigma_old) / real_dt;
    const double dmori = (surfaceChargeScale / valence) * (gn[j][c.i] * C_m * (1e-3 * (dJump * mV))) / real_dt;
    y += moriFlux * f.n_y;
    dy += dmori * f.n_y;
  }
  if (d_duPot) *d_duPot = dy;
  return y;
}

}  // namespace

// ---- charge layer state (ChargeLayerMoriImplicit::updateState / init / timeStep / acceptTimeStep) -----------------
void mori_update_state(Problem& p, const double* u) {
  const int nx = p.nx;
  if (!p.mori_initialized && p.mori_partially_initialized) return;
  if (p.mori_gamma_int_old[0].empty())
    for (int j = 0; j < NSPEC; ++j) {
      p.mori_gamma_int_old[j].assign(nx, 0.0); p.mori_gamma_int_new[j].assign(nx, 0.0);
      p.mori_gamma_ext_old[j].assign(nx, 0.0); p.mori_gamma_ext_new[j].assign(nx, 0.0);
    }
  const Config& g = p.cfg;
  const bool do_init = !p.mori_initialized;
  if (do_init && g.do_equilibration && !(std::fabs(p.time - g.t_equilibrium) < 1e-6 || p.time > g.t_equilibrium)) return;
  const double tau = 1e-9;
  const double dt_ = p.dt * g.time_scale;
  for (int i = 0; i < nx; ++i) {
    double conC[NSPEC], conE[NSPEC], sum_ext = 0.0, sum_int = 0.0;
    for (int j = 0; j < NSPEC; ++j) {
      conC[j] = face_centre(p, u, i, p.jm, j);
      conE[j] = face_centre(p, u, i, p.jm + 1, j);
      const double z2 = g.valence[j] * g.valence[j];
      sum_ext += z2 * conE[j];
      sum_int += z2 * conC[j];
    }
    for (int j = 0; j < NSPEC; ++j) {
      const double z2 = g.valence[j] * g.valence[j];
      const double snake_ext = z2 * conE[j] / sum_ext, snake_int = z2 * conC[j] / sum_int;
      if (do_init) {
        p.mori_gamma_ext_old[j][i] = snake_ext; p.mori_gamma_int_old[j][i] = snake_int;
        p.mori_gamma_ext_new[j][i] = snake_ext; p.mori_gamma_int_new[j][i] = snake_int;
      } else {
        p.mori_gamma_ext_new[j][i] = (tau * p.mori_gamma_ext_old[j][i] + dt_ * snake_ext) / (tau + dt_);
        p.mori_gamma_int_new[j][i] = (tau * p.mori_gamma_int_old[j][i] + dt_ * snake_int) / (tau + dt_);
      }
    }
  }
  if (do_init) p.mori_partially_initialized = true;
}

void mori_accept_state(Problem& p) {
  if (p.mori_partially_initialized) { p.mori_initialized = true; p.mori_partially_initialized = false; }
  for (int j = 0; j < NSPEC; ++j) {
    p.mori_gamma_ext_old[j] = p.mori_gamma_ext_new[j];
    p.mori_gamma_int_old[j] = p.mori_gamma_int_new[j];
  }
}

// ---- potential sub-problem --------------------------------------------------------------------------------------
// r = residual of the potential rows at u (concentration rows 0); A (optional) = its Jacobian w.r.t. the potential
// unknowns with identity on the concentration rows. ulast supplies the concentrations, uold the previous potential.
void mori_potential_assemble(const Problem& p, const double* u, const double* ulast, double* r, BCRS* A) {
  const Config& g = p.cfg;
  const size_t n = 4 * (size_t)p.nv;
  std::memset(r, 0, n * sizeof(double));
  if (A) std::fill(A->val.begin(), A->val.end(), 0.0);
  const double S = g.pot_residual_scale;
  const double cond_unit = g.length_scale * con_e * con_mol;          // lengthScale * e * N_A
  const double jscale = p.PC * g.time_scale / (g.length_scale * g.length_scale);
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      if (p.cell_sub[p.cid(i, j)] == MEMBRANE) continue;                // useMembraneContributions = false
      const MCell c = make_cell(p, i, j);
      double xl[16], xc[16], xo[16], rl[4] = {0, 0, 0, 0}, ml[16] = {0};
      gather(c, u, xl);
      gather(c, ulast, xc);
      gather(c, p.uold.data(), xo);
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          double phi[4], dxi[4], deta[4], gx[4], gy[4];
          q1(GP[a], GP[b], phi, dxi, deta);
          for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
          double uCon[NSPEC], gcx[NSPEC], gcy[NSPEC];
          for (int s = 0; s < NSPEC; ++s) {
            double v = 0, sx = 0, sy = 0;
            for (int k = 0; k < 4; ++k) { v += xc[4 * s + k] * phi[k]; sx += xc[4 * s + k] * gx[k]; sy += xc[4 * s + k] * gy[k]; }
            uCon[s] = v; gcx[s] = sx; gcy[s] = sy;
          }
          double cond = 0.0;                                              // MoriPoissonParameters::A
          for (int s = 0; s < NSPEC; ++s) cond += cond_unit * g.valence[s] * g.valence[s] * p.D[s] * uCon[s];
          double bgx = 0.0, bgy = 0.0;                                    // bGrad: sum_j LS e N_A z_j (D_j grad c_j)
          for (int s = 0; s < NSPEC; ++s) {
            double tx = p.D[s] * gcx[s], ty = p.D[s] * gcy[s];
            tx *= cond_unit * g.valence[s]; ty *= cond_unit * g.valence[s];
            bgx += tx; bgy += ty;
          }
          double gpx = 0, gpy = 0;
          for (int k = 0; k < 4; ++k) { gpx += xl[12 + k] * gx[k]; gpy += xl[12 + k] * gy[k]; }
          const double Ax = cond * gpx, Ay = cond * gpy;
          const double yq = c.y0 + GP[b] * c.hy;
          const double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
          for (int k = 0; k < 4; ++k) {
            rl[k] += ((Ax * gx[k] + Ay * gy[k]) + (bgx * gx[k] + bgy * gy[k])) * factor * c.vs[k] * S;
            if (A)
              for (int m = 0; m < 4; ++m)
                ml[k * 4 + m] += ((cond * gx[m]) * gx[k] + (cond * gy[m]) * gy[k]) * factor * c.vs[k] * S;
          }
        }
      if (is_interface_cell(p, j) && g.mori_membrane_neumann) {
        const Face f = make_face(p, c);
        const double pot2 = face_centre(p, u, i, f.jopp, POT);
        const double potJumpOld = (face_centre(p, p.uold.data(), i, f.jopp, POT) -
                                   face_centre(p, p.uold.data(), i, f.jthis, POT)) * f.ySign;
        for (int s = 0; s < 2; ++s) {
          double phi[4], dxi[4], deta[4];
          q1(GP[s], f.eta, phi, dxi, deta);
          double uPot = 0.0;
          for (int k = 0; k < 4; ++k) uPot += xl[12 + k] * phi[k];
          double jPot = 0.0, djPot = 0.0;
          for (int sp = 0; sp < NSPEC; ++sp) {
            double uc = 0.0;                                           // this side: getSolutionConOld (:346-350)
            for (int k = 0; k < 4; ++k) uc += xo[4 * sp + k] * phi[k];
            double d;
            double jc = interface_flux(p, c, f, sp, uPot, uc, pot2, face_centre(p, ulast, i, f.jopp, sp), potJumpOld, &d);
            jc *= f.surfaceScale; d *= f.surfaceScale;
            jPot += g.valence[sp] * jc;
            djPot += g.valence[sp] * d;
          }
          jPot *= jscale; djPot *= jscale;
          for (int k = 0; k < 4; ++k) {
            rl[k] += jPot * phi[k] * f.factor * c.vs[k] * S;
            if (A)
              for (int m = 0; m < 4; ++m) ml[k * 4 + m] += (djPot * phi[m]) * phi[k] * f.factor * c.vs[k] * S;
          }
        }
      }
      for (int k = 0; k < 4; ++k) {
        if ((p.dirichlet[c.v[k]] >> POT) & 1) continue;
        r[4 * c.v[k] + POT] += rl[k];
        if (A)
          for (int m = 0; m < 4; ++m) {
            if (ml[k * 4 + m] == 0.0) continue;
            A->block(A->find(c.v[k], c.v[m]))[POT * 4 + POT] += ml[k * 4 + m];
          }
      }
    }
  if (A)
    for (int v = 0; v < p.nv; ++v) {
      double* d = A->block(A->find(v, v));
      for (int comp = 0; comp < 3; ++comp) d[comp * 4 + comp] = 1.0;
      if ((p.dirichlet[v] >> POT) & 1) d[POT * 4 + POT] = 1.0;
    }
}

// ---- concentration sub-problem (implicit Euler: r = M(c) - M(c_old) + dt A_con(c)) -------------------------------
// u: current concentrations (unknowns); ulast: concentrations / potential of the last iteration (drift term, opposite
// side of the membrane, potential of the membrane fluxes); p.uold: previous time step (mass term, this side's
// concentration in the membrane flux, old potential jump of the capacitive flux).
void mori_concentration_assemble(const Problem& p, const double* u, const double* ulast, double* r, BCRS* A) {
  const Config& g = p.cfg;
  const size_t n = 4 * (size_t)p.nv;
  std::memset(r, 0, n * sizeof(double));
  if (A) std::fill(A->val.begin(), A->val.end(), 0.0);
  const double* uold = p.uold.data();
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      if (p.cell_sub[p.cid(i, j)] == MEMBRANE) continue;
      const MCell c = make_cell(p, i, j);
      double xl[16], xc[16], xo[16], rl[12] = {0}, ml[48] = {0};   // ml[species][k][m]
      gather(c, u, xl);
      gather(c, ulast, xc);
      gather(c, uold, xo);
      double fsrc[NSPEC];
      for (int s = 0; s < NSPEC; ++s) fsrc[s] = source_f(p, c, s);
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          double phi[4], dxi[4], deta[4], gx[4], gy[4];
          q1(GP[a], GP[b], phi, dxi, deta);
          for (int k = 0; k < 4; ++k) { gx[k] = c.ihx * dxi[k]; gy[k] = c.ihy * deta[k]; }
          double gpx = 0, gpy = 0;                                   // OLD potential gradient (operator split)
          for (int k = 0; k < 4; ++k) { gpx += xc[12 + k] * gx[k]; gpy += xc[12 + k] * gy[k]; }
          const double yq = c.y0 + GP[b] * c.hy;
          const double factor = 0.25 * ((2 * con_pi * yq) * (c.hx * c.hy));
          for (int s = 0; s < NSPEC; ++s) {
            double uC = 0, uCl = 0, uCo = 0, sx = 0, sy = 0;
            for (int k = 0; k < 4; ++k) {
              uC += xl[4 * s + k] * phi[k]; uCl += xc[4 * s + k] * phi[k]; uCo += xo[4 * s + k] * phi[k];
              sx += xl[4 * s + k] * gx[k]; sy += xl[4 * s + k] * gy[k];
            }
            const double D = p.D[s];
            const double Acx = D * sx, Acy = D * sy;
            const double bx = -D * g.valence[s] * gpx, by = -D * g.valence[s] * gpy;
            for (int k = 0; k < 4; ++k) {
              const double spatial = (Acx * gx[k] + Acy * gy[k]) - uCl * (bx * gx[k] + by * gy[k]) + (0.0 * uC - fsrc[s]) * phi[k];
              rl[4 * s + k] += p.dt * (spatial * factor * c.vs[k]);
              rl[4 * s + k] += (uC * phi[k] * factor * c.vs[k]);        // M(c)
              rl[4 * s + k] += -(uCo * phi[k] * factor * c.vs[k]);      // - M(c_old)
              if (A)
                for (int m = 0; m < 4; ++m) {
                  ml[(4 * s + k) * 4 + m] += p.dt * (((D * gx[m]) * gx[k] + (D * gy[m]) * gy[k]) * factor * c.vs[k]);
                  ml[(4 * s + k) * 4 + m] += (phi[m] * phi[k] * factor * c.vs[k]);
                }
            }
          }
        }
      if (is_interface_cell(p, j)) {
        const Face f = make_face(p, c);
        const double pot2 = face_centre(p, ulast, i, f.jopp, POT);
        const double potJumpOld = (face_centre(p, uold, i, f.jopp, POT) - face_centre(p, uold, i, f.jthis, POT)) * f.ySign;
        for (int s = 0; s < 2; ++s) {
          double phi[4], dxi[4], deta[4];
          q1(GP[s], f.eta, phi, dxi, deta);
          double uPot = 0.0;                                           // "new" potential of this phase = ulast
          for (int k = 0; k < 4; ++k) uPot += xc[12 + k] * phi[k];
          for (int sp = 0; sp < NSPEC; ++sp) {
            double uc = 0.0;                                           // this side: OLD concentration (getSolutionConOld)
            for (int k = 0; k < 4; ++k) uc += xo[4 * sp + k] * phi[k];
            double jc = interface_flux(p, c, f, sp, uPot, uc, pot2, face_centre(p, ulast, i, f.jopp, sp), potJumpOld, nullptr);
            jc *= f.surfaceScale;
            for (int k = 0; k < 4; ++k) rl[4 * sp + k] += p.dt * (jc * phi[k] * f.factor * c.vs[k]);
          }
        }
      }
      for (int s = 0; s < NSPEC; ++s)
        for (int k = 0; k < 4; ++k) {
          if ((p.dirichlet[c.v[k]] >> s) & 1) continue;
          r[4 * c.v[k] + s] += rl[4 * s + k];
          if (A)
            for (int m = 0; m < 4; ++m) A->block(A->find(c.v[k], c.v[m]))[s * 4 + s] += ml[(4 * s + k) * 4 + m];
        }
    }
  if (A)
    for (int v = 0; v < p.nv; ++v) {
      double* d = A->block(A->find(v, v));
      d[POT * 4 + POT] = 1.0;
      for (int s = 0; s < NSPEC; ++s)
        if ((p.dirichlet[v] >> s) & 1) d[s * 4 + s] = 1.0;
    }
}

// ---- one time step -----------------------------------------------------------------------------------------------
// u: in = solution of the previous step (= uold = ulast: no restart), out = new solution. The caller has set time / dt,
// uold (set_uold) and updated the channel states (update_state) as the time loop does.
void mori_step(Problem& p, double* u, const NewtonOpts& o, MoriResult& res) {
  const size_t n = 4 * (size_t)p.nv;
  res = MoriResult();
  if (!p.mori_initialized && !p.mori_partially_initialized) {   // [init] block of the simulation (:129-142)
    mori_update_state(p, u);
    mori_accept_state(p);
  }
  std::vector<double> ulast(u, u + n), r(n), z(n);
  BCRS A;
  build_pattern(p, A);
  ILU M;
  auto solve = [&](double& first_defect, double& defect, int& its) {
    double d0 = 0.0;
    for (size_t k = 0; k < n; ++k) d0 += r[k] * r[k];
    d0 = std::sqrt(d0);
    first_defect = d0;
    if (!std::isfinite(d0)) throw std::runtime_error("Initial residual is NaN!");
    std::fill(z.begin(), z.end(), 0.0);
    const double red = std::max(o.reduction, o.abs_limit / d0);   // ax1_linearsubproblem.hh: max(reduction, min_defect / defect)
    LinResult lr;
    if (d0 > 0.0) {
      ilu_factor(p, A, o.ilu_fill, o.slab_w, M);
      bicgstab(p, A, M, z.data(), r.data(), red, o.lin_maxit, nullptr, lr);
      if (!lr.converged) throw std::runtime_error("linear sub-problem solver did not converge");
    }
    its = lr.iterations;
    defect = d0 * lr.reduction;
    for (size_t k = 0; k < n; ++k) u[k] -= z[k];
  };
  // (1) potential
  mori_potential_assemble(p, u, ulast.data(), r.data(), &A);
  solve(res.pot_first_defect, res.pot_defect, res.pot_iterations);
  // (2) charge layer
  mori_update_state(p, ulast.data());
  // (3) concentrations (potential and concentrations of the membrane / drift terms: ulast)
  mori_concentration_assemble(p, u, ulast.data(), r.data(), &A);
  solve(res.con_first_defect, res.con_defect, res.con_iterations);
  res.status = 0;
}

}  // namespace ax1o
