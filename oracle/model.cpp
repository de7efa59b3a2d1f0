// oracle/model.cpp -- TEST INFRASTRUCTURE ONLY.
// Restatement of the host-side "Physics" precomputations and the channel model of dune-ax1:
//   subdomain tags            dune/ax1/common/ax1_gridgenerator.hh:776-860
//   node volumes              dune/ax1/acme2_cyl/common/acme2_cyl_physics.hh:2286-2397
//   Dirichlet constraints     configurations/default/default_{nernst_planck,poisson}_parameters.hh
//                             (bctype, :219-367 / :161-292) + ax1_parallelconstraintshelper.hh:36-41
//   sparsity pattern          FullVolumePattern on Q1 vertices
//   channels                  dune/ax1/channels/channel.hh, channelset.hh
//   gating update             dune/ax1/common/membranefluxgridfunction_implicit.hh:429-546
#include "oracle.hpp"
#include <algorithm>
#include <stdexcept>

namespace ax1o {

static double cyl_volume(double x0, double x1, double y0, double y1) {
  // Acme2CylGeometry::volume, acme2_cyl_geometrywrapper.hh:64-72
  const double factor = con_pi * (y1 + y0);
  return factor * ((x1 - x0) * (y1 - y0));
}

void setup_problem(Problem& p) {
  const Config& c = p.cfg;
  p.nvx = p.nx + 1; p.nvy = p.ny + 1; p.nv = p.nvx * p.nvy;
  if ((int)p.x.size() != p.nvx || (int)p.y.size() != p.nvy) throw std::runtime_error("bad grid");
  // electrolyte.hh:61-67, default_config.hh:79-91
  p.PC = con_e * con_e * con_mol * c.length_scale * c.length_scale / (con_eps0 * con_k * c.temperature);
  p.kT_e = (con_k * c.temperature) / con_e;
  for (int j = 0; j < NSPEC; ++j) p.D[j] = c.diff_water[j] * c.time_scale / (c.length_scale * c.length_scale);

  // --- subdomain tags (single membrane), ax1_gridgenerator.hh:806-828
  p.cell_sub.assign((size_t)p.nx * p.ny, CYTOSOL);
  p.jm = -1;
  for (int j = 0; j < p.ny; ++j) {
    double yc = 0.5 * (p.y[j] + p.y[j + 1]);
    int sd = CYTOSOL;
    if (yc > c.ymemb && yc < c.ymemb + c.dmemb) sd = MEMBRANE;
    if (yc > c.ymemb + c.dmemb) sd = ES;
    if (sd == MEMBRANE) {
      if (p.jm >= 0) throw std::runtime_error("oracle supports exactly one membrane cell layer");
      p.jm = j;
    }
    for (int i = 0; i < p.nx; ++i) p.cell_sub[p.cid(i, j)] = (uint8_t)sd;
  }
  if (p.jm < 1 || p.jm > p.ny - 2) throw std::runtime_error("membrane row not found");
  if ((int)p.eps_memb.size() != p.nx) p.eps_memb.assign(p.nx, 2.0);

  // --- node volumes: literal cell loop, physics.hh:2286-2397 (order dependent by construction)
  p.node_volume.assign(p.nv, 0.0);
  {
    std::vector<char> exists(p.nv, 0);
    double reference_volume = -1.0;
    double cellWidth = c.ymemb;  // one membrane: yMemb[0]
    double yStart = c.volume_scaling_threshold >= 0 ? c.volume_scaling_threshold
                                                    : c.ymemb + c.dmemb + cellWidth;
    for (int j = 0; j < p.ny; ++j)
      for (int i = 0; i < p.nx; ++i) {
        double yc = 0.5 * (p.y[j] + p.y[j + 1]);
        int corners[4] = {p.vid(i, j), p.vid(i + 1, j), p.vid(i, j + 1), p.vid(i + 1, j + 1)};
        for (int k = 0; k < 4; ++k) {
          int v = corners[k];
          if (yc > yStart && c.do_volume_scaling) {
            double volume = cyl_volume(p.x[i], p.x[i + 1], p.y[j], p.y[j + 1]);
            if (reference_volume < 0) reference_volume = volume;
            if (!exists[v]) p.node_volume[v] = volume / reference_volume;
            else p.node_volume[v] = std::min(p.node_volume[v], volume / reference_volume);
          } else {
            p.node_volume[v] = 1.0;
          }
          exists[v] = 1;
        }
      }
  }

  // --- Dirichlet constraints from boundary-face bctype at face centres
  p.dirichlet.assign(p.nv, 0);
  auto side_class = [&](double yc) {  // 0 cytosol, 1 debye, 2 extracellular (SURVEY A.4 note)
    double w = c.debye_layer_width;
    if (yc - 1e-6 < c.ymemb - w) return 0;
    if (yc + 1e-6 > c.ymemb - w && yc - 1e-6 < c.ymemb + c.dmemb + w) return 1;
    return 2;
  };
  auto mark = [&](int va, int vb, bool con, bool pot) {
    uint8_t m = (con ? 0x7 : 0) | (pot ? 0x8 : 0);
    p.dirichlet[va] |= m; p.dirichlet[vb] |= m;
  };
  auto face_bc = [&](double xc, double yc, int which /*0 con, 1 pot*/) -> bool {
    bool top = yc + 1e-6 > c.ymax, bottom = yc - 1e-6 < c.ymin;
    bool left = xc - 1e-6 < c.xmin_global, right = xc + 1e-6 > c.xmax_global;
    int sc = (left || right) ? side_class(yc) : 2;
    if (top) return c.top_dirichlet[which];
    if (bottom) return c.bottom_dirichlet[which];
    if (left) return sc == 0 ? c.left_cytosol_dirichlet[which] : sc == 1 ? c.left_debye_dirichlet[which]
                                                                          : c.left_extra_dirichlet[which];
    if (right) return sc == 0 ? c.right_cytosol_dirichlet[which] : sc == 1 ? c.right_debye_dirichlet[which]
                                                                           : c.right_extra_dirichlet[which];
    throw std::runtime_error("Error checking this boundary's position!");
  };
  for (int j = 0; j < p.ny; ++j)
    for (int i = 0; i < p.nx; ++i) {
      bool memb = p.cell_sub[p.cid(i, j)] == MEMBRANE;  // membrane cells carry pot only
      double xc = 0.5 * (p.x[i] + p.x[i + 1]), yc = 0.5 * (p.y[j] + p.y[j + 1]);
      if (i == 0 && !p.left_proc_boundary)
        mark(p.vid(0, j), p.vid(0, j + 1), !memb && face_bc(p.x[0], yc, 0), face_bc(p.x[0], yc, 1));
      if (i == p.nx - 1 && !p.right_proc_boundary)
        mark(p.vid(p.nx, j), p.vid(p.nx, j + 1), !memb && face_bc(p.x[p.nx], yc, 0), face_bc(p.x[p.nx], yc, 1));
      if (j == 0) mark(p.vid(i, 0), p.vid(i + 1, 0), !memb && face_bc(xc, p.y[0], 0), face_bc(xc, p.y[0], 1));
      if (j == p.ny - 1)
        mark(p.vid(i, p.ny), p.vid(i + 1, p.ny), !memb && face_bc(xc, p.y[p.ny], 0), face_bc(xc, p.y[p.ny], 1));
    }
  // processor-boundary vertices of overlap cells: all components (ax1_parallelconstraintshelper.hh:36-41)
  for (int j = 0; j < p.nvy; ++j) {
    if (p.left_proc_boundary) p.dirichlet[p.vid(0, j)] = 0xF;
    if (p.right_proc_boundary) p.dirichlet[p.vid(p.nx, j)] = 0xF;
  }
  // disjoint ownership (lowest rank owns shared border vertices; front vertices are never owned)
  p.own_lo = p.left_proc_boundary ? 2 : 0;
  p.own_hi = p.right_proc_boundary ? p.nvx - 2 : p.nvx - 1;

  channels_static_init(p);
}

void build_pattern(const Problem& p, BCRS& A) {
  A.n = p.nv;
  A.rowptr.assign(p.nv + 1, 0);
  A.col.clear();
  for (int j = 0; j < p.nvy; ++j)
    for (int i = 0; i < p.nvx; ++i) {
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          int ii = i + di, jj = j + dj;
          if (ii < 0 || ii >= p.nvx || jj < 0 || jj >= p.nvy) continue;
          A.col.push_back(p.vid(ii, jj));
        }
      A.rowptr[p.vid(i, j) + 1] = (int)A.col.size();
    }
  A.val.assign(A.col.size() * 16, 0.0);
}

// ---------------------------------------------------------------------------------------------
// channels
static inline double v_trap(double x, double y) {  // channel.hh:185-194
  if (std::fabs(x / y) < 1e-8) return y * (1 - x / y / 2);
  return x / (std::exp(x / y) - 1);
}
// gate 0 = Nav m, 1 = Nav h, 2 = Kv n; v_ = v - v_rest [mV]; temp_fac = 3^0 = 1
static inline double alpha_gate(int g, double v_) {
  switch (g) {
    case 0: return 1.0 * (0.1 * v_trap((25 - v_), 10));     // channel.hh:606-618
    case 1: return 1.0 * (0.07 * std::exp(-v_ / 20));
    default: return 1.0 * (0.01 * v_trap((10 - v_), 10));   // :667-676
  }
}
static inline double beta_gate(int g, double v_) {
  switch (g) {
    case 0: return 1.0 * (4 * std::exp(-v_ / 18));           // channel.hh:624-642
    case 1: return 1.0 * (1. / (std::exp((30 - v_) / 10) + 1));
    default: return 1.0 * (0.125 * std::exp(-v_ / 80));      // :679-687
  }
}

void channels_static_init(Problem& p) {
  for (int k = 0; k < 3; ++k)
    if ((int)p.gbar[k].size() != p.nx) p.gbar[k].assign(p.nx, 0.0);
  for (int k = 0; k < 2; ++k) p.ratio[k].assign(p.nx, k == 0 ? p.cfg.leak_ratio_na : p.cfg.leak_ratio_k);
  for (int g = 0; g < 3; ++g) {
    // VoltageGatedChannel::init(): steady state at v = v_rest (channel.hh:236-247)
    double a = alpha_gate(g, 0.0), b = beta_gate(g, 0.0);
    p.gate[g].assign(p.nx, a / (a + b));
    p.gate_new[g] = p.gate[g];
  }
  p.channels_initialized = false;
  p.channels_partially_initialized = false;
}

// new effective conductance of channel k on edge i (channel.hh:356-366, 497-501)
double eff_conductance_new(const Problem& p, int k, int i) {
  switch (k) {
    case CH_NA_LEAK: return p.gbar[0][i] * p.ratio[0][i];
    case CH_K_LEAK: return p.gbar[0][i] * p.ratio[1][i];
    case CH_NAV: {
      double g = p.gbar[1][i];
      g *= std::pow(p.gate_new[0][i], 3);
      g *= std::pow(p.gate_new[1][i], 1);
      return g;
    }
    default: {
      double g = p.gbar[2][i];
      g *= std::pow(p.gate_new[2][i], 4);
      return g;
    }
  }
}

static void recalc_leak(Problem& p, int i) {  // channelset.hh:234-311
  static const int species_of[NCHAN] = {Na, K, Na, K};
  for (int j = 0; j < NSPEC; ++j)
    for (int k = 0; k < 2; ++k) {  // leak channels only
      if (species_of[k] != j) continue;
      double total = p.gbar[0][i];
      if (total > 0.0) {
        double sumThis = 0.0, sumOther = 0.0;
        for (int l = CH_NAV; l <= CH_KV; ++l) {
          if (species_of[l] != j) sumOther += eff_conductance_new(p, l, i);
          else sumThis += eff_conductance_new(p, l, i);
        }
        double ratio = p.ratio[k][i];
        double leak = (ratio / (1. - ratio)) * (total + sumOther) - sumThis;
        leak /= 1. + (ratio / (1. - ratio));
        double newRatio = leak / total;
        if (newRatio < 0.0 || newRatio > 1.0) throw std::runtime_error("leak ratio out of [0,1]");
        p.ratio[k][i] = newRatio;
      }
    }
}

void membrane_potential_mV(const Problem& p, const double* u, double* vm) {
  // physics.hh:562-603 + convertTo_mV :1062-1067; face centres of the membrane cell
  for (int i = 0; i < p.nx; ++i) {
    double in  = 0.5 * u[4 * p.vid(i, p.jm) + POT] + 0.5 * u[4 * p.vid(i + 1, p.jm) + POT];
    double out = 0.5 * u[4 * p.vid(i, p.jm + 1) + POT] + 0.5 * u[4 * p.vid(i + 1, p.jm + 1) + POT];
    double membPot = in;
    membPot -= out;
    vm[i] = membPot * con_k * p.cfg.temperature / con_e * 1.0e3;
  }
}

void update_state(Problem& p, const double* u) {  // membranefluxgridfunction_implicit.hh:429-501
  const Config& c = p.cfg;
  if (!c.membrane_active) return;
  if (!p.channels_initialized && p.channels_partially_initialized) return;
  std::vector<double> vm(p.nx);
  membrane_potential_mV(p, u, vm.data());
  for (int i = 0; i < p.nx; ++i) {
    if (!p.channels_initialized) {
      if (!c.do_equilibration || (std::fabs(p.time - c.t_equilibrium) < 1e-6 || p.time > c.t_equilibrium)) {
        double v = c.automatic_channel_init ? vm[i] : c.channel_resting_potential;
        // DynamicChannelSet::initChannels (channelset.hh:186-228): v_rest of the shared channel
        // objects is overwritten per edge, gates go to steady state at v - v_rest = 0
        p.v_rest = v;
        for (int g = 0; g < 3; ++g) {
          double a = alpha_gate(g, v - p.v_rest), b = beta_gate(g, v - p.v_rest);
          p.gate[g][i] = a / (a + b);
          p.gate_new[g][i] = p.gate[g][i];
        }
        recalc_leak(p, i);
        p.channels_partially_initialized = true;
      }
    } else {
      double real_dt = p.dt * c.time_scale;
      double dt_ms = real_dt / 1e-3;  // DynamicChannelSet::TIME_SCALE
      for (int g = 0; g < 3; ++g) {   // step_be, channel.hh:394-402
        double v_ = vm[i] - p.v_rest;
        double a = alpha_gate(g, v_), b = beta_gate(g, v_);
        p.gate_new[g][i] = (dt_ms * a + p.gate[g][i]) / (1 + dt_ms * (a + b));
      }
    }
  }
}

void accept_state(Problem& p) {  // membranefluxgridfunction_implicit.hh:524-546
  if (!p.cfg.membrane_active) return;
  if (p.channels_partially_initialized) {
    p.channels_initialized = true;
    p.channels_partially_initialized = false;
  }
  for (int g = 0; g < 3; ++g) p.gate[g] = p.gate_new[g];
}

}  // namespace ax1o
