// host_model.cpp -- host-side model of the reference's "Physics" precomputations, part of the
// product (libax1gpu.so). It lets the device path be driven without DUNE: radial grid vector,
// subdomain tags, node volumes, Dirichlet masks and the z-slab partition. When the library is
// dropped into acme2_cyl_par these arrays come from the reference's own Physics object instead
//  and this file is not needed.
//
//   radial grid vector   dune/ax1/common/ax1_gridvector.hh, ax1_gridgenerator.hh:74-149,232-277,357-450
//   subdomain tags       dune/ax1/common/ax1_gridgenerator.hh:806-828
//   node volumes         dune/ax1/acme2_cyl/common/acme2_cyl_physics.hh:2286-2397
//   boundary types       configurations/default/default_{nernst_planck,poisson}_parameters.hh (bctype)
//   overlap constraints  dune/ax1/common/ax1_parallelconstraintshelper.hh:36-41
//   slab partition       dune/ax1/common/ax1_gridgenerator.hh:601-631, ax1_yaspgrid_loadbalancer.hh:17-38
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/ax1gpu.h"

namespace ax1 { void set_error(const std::string& s); }

namespace {

const double kPi = 3.1415926535897932384626433;

// ---- grid vector pieces -------------------------------------------------------------------
struct Axis {
  std::vector<double> p;
  double last() const { return p.back(); }
  void uniform_steps(int n, double h) { for (int k = 0; k < n; ++k) p.push_back(last() + h); }
  void uniform_to(int n, double end) {
    const double h = (end - last()) / n;
    for (int k = 0; k + 1 < n; ++k) p.push_back(last() + h);
    p.push_back(end);
  }
  // root of  -q^n + m q - m + 1 = 0  by Newton (ratio of a geometric progression of n steps
  // starting with step h that exactly spans [s,e])
  static double ratio(int n, double s, double e, double h) {
    const double m = (e - s) / h;
    auto iterate = [&](double from, double prev) {
      double cur = from;
      while (std::fabs(cur - prev) > 1E-8) {
        prev = cur;
        cur = prev - (-std::pow(prev, n) + m * prev - m + 1) / (-n * std::pow(prev, n - 1) + m);
      }
      return cur;
    };
    double q = iterate(e - s, 0.0);
    if (std::fabs(q - 1) < 1E-6) q = iterate(0.0, e - s);
    return q;
  }
  void geometric_first(int n, double h0, double end) {  // first step fixed
    const double q = ratio(n, last(), end, h0);
    double h = h0;
    for (int k = 0; k + 1 < n; ++k) { p.push_back(last() + h); h *= q; }
    p.push_back(end);
  }
  void geometric_last(int n, double hend, double end) {  // last step fixed
    const double q = ratio(n, last(), end, hend);
    double h = hend * std::pow(q, n - 1);
    for (int k = 0; k + 1 < n; ++k) { p.push_back(last() + h); h /= q; }
    p.push_back(end);
  }
  void geometric(double h0, double hend, double end, bool fix_last) {
    const double q = (end - last() - h0) / (end - last() - hend);
    const int n = (int)(std::log(hend / h0) / std::log(q)) + 1;
    if (fix_last) geometric_last(n, hend, end); else geometric_first(n, h0, end);
  }
};

}  // namespace

extern "C" {

int ax1host_generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell, double dy_cell_min,
                       int n_dy_cell_min, double dy, double dy_min, int n_dy_min, double* out, int cap) {
  const double q = 0.7;  // qTransition
  Axis y;
  y.p.push_back(ymin);
  if (ymemb > y.last() + 1e-6) {  // cytosol: equidistant, then geometric refinement towards the membrane
    const int nt = (int)std::ceil(std::fabs(std::log(dy_cell_min / dy_cell) / std::log(q)));
    const double range = std::max(dy_cell, dy_cell_min) * (std::pow(q, nt + 1) - 1) / (q - 1);
    const double dist = n_dy_cell_min * dy_cell_min + range;
    if (ymemb - dist < y.last()) { ax1::set_error("cytosol too thin for the requested dy_cell_min"); return -1; }
    const int neq = (int)std::floor((ymemb - dist - y.last()) / dy_cell);
    if (neq > 0) y.uniform_steps(neq, dy_cell);
    y.geometric(dy_cell, dy_cell_min, ymemb - n_dy_cell_min * dy_cell_min, false);
    y.uniform_steps(n_dy_cell_min, dy_cell_min);
  }
  if (ymemb + dmemb > y.last() + 1e-6 && dmemb > 0) y.uniform_steps(1, dmemb);  // one membrane cell layer
  {  // extracellular: fine layer, geometric coarsening, equidistant far field
    const double y0 = y.last();
    y.uniform_steps(n_dy_min, dy_min);
    const int nt = (int)std::ceil(std::fabs(std::log(dy_min / dy) / std::log(q)));
    const double range = dy * (std::pow(q, nt + 1) - 1) / (q - 1);
    const int neq = (int)std::floor((ymax - y0 - (n_dy_min * dy_min + range)) / dy);
    y.geometric(dy_min, dy, ymax - (neq * dy), true);
    if (neq > 0) y.uniform_to(neq, ymax);
  }
  if (std::fabs(y.last() - ymax) > 1e-6) { ax1::set_error("radial grid does not end at ymax"); return -1; }
  if (out) for (int k = 0; k < (int)y.p.size() && k < cap; ++k) out[k] = y.p[k];
  return (int)y.p.size();
}

int ax1host_physics_maps(const ax1gpu_grid* grid, double ymemb, double dmemb, double xmin_global, double xmax_global,
                         double ymin, double ymax, int do_volume_scaling, double volume_scaling_threshold,
                         const ax1host_bc* bc, unsigned char* cell_subdomain, double* node_volume,
                         unsigned char* dirichlet_mask) {
  if (!grid || !bc) { ax1::set_error("null argument"); return AX1GPU_EINVAL; }
  const int nx = grid->nx, ny = grid->ny, nvx = nx + 1, nvy = ny + 1;
  const double* x = grid->x;
  const double* y = grid->y;
  // row classification by cell-centre height
  std::vector<unsigned char> row_sd(ny);
  for (int j = 0; j < ny; ++j) {
    const double yc = 0.5 * (y[j] + y[j + 1]);
    row_sd[j] = (yc > ymemb && yc < ymemb + dmemb) ? 2 : (yc > ymemb + dmemb ? 1 : 0);
  }
  if (cell_subdomain)
    for (int j = 0; j < ny; ++j) std::memset(cell_subdomain + (size_t)j * nx, row_sd[j], nx);

  if (node_volume) {
    // Order-dependent by construction in the reference (cells visited x-fastest; min over scaled
    // cells, overwrite with 1 by unscaled ones; reference volume = first scaled cell on this rank).
    // Rows are visited bottom-up, so a vertex row j sees cell rows j-1 then j; within a row the
    // left cell comes before the right one.
    const double threshold = volume_scaling_threshold >= 0 ? volume_scaling_threshold : ymemb + dmemb + ymemb;
    std::vector<char> scaled(ny);
    int jfirst = -1;
    for (int j = 0; j < ny; ++j) {
      scaled[j] = do_volume_scaling && (0.5 * (y[j] + y[j + 1]) > threshold);
      if (scaled[j] && jfirst < 0) jfirst = j;
    }
    auto cellvol = [&](int i, int j) { return (kPi * (y[j + 1] + y[j])) * ((x[i + 1] - x[i]) * (y[j + 1] - y[j])); };
    const double ref = jfirst >= 0 ? cellvol(0, jfirst) : 1.0;
    for (int jv = 0; jv < nvy; ++jv)
      for (int iv = 0; iv < nvx; ++iv) {
        double val = 0.0;
        bool have = false;
        for (int cj = jv - 1; cj <= jv; ++cj) {
          if (cj < 0 || cj >= ny) continue;
          for (int ci = iv - 1; ci <= iv; ++ci) {
            if (ci < 0 || ci >= nx) continue;
            if (scaled[cj]) {
              const double r = cellvol(ci, cj) / ref;
              val = have ? std::min(val, r) : r;
            } else {
              val = 1.0;
            }
            have = true;
          }
        }
        node_volume[(size_t)iv + (size_t)nvx * jv] = val;
      }
  }

  if (dirichlet_mask) {
    std::memset(dirichlet_mask, 0, (size_t)nvx * nvy);
    const double w = bc->debye_layer_width;
    auto side_flags = [&](bool left, double yc, int which) {
      if (yc - 1e-6 < ymemb - w) return left ? bc->left_cytosol[which] : bc->right_cytosol[which];
      if (yc + 1e-6 > ymemb - w && yc - 1e-6 < ymemb + dmemb + w) return left ? bc->left_debye[which] : bc->right_debye[which];
      return left ? bc->left_extra[which] : bc->right_extra[which];
    };
    auto bits = [&](bool memb_cell, int con, int pot) -> unsigned char {
      return (unsigned char)(((con && !memb_cell) ? 0x7 : 0) | (pot ? 0x8 : 0));
    };
    // bottom / top faces (face centre decides: top tested before bottom, both before the sides)
    for (int i = 0; i < nx; ++i) {
      const bool top_is_top = y[ny] + 1e-6 > ymax, bot_is_bot = y[0] - 1e-6 < ymin;
      if (bot_is_bot) {
        const unsigned char m = bits(row_sd[0] == 2, bc->bottom[0], bc->bottom[1]);
        dirichlet_mask[i] |= m; dirichlet_mask[i + 1] |= m;
      }
      if (top_is_top) {
        const unsigned char m = bits(row_sd[ny - 1] == 2, bc->top[0], bc->top[1]);
        dirichlet_mask[(size_t)nvx * ny + i] |= m; dirichlet_mask[(size_t)nvx * ny + i + 1] |= m;
      }
    }
    for (int j = 0; j < ny; ++j) {
      const double yc = 0.5 * (y[j] + y[j + 1]);
      const bool memb = row_sd[j] == 2;
      if (!grid->left_proc_boundary && x[0] - 1e-6 < xmin_global) {
        const unsigned char m = bits(memb, side_flags(true, yc, 0), side_flags(true, yc, 1));
        dirichlet_mask[(size_t)nvx * j] |= m; dirichlet_mask[(size_t)nvx * (j + 1)] |= m;
      }
      if (!grid->right_proc_boundary && x[nx] + 1e-6 > xmax_global) {
        const unsigned char m = bits(memb, side_flags(false, yc, 0), side_flags(false, yc, 1));
        dirichlet_mask[(size_t)nvx * j + nx] |= m; dirichlet_mask[(size_t)nvx * (j + 1) + nx] |= m;
      }
    }
    for (int j = 0; j < nvy; ++j) {  // processor boundaries of the overlap: all four components
      if (grid->left_proc_boundary) dirichlet_mask[(size_t)nvx * j] = 0xF;
      if (grid->right_proc_boundary) dirichlet_mask[(size_t)nvx * j + nx] = 0xF;
    }
  }
  return AX1GPU_OK;
}

int ax1host_slab(int nx_global, int nranks, int rank, int* c0, int* c1, int* l0, int* l1) {
  if (nranks < 1 || rank < 0 || rank >= nranks || nranks > nx_global) {
    ax1::set_error("invalid slab partition (px must not exceed Nx)");
    return AX1GPU_EINVAL;
  }
  // YaspGrid-style block partition of the cell columns: the first (nx % p) ranks get one more
  const int base = nx_global / nranks, rem = nx_global % nranks;
  const int a = rank * base + std::min(rank, rem);
  const int b = a + base + (rank < rem ? 1 : 0);
  if (c0) *c0 = a;
  if (c1) *c1 = b;
  if (l0) *l0 = a - (rank > 0 ? 1 : 0);
  if (l1) *l1 = b + (rank < nranks - 1 ? 1 : 0);
  return AX1GPU_OK;
}

// ---- membrane groups (myelin / nodes of Ranvier) ------------------------------------------------
// Ax1GridGenerator::setupXPartition (ax1_gridgenerator.hh:1238-1312): intervals of the groups that carry
// position information (start, width, optional stride), sorted; gaps are filled with the default group.
int ax1host_group_partition(const ax1host_membrane_group* groups, int ngroups, int default_group, double xmax,
                            int* ival_group, double* ival_start, double* ival_end, int cap) {
  if (!groups || ngroups < 1 || default_group < 0 || default_group >= ngroups) { ax1::set_error("bad membrane groups"); return AX1GPU_EINVAL; }
  struct Ival { int g; double s, e; };
  std::vector<Ival> iv;
  for (int g = 0; g < ngroups; ++g) {
    const double start = groups[g].start, width = groups[g].width, stride = groups[g].stride;
    if (start > -1 && width > -1) {
      iv.push_back({g, start, std::min(start + width, xmax)});
      if (stride > -1) {
        if (width > stride) { ax1::set_error("membrane group: width > stride"); return AX1GPU_EINVAL; }
        double xn = start + stride;
        while (xn < xmax) { iv.push_back({g, xn, std::min(xn + width, xmax)}); xn += stride; }
      }
    }
  }
  std::sort(iv.begin(), iv.end(), [](const Ival& a, const Ival& b) { return a.s < b.s; });
  std::vector<Ival> out;
  double last_end = 0.0;
  for (const Ival& k : iv) {
    if (k.s > last_end) out.push_back({default_group, last_end, k.s});
    out.push_back(k);
    last_end = k.e;
  }
  if (last_end < xmax) out.push_back({default_group, last_end, xmax});
  last_end = 0.0;
  for (const Ival& k : out) {  // checkPartition (:1314-1334)
    if (k.s != last_end) { ax1::set_error("inconsistent membrane group partition"); return AX1GPU_EINVAL; }
    last_end = k.e;
  }
  for (int k = 0; k < (int)out.size() && k < cap; ++k) {
    if (ival_group) ival_group[k] = out[k].g;
    if (ival_start) ival_start[k] = out[k].s;
    if (ival_end) ival_end[k] = out[k].e;
  }
  return (int)out.size();
}

// Ax1GridGenerator::intervalVectorToGridVector (ax1_gridgenerator.hh:1021-1234): every interval is filled
// with cells of the group's dx; a group with smoothTransition gets n_transition cells of dx_transition next
// to a different group and (geometricRefinement) a geometric ramp (q = 0.7) between dx_transition and dx.
int ax1host_generate_x_groups(const ax1host_membrane_group* groups, int ngroups, int nival, const int* ival_group,
                              const double* ival_start, const double* ival_end, double dx_default, double xmax,
                              double* out, int cap) {
  const double q = 0.7;  // qTransition, ax1_gridgenerator.hh:62
  Axis x;
  x.p.push_back(0.0);
  for (int it = 0; it < nival; ++it) {
    const ax1host_membrane_group& G = groups[ival_group[it]];
    const double start = ival_start[it], end = ival_end[it], my_width = end - start;
    double remaining = my_width;
    const double dx = G.dx > 0 ? G.dx : dx_default;
    const bool trans_begin = it > 0 && ival_group[it] != ival_group[it - 1];
    const bool trans_end = it < nival - 1 && ival_group[it] != ival_group[it + 1];
    bool smooth_begin = false, smooth_end = false;
    double dxt_l = 0.0, dxt_r = 0.0, geo_l = 0.0, geo_r = 0.0;
    int nt_l = 0, nt_r = 0;
    if (trans_begin && G.smooth_transition) {
      const double last_w = ival_end[it - 1] - ival_start[it - 1];
      dxt_l = G.dx_transition > 0 ? G.dx_transition : std::min(my_width, last_w);
      nt_l = std::min(G.n_transition, (int)std::floor(remaining / dxt_l));
      if (nt_l > 0) {
        smooth_begin = true;
        x.uniform_steps(nt_l, dxt_l);
        remaining -= nt_l * dxt_l;
      }
    }
    if (trans_end && G.smooth_transition) {
      const double next_w = ival_end[it + 1] - ival_start[it + 1];
      dxt_r = G.dx_transition > 0 ? G.dx_transition : std::min(my_width, next_w);
      nt_r = std::min(G.n_transition, (int)std::floor(remaining / dxt_r));
      if (nt_r > 0) {
        smooth_end = true;
        remaining -= nt_r * dxt_r;
      }
    }
    if (G.geometric_refinement && remaining < my_width) {
      if (smooth_begin) {
        const int ng = (int)std::ceil(std::fabs(std::log(dxt_l / dx) / std::log(q)));
        geo_l = dx * (std::pow(q, ng + 1) - 1) / (q - 1);
      }
      if (smooth_end) {
        const int ng = (int)std::ceil(std::fabs(std::log(dxt_r / dx) / std::log(q)));
        geo_r = dx * (std::pow(q, ng + 1) - 1) / (q - 1);
      }
      double rest = remaining - (geo_l + geo_r);
      if (rest > dx) rest -= (int)std::floor(rest / dx) * dx;
      if (smooth_begin && smooth_end) { geo_l += 0.5 * rest; geo_r += 0.5 * rest; }
      else if (smooth_begin) geo_l += rest;
      else geo_r += rest;
      if (geo_l + geo_r > remaining) { ax1::set_error("membrane group too narrow for its geometric transition"); return AX1GPU_EINVAL; }
      remaining -= geo_l + geo_r;
      if (geo_l > 0.0) x.geometric(dxt_l, dx, x.last() + geo_l, true);
    }
    int n_x = 1;
    double dx_eq = remaining;
    if (remaining >= dx) { n_x = (int)std::round(remaining / dx); dx_eq = remaining / n_x; }
    const double eq_range = n_x * dx_eq;
    if (eq_range > 0) x.uniform_to(n_x, x.last() + eq_range);
    if (trans_end && nt_r > 0) {
      if (geo_r > 0.0) {
        if (std::fabs(geo_r - (end - nt_r * dxt_r - x.last())) > 1e-6) { ax1::set_error("geometric refinement to the right went wrong"); return AX1GPU_EINVAL; }
        x.geometric(dx, dxt_r, end - nt_r * dxt_r, false);
      }
      x.uniform_to(nt_r, end);
    }
  }
  if (std::fabs(x.last() - xmax) > 1e-6) { ax1::set_error("x grid does not end at xmax"); return AX1GPU_EINVAL; }
  if (out) for (int k = 0; k < (int)x.p.size() && k < cap; ++k) out[k] = x.p[k];
  return (int)x.p.size();
}

// Per membrane cell column of a rank: group index (first interval with start < centre < end,
// ax1_gridgenerator.hh:868-895), channel densities of that group, and the membrane permittivity of
// Physics::initPermittivities (acme2_cyl_physics.hh:2424-2606) incl. the smooth-step blending
// x^2 (3 - 2 x) over n_transition * dx_transition next to a group transition.
int ax1host_membrane_arrays(const ax1host_membrane_group* groups, int ngroups, int nival, const int* ival_group,
                            const double* ival_start, const double* ival_end, int smooth_permittivities,
                            double default_permittivity, double default_dmemb, int nx, const double* x,
                            double* eps_memb, double* g_leak, double* g_nav, double* g_kv, int* cell_group) {
  if (nival < 1 || nx < 1) { ax1::set_error("bad arguments"); return AX1GPU_EINVAL; }
  auto eff_perm = [&](int g) {
    const double perm = groups[g].permittivity > 0 ? groups[g].permittivity : default_permittivity;
    const double dm = groups[g].d_memb > 0 ? groups[g].d_memb : default_dmemb;
    return perm * (default_dmemb / dm);
  };
  auto width_of = [&](int k) { return ival_end[k] - ival_start[k]; };
  auto smooth_of = [&](int k) { return smooth_permittivities && groups[ival_group[k]].smooth_transition; };
  auto dtrans_of = [&](int k, int other) {
    const ax1host_membrane_group& G = groups[ival_group[k]];
    const double dxt = G.dx_transition > 0 ? G.dx_transition : std::min(width_of(k), width_of(other));
    return dxt * G.n_transition;
  };
  const double x_start = x[0];
  int git = 0;
  while (ival_end[git] < x_start && git != nival - 1) ++git;
  bool right = x_start + 1e-6 > ival_start[git];
  double cur = right ? ival_end[git] : ival_start[git];
  double d_transition = dtrans_of(git, git + 1 != nival ? git + 1 : git);
  bool smooth = smooth_of(git);
  if (nival == 1) { d_transition = 0.0; smooth = false; }
  for (int i = 0; i < nx; ++i) {
    const double xc = 0.5 * (x[i] + x[i + 1]);
    int grp = -1;
    for (int k = 0; k < nival; ++k)
      if (xc > ival_start[k] && xc < ival_end[k]) { grp = ival_group[k]; break; }
    if (grp < 0) { ax1::set_error("membrane element could not be assigned to a group"); return AX1GPU_EINVAL; }
    if (cell_group) cell_group[i] = grp;
    if (g_leak) g_leak[i] = groups[grp].g_leak;
    if (g_nav) g_nav[i] = groups[grp].g_nav;
    if (g_kv) g_kv[i] = groups[grp].g_kv;
    if (!eps_memb) continue;
    bool updated = false;
    if (!right && xc > cur + d_transition && git != nival - 1) {
      cur = ival_end[git];
      right = true;
      updated = true;
    } else if (right && xc > cur && git != nival - 1) {
      ++git;
      smooth = smooth_of(git);
      if (!smooth && git != nival - 1) { cur = ival_end[git]; right = true; }
      else { cur = ival_start[git]; right = false; }
      updated = true;
    }
    if (updated) {
      if (smooth && (right ? (git + 1 != nival) : (git != 0))) d_transition = dtrans_of(git, right ? git + 1 : git - 1);
      else d_transition = 0.0;
    }
    if (smooth && d_transition > 0.0) {
      const int gl = right ? git : git - 1, gr = right ? git + 1 : git;
      if (smooth_of(gl) && smooth_of(gr)) { ax1::set_error("only one side of a group transition may have smoothTransition"); return AX1GPU_EINVAL; }
      const double pl = eff_perm(ival_group[gl]), pr = eff_perm(ival_group[gr]);
      double xs;
      if (std::fabs(xc - cur) < d_transition) xs = right ? (xc - (cur - d_transition)) / d_transition : (xc - cur) / d_transition;
      else xs = right ? 0.0 : 1.0;
      eps_memb[i] = pl + (pr - pl) * (xs * xs * (3 - 2 * xs));
    } else {
      eps_memb[i] = eff_perm(grp);
    }
  }
  return AX1GPU_OK;
}

// guard of makeMultidomainGrid (ax1_gridgenerator.hh:903-942): an interior processor boundary must not lie
// inside group `forbidden_group` (node_of_ranvier) nor closer to a group transition than the safety distance
int ax1host_check_processor_boundary(const ax1host_membrane_group* groups, int nival, const int* ival_group,
                                     const double* ival_start, const double* ival_end, int forbidden_group,
                                     double dx_default, double xmax_rank, double xmax_global) {
  int gb = -1;
  double nearest = 0.0;
  for (int k = 0; k < nival; ++k) {
    if (ival_start[k] < xmax_rank && ival_end[k] > xmax_rank) gb = ival_group[k];
    if (std::fabs(ival_end[k] - xmax_rank) < std::fabs(nearest - xmax_rank)) nearest = ival_end[k];
  }
  if (gb >= 0 && gb == forbidden_group) { ax1::set_error("processor boundary is located at membrane group 'node_of_ranvier'"); return AX1GPU_EINVAL; }
  double safety = dx_default;
  if (gb >= 0 && groups[gb].smooth_transition) {
    const double node_width = forbidden_group >= 0 ? groups[forbidden_group].width : 1.e3;
    safety = (groups[gb].dx_transition > 0 ? groups[gb].dx_transition : node_width) * groups[gb].n_transition;
  }
  if (std::fabs(nearest - xmax_rank) < safety && xmax_rank != xmax_global) {
    ax1::set_error("processor boundary is located very close to the next membrane group transition");
    return AX1GPU_EINVAL;
  }
  return AX1GPU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Simulation-state files: the ASCII checkpoint format of Ax1SimulationState
// (dune/ax1/common/ax1_simulationstate.hh:30-57; written by Acme2CylOutput::saveState,
// acme2_cyl_output.hh:937-979, one file per rank "*_p<rank>.dat") and the x-extrapolation of a loaded
// state onto the simulation grid (doExtrapolateXData, acme2_cyl_simloader.hh:364-...: the loaded state
// is evaluated as a Q1 grid function at the vertices of the new grid).
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>

struct ax1host_state {
  int nelements = 0, timestep = 0, mpi_rank = 0, mpi_size = 1;
  double time = 0.0, dt = 0.0;
  std::vector<std::string> names;
  std::map<std::string, std::vector<double>> vec;
};

extern "C" {

int ax1host_state_load(const char* path, ax1host_state** out) {
  std::ifstream in(path);
  if (!in.good()) { ax1::set_error(std::string("cannot read state file ") + path); return AX1GPU_EINVAL; }
  ax1host_state* s = new ax1host_state;
  std::string key;
  int nvec = -1;
  while (nvec < 0 && (in >> key)) {  // header: older files have no timestep / mpi_* lines
    if (key == "#elements") in >> s->nelements;
    else if (key == "timestep") in >> s->timestep;
    else if (key == "time") in >> s->time;
    else if (key == "dt") in >> s->dt;
    else if (key == "mpi_rank") in >> s->mpi_rank;
    else if (key == "mpi_size") in >> s->mpi_size;
    else if (key == "#vectors") in >> nvec;
    else { delete s; ax1::set_error("unknown key '" + key + "' in state file header"); return AX1GPU_EINVAL; }
  }
  for (int k = 0; k < nvec; ++k) {
    std::string name;
    size_t n = 0;
    if (!(in >> name >> n)) { delete s; ax1::set_error("truncated state file"); return AX1GPU_EINVAL; }
    std::vector<double>& v = s->vec[name];
    v.resize(n);
    for (size_t i = 0; i < n; ++i)
      if (!(in >> v[i])) { delete s; ax1::set_error("truncated vector '" + name + "'"); return AX1GPU_EINVAL; }
    s->names.push_back(name);
  }
  *out = s;
  return AX1GPU_OK;
}

void ax1host_state_free(ax1host_state* s) { delete s; }

int ax1host_state_info(const ax1host_state* s, int* nelements, double* time, double* dt, int* timestep, int* mpi_rank,
                       int* mpi_size, int* nvectors) {
  if (!s) return AX1GPU_EINVAL;
  if (nelements) *nelements = s->nelements;
  if (time) *time = s->time;
  if (dt) *dt = s->dt;
  if (timestep) *timestep = s->timestep;
  if (mpi_rank) *mpi_rank = s->mpi_rank;
  if (mpi_size) *mpi_size = s->mpi_size;
  if (nvectors) *nvectors = (int)s->names.size();
  return AX1GPU_OK;
}

// returns the length of vector `name` (-1 if absent); copies min(len, cap) values if out != NULL
int ax1host_state_vector(const ax1host_state* s, const char* name, double* out, int cap) {
  auto it = s->vec.find(name);
  if (it == s->vec.end()) return -1;
  if (out) std::memcpy(out, it->second.data(), sizeof(double) * std::min<size_t>(it->second.size(), (size_t)cap));
  return (int)it->second.size();
}

int ax1host_state_save(const char* path, int nelements, int timestep, double time, double dt, int mpi_rank,
                       int mpi_size, int nvec, const char* const* names, const double* const* data, const int* sizes) {
  std::ofstream out(path);
  if (!out.good()) { ax1::set_error(std::string("cannot create state file ") + path); return AX1GPU_EINVAL; }
  out << std::scientific << std::setprecision(16);
  out << "#elements " << nelements << std::endl;
  out << "timestep " << timestep << std::endl;
  out << "time " << time << std::endl;
  out << "dt " << dt << std::endl;
  out << "mpi_rank " << mpi_rank << std::endl;
  out << "mpi_size " << mpi_size << std::endl;
  out << "#vectors " << nvec << std::endl;
  for (int i = 0; i < nvec; i++) {
    out << names[i] << " " << sizes[i] << std::endl;
    for (int j = 0; j < sizes[i]; j++) out << data[i][j] << " ";
    out << std::endl;
  }
  return out.good() ? AX1GPU_OK : AX1GPU_EINVAL;
}

// Q1 evaluation of a vertex-blocked state given on the tensor grid (xs, ys) at the vertices of (x, y).
int ax1host_state_extrapolate(int nxs, const double* xs, int nys, const double* ys, const double* us, int nx,
                              const double* x, int ny, const double* y, double* u) {
  if (nxs < 2 || nys < 2) { ax1::set_error("state grid needs at least 2x2 vertices"); return AX1GPU_EINVAL; }
  auto locate = [](const double* g, int n, double p, int& cell, double& loc) {
    int c = (int)(std::upper_bound(g, g + n, p) - g) - 1;
    c = std::max(0, std::min(c, n - 2));
    cell = c;
    loc = (p - g[c]) / (g[c + 1] - g[c]);
    loc = std::max(0.0, std::min(1.0, loc));
  };
  for (int j = 0; j < ny; ++j) {
    int cj; double eta;
    locate(ys, nys, y[j], cj, eta);
    for (int i = 0; i < nx; ++i) {
      int ci; double xi;
      locate(xs, nxs, x[i], ci, xi);
      const double* a = us + 4 * ((size_t)ci + (size_t)nxs * cj);
      const double* b = a + 4;
      const double* c = a + 4 * (size_t)nxs;
      const double* d = c + 4;
      double* o = u + 4 * ((size_t)i + (size_t)nx * j);
      // exact at the nodes of the loaded grid and when neighbouring values coincide
      auto lerp = [](double p, double q_, double t) { return t <= 0.0 ? p : (t >= 1.0 ? q_ : p + t * (q_ - p)); };
      for (int k = 0; k < 4; ++k) o[k] = lerp(lerp(a[k], b[k], xi), lerp(c[k], d[k], xi), eta);
    }
  }
  return AX1GPU_OK;
}

}  // extern "C"
