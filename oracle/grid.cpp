// oracle/grid.cpp -- TEST INFRASTRUCTURE ONLY.
// Restatement of the radial grid-vector generator of dune-ax1 for the single-membrane,
// geometrically refined case (every BASELINE config uses it):
//   dune/ax1/common/ax1_gridvector.hh:15-31,108-170      (GridVector primitives)
//   dune/ax1/common/ax1_gridgenerator.hh:74-130          (onesidedGeometricTowardsMembrane)
//   dune/ax1/common/ax1_gridgenerator.hh:132-149         (equidistantMembrane)
//   dune/ax1/common/ax1_gridgenerator.hh:232-277         (onesidedGeometricFromMembrane)
//   dune/ax1/common/ax1_gridgenerator.hh:357-450         (generateTensorGrid, y part)
// Pinned bit-exactly by the 181-point y vector shipped in
//   src/simulation_states/equilibrium/cyl/yasp_square10mm_p0.dat:7-8  (tests/golden/).
#include "oracle.hpp"
#include <algorithm>
#include <stdexcept>

namespace ax1o {
namespace {

struct GV : std::vector<double> {
  void start(double v) { if (empty()) push_back(v); }
  void equidistant_n_h(int n, double h) {                     // gridvector.hh:15-21
    if (empty()) push_back(0.0);
    for (int i = 0; i < n; i++) push_back(back() + h);
  }
  void equidistant_n_xend(int n, double xend) {               // gridvector.hh:23-31
    if (empty()) push_back(0.0);
    double h = (xend - back()) / n;
    for (int i = 0; i < n - 1; i++) push_back(back() + h);
    push_back(xend);
  }
  double newton(int n, double x_s, double x_e, double h) {    // gridvector.hh:108-129
    double m = (x_e - x_s) / h;
    double xold = 0.0, xnew = x_e - x_s;
    while (std::fabs(xnew - xold) > 1E-8) {
      xold = xnew;
      xnew = xold - (-std::pow(xold, n) + m * xold - m + 1) / (-n * std::pow(xold, n - 1) + m);
    }
    if (std::fabs(xnew - 1) < 1E-6) {
      xold = x_e - x_s;
      xnew = 0.0;
      while (std::fabs(xnew - xold) > 1E-8) {
        xold = xnew;
        xnew = xold - (-std::pow(xold, n) + m * xold - m + 1) / (-n * std::pow(xold, n - 1) + m);
      }
    }
    return xnew;
  }
  void geometric_n_h0_xend(int n, double h0, double xend) {   // gridvector.hh:131-142
    if (empty()) push_back(0.0);
    double q = newton(n, back(), xend, h0);
    double h = h0;
    for (int i = 0; i < n - 1; i++) { push_back(back() + h); h *= q; }
    push_back(xend);
  }
  void geometric_n_hend_xend(int n, double hend, double xend) {  // gridvector.hh:144-156
    if (empty()) push_back(0.0);
    double p = newton(n, back(), xend, hend);
    double h = hend * std::pow(p, n - 1);
    for (int i = 0; i < n - 1; i++) { push_back(back() + h); h /= p; }
    push_back(xend);
  }
  void geometric_h0_hend_xend(double h0, double hend, double xend, int guarantee = 0) {  // :158-168
    if (empty()) push_back(0.0);
    double q = (xend - back() - h0) / (xend - back() - hend);
    int n = (int)(std::log(hend / h0) / std::log(q)) + 1;
    if (guarantee == 0) geometric_n_h0_xend(n, h0, xend);
    else geometric_n_hend_xend(n, hend, xend);
  }
};

const double qTransition = 0.7;  // ax1_gridgenerator.hh:62

void onesided_towards_membrane(GV& y, double dy_m, int n_dy_m, double max_dy, double yMemb) {
  double yStart = y.back();
  double q = qTransition;
  int n_transition = (int)std::ceil(std::fabs(std::log(dy_m / max_dy) / std::log(q)));
  double geoRefineRange = std::max(max_dy, dy_m) * (std::pow(q, n_transition + 1) - 1) / (q - 1);
  double geoRefineDistance = n_dy_m * dy_m + geoRefineRange;
  double yrange_equidistant = yMemb - geoRefineDistance - yStart;
  int nEquidistant = (int)std::floor(yrange_equidistant / max_dy);
  if (yMemb - geoRefineDistance < yStart)
    throw std::runtime_error("cannot reach requested membrane resolution with this dy_cell");
  if (nEquidistant > 0) y.equidistant_n_h(nEquidistant, max_dy);
  y.geometric_h0_hend_xend(max_dy, dy_m, yMemb - n_dy_m * dy_m);
  y.equidistant_n_h(n_dy_m, dy_m);
}

void onesided_from_membrane(GV& y, double dy_m, int n_dy_m, double max_dy, double yEnd) {
  double yStart = y.back();
  y.equidistant_n_h(n_dy_m, dy_m);
  double q = qTransition;
  int n_transition = (int)std::ceil(std::fabs(std::log(dy_m / max_dy) / std::log(q)));
  double geoRefineRange = max_dy * (std::pow(q, n_transition + 1) - 1) / (q - 1);
  double geoRefineDistance = n_dy_m * dy_m + geoRefineRange;
  double yrange_equidistant = yEnd - yStart - geoRefineDistance;
  int nEquidistant = (int)std::floor(yrange_equidistant / max_dy);
  y.geometric_h0_hend_xend(dy_m, max_dy, yEnd - (nEquidistant * max_dy), 1);
  if (nEquidistant > 0) y.equidistant_n_xend(nEquidistant, yEnd);
}

}  // namespace

std::vector<double> generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell,
                               double dy_cell_min, int n_dy_cell_min, double dy, double dy_min,
                               int n_dy_min) {
  GV y;
  y.start(ymin);
  if (ymemb > y.back() + 1e-6) onesided_towards_membrane(y, dy_cell_min, n_dy_cell_min, dy_cell, ymemb);
  if (ymemb + dmemb > y.back() + 1e-6 && dmemb > 0) y.equidistant_n_h(1, dmemb);
  onesided_from_membrane(y, dy_min, n_dy_min, dy, ymax);
  if (std::fabs(y.back() - ymax) > 1e-6) throw std::runtime_error("y grid does not end at ymax");
  return y;
}

// the GridVector primitives on their own (checked against the reference's compiled ax1_gridvector.hh through
// oracle/_ref, tests/test_reference_pins.py); op numbers as in oracle/ref_driver.cpp
std::vector<double> gridvector_op(int op, double start, int n, double a, double b, double c) {
  GV g;
  g.push_back(start);
  switch (op) {
    case 0: g.equidistant_n_h(n, a); break;
    case 1: g.equidistant_n_xend(n, a); break;
    case 6: g.geometric_n_h0_xend(n, a, b); break;
    case 7: g.geometric_n_hend_xend(n, a, b); break;
    case 8: g.geometric_h0_hend_xend(a, b, c, 0); break;
    case 9: g.geometric_h0_hend_xend(a, b, c, 1); break;
    default: throw std::runtime_error("gridvector_op: primitive not restated");
  }
  return g;
}

// refineXDirection branch of generateTensorGrid (ax1_gridgenerator.hh:536-547)
std::vector<double> equidistant_x(double xmin, double xmax, double dx) {
  GV x;
  x.start(xmin);
  int n_x = (int)std::ceil((xmax - xmin) / dx);
  x.equidistant_n_xend(n_x, xmax);
  return x;
}

}  // namespace ax1o
