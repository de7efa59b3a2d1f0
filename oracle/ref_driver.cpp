// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around three headers of the reference, compiled FROM /root/reference (never copied):
//   dune/ax1/common/ax1_gridvector.hh   GridVector primitives behind the radial / axial grid vectors
//   dune/ax1/channels/channel.hh        LeakChannel, NavChannel, KvChannel (rates, steady state, BE step)
//   dune/ax1/channels/channelset.hh     DynamicChannelSet (initChannels, leak re-balancing, timeStep, acceptState)
//   dune/ax1/common/electrolyte.hh      Electrolyte (Poisson constant)
//   dune/ax1/acme2_cyl/common/acme2_cyl_geometrywrapper.hh   Acme2CylGeometry (2 pi r weights of cells and faces)
// Built into oracle/_ref/libax1ref.so by `make -C oracle _ref` when /root/reference is present. It pins the
// oracle's restatement of these parts (SURVEY rows a10 and the grid generator primitives) against the reference's
// own code; everything else on the hot path needs DUNE and stays "parity unpinned" (DESIGN 3).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <dune/common/exceptions.hh>
#include <dune/common/fvector.hh>
#include <dune/common/shared_ptr.hh>
// stand-in for the reference's printing helper (dune/ax1/common/output.hh needs DUNE): channelset.hh calls
// Output::printVector in deserializeChannelStates
struct Output {
  template <class V>
  static void printVector(const V&) {}
};
#include <dune/ax1/channels/channel.hh>          // pulls in constants.hh (con_pi, debug streams) first
#include <dune/ax1/common/ax1_gridvector.hh>
#include <dune/ax1/channels/channelset.hh>
#include <dune/ax1/common/electrolyte.hh>
#include <dune/ax1/acme2_cyl/common/acme2_cyl_geometrywrapper.hh>

namespace {
// vector type handed to the channel templates (they need resize, [], size, begin/end and "= scalar")
struct Vec : std::vector<double> {
  Vec() {}
  explicit Vec(int n) : std::vector<double>((size_t)n) {}   // the channel base class constructs g(0)
  Vec& operator=(double v) { std::fill(begin(), end(), v); return *this; }
};
// axis-aligned host geometry (what YaspGrid hands out): a cell [x0,x1] x [y0,y1] (mydim 2) or one of its edges
// (mydim 1, from corner a to corner b); local coordinates in [0,1]^mydim
template <int mydim>
struct BoxGeo {
  typedef double ctype;
  static const int dimension = 2, dimensionworld = 2, mydimension = mydim, coorddimension = 2;
  typedef Dune::FieldVector<double, 2> Global;
  typedef Dune::FieldVector<double, mydim> Local;
  double a[2], b[2];
  BoxGeo(double ax, double ay, double bx, double by) { a[0] = ax; a[1] = ay; b[0] = bx; b[1] = by; }
  int corners() const { return 1 << mydim; }
  Global corner(int i) const {
    Global c;
    if (mydim == 2) { c[0] = (i & 1) ? b[0] : a[0]; c[1] = (i & 2) ? b[1] : a[1]; }
    else { c[0] = i ? b[0] : a[0]; c[1] = i ? b[1] : a[1]; }
    return c;
  }
  Global global(const Local& l) const {
    Global g;
    if (mydim == 2) { g[0] = a[0] + l[0] * (b[0] - a[0]); g[1] = a[1] + l[1] * (b[1] - a[1]); }
    else { g[0] = a[0] + l[0] * (b[0] - a[0]); g[1] = a[1] + l[0] * (b[1] - a[1]); }
    return g;
  }
  double volume() const {
    if (mydim == 2) return (b[0] - a[0]) * (b[1] - a[1]);
    return std::fabs(b[0] - a[0]) + std::fabs(b[1] - a[1]);   // axis-aligned edge
  }
  double integrationElement(const Local&) const { return volume(); }
};
typedef DynamicChannelSet<double, Vec> ChannelSet;
typedef ChannelSet::VCON VCON;
}  // namespace

extern "C" {

// GridVector primitives; `start` is the point the vector already holds (the generator appends to it)
// op: 0 equidistant_n_h(n,a)  1 equidistant_n_xend(n,a)  2 equidistant_h_xend(a,b)  3 geometric_n_h0_q(n,a,b)
//     4 geometric_h0_q_xend(a,b,c)  5 geometric_n_q_xend(n,a,b)  6 geometric_n_h0_xend(n,a,b)
//     7 geometric_n_hend_xend(n,a,b)  8 / 9 geometric_h0_hend_xend(a,b,c, guarantee = 0 / 1)
int ax1ref_gridvector(int op, double start, int n, double a, double b, double c, double* out, int cap) {
  GridVector<double> g;
  g.push_back(start);
  try {
    switch (op) {
      case 0: g.equidistant_n_h(n, a); break;
      case 1: g.equidistant_n_xend(n, a); break;
      case 2: g.equidistant_h_xend(a, b); break;
      case 3: g.geometric_n_h0_q(n, a, b); break;
      case 4: g.geometric_h0_q_xend(a, b, c); break;
      case 5: g.geometric_n_q_xend(n, a, b); break;
      case 6: g.geometric_n_h0_xend(n, a, b); break;
      case 7: g.geometric_n_hend_xend(n, a, b); break;
      case 8: g.geometric_h0_hend_xend(a, b, c, 0); break;
      case 9: g.geometric_h0_hend_xend(a, b, c, 1); break;
      default: return -1;
    }
  } catch (...) { return -2; }
  for (int i = 0; i < (int)g.size() && i < cap; ++i) out[i] = g[i];
  return (int)g.size();
}

// Electrolyte::getPoissonConstant / getDebyeLength (electrolyte.hh): the scaling constant of the Poisson equation
void ax1ref_electrolyte(double permittivity, double temperature, double std_con, double length_scale, double* out) {
  Electrolyte<double> e(permittivity, temperature, std_con, length_scale);
  out[0] = e.getPoissonConstant();
  out[1] = e.getDebyeLength();
}

// Acme2CylGeometry (acme2_cyl_geometrywrapper.hh:56-72) over an axis-aligned cell / edge:
// out[0] = integrationElement(local) = 2 pi r(local) |J|, out[1] = volume() = pi (r_last + r_first) vol2D
void ax1ref_cyl_geometry(int mydim, double ax, double ay, double bx, double by, double l0, double l1, double* out) {
  if (mydim == 2) {
    BoxGeo<2> host(ax, ay, bx, by);
    Acme2CylGeometry<BoxGeo<2> > g(host);
    BoxGeo<2>::Local l; l[0] = l0; l[1] = l1;
    out[0] = g.integrationElement(l); out[1] = g.volume();
  } else {
    BoxGeo<1> host(ax, ay, bx, by);
    Acme2CylGeometry<BoxGeo<1> > g(host);
    BoxGeo<1>::Local l; l[0] = l0;
    out[0] = g.integrationElement(l); out[1] = g.volume();
  }
}

// One membrane element with the default channel set [Na_leak, K_leak, Nav, Kv] (channel_builder order):
//  - set maximal conductances and the [equilibration] leak ratios,
//  - initChannels(v_init): steady state at v_init, v_rest := v_init, leak re-balancing,
//  - nsteps times: timeStep(dt_s, v[k]) + acceptState (backward Euler from the accepted state).
// out (per stage, stage 0 = after initChannels): [ratioNa, ratioK, m, h, n, g_eff(Na_leak, K_leak, Nav, Kv)] = 9 doubles
int ax1ref_channelset(double g_leak, double g_nav, double g_kv, double ratio_na, double ratio_k, double v_init,
                      double dt_s, const double* v, int nsteps, double* out) {
  try {
    ChannelSet cs;
    cs.addChannel(Dune::shared_ptr<Channel<double, Vec> >(new LeakChannel<double, Vec>(Na)));
    cs.addChannel(Dune::shared_ptr<Channel<double, Vec> >(new LeakChannel<double, Vec>(K)));
    cs.addChannel(Dune::shared_ptr<Channel<double, Vec> >(new NavChannel<double, Vec>()));
    cs.addChannel(Dune::shared_ptr<Channel<double, Vec> >(new KvChannel<double, Vec>()));
    const int elem = 7;   // element index (arbitrary), mapped to membrane index 0
    cs.addMembraneElement(elem);
    cs.resize();
    for (int k = 0; k < cs.size(); ++k) cs.getChannelPtr(k)->init();
    cs.setConductance(0, elem, g_leak); cs.setRatio(0, elem, ratio_na);
    cs.setConductance(1, elem, g_leak); cs.setRatio(1, elem, ratio_k);
    cs.setConductance(2, elem, g_nav);
    cs.setConductance(3, elem, g_kv);
    VCON cin(0.0), cex(0.0);
    cs.initChannels(elem, v_init, cin, cex);
    auto dump = [&](double* o) {
      o[0] = cs.getGatingParticle(0, 0, elem); o[1] = cs.getGatingParticle(1, 0, elem);
      o[2] = cs.getNewGatingParticle(2, 0, elem); o[3] = cs.getNewGatingParticle(2, 1, elem);
      o[4] = cs.getNewGatingParticle(3, 0, elem);
      for (int k = 0; k < 4; ++k) o[5 + k] = cs.getNewEffConductance(k, elem);
    };
    dump(out);
    for (int s = 0; s < nsteps; ++s) {
      cs.timeStep(elem, dt_s, v[s], cin, cex);
      dump(out + 9 * (s + 1));
      cs.acceptState();
    }
    return 0;
  } catch (std::exception& e) {
    return 1;
  }
}

}  // extern "C"
