// oracle/ref_operator_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the reference's OWN local operators and parameter classes on single cells, compiled FROM /root/reference
// (never copied) against the stand-ins of oracle/ref_stubs for the DUNE interfaces they are templated on:
//   dune/ax1/acme2_cyl/operator/acme2_cyl_operator_fully_coupled.hh          Acme2CylOperatorFullyCoupled::alpha_volume
//   dune/ax1/acme2_cyl/configurations/default/default_poisson_parameters.hh  DefaultPoissonParameters::A, b, c, f
//   dune/ax1/acme2_cyl/configurations/default/default_nernst_planck_parameters.hh  DefaultNernstPlanckParameters::A, b, c, f
//   dune/ax1/acme2_cyl/operator/acme2_cyl_toperator.hh                       NernstPlanckTimeLocalOperator::alpha_volume
//   dune/ax1/acme2_cyl/operator/convectiondiffusionfem.hh                    ConvectionDiffusionFEM::alpha_volume (membrane)
//   ... and the analytic jacobian_volume of the three operators
//   Acme2CylOperatorFullyCoupled::alpha_boundary on a membrane-interface face, with DefaultNernstPlanckParameters::
//   bctype / j (surface scaling, flux stimulation) and dune/ax1/common/membranefluxgridfunction_implicit.hh
//   MembraneFluxGridFunctionImplicit::evaluate over the reference's DynamicChannelSet (channel flux arithmetic)
//   dune/ax1/common/rownorm_preconditioner.hh                                RowNormPreconditioner on a block CSR matrix
//   dune/ax1/common/power_parameters.hh, dune/ax1/acme2_cyl/common/acme2_cyl_geometrywrapper.hh
// What is stood in for: the grid (one axis-aligned cell), the Q1 local function space tree, quadrature, and a
// "Physics" object that hands back the coefficients the test chooses (permittivity, diffusion coefficients,
// valences, node volumes, scales) -- i.e. the look-ups, not the arithmetic. Built into oracle/_ref/libax1refop.so.
#define USE_CYLINDER_COORDS 1
#define AX1_USE_JACOBIAN_METHODS_IN_OPERATOR 1   // as in the reference's build flags (src/Makefile.am)
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <dune_standins.hh>

// the reference's printing / comparison helpers live in dune/ax1/common/tools.hh, which needs DUNE; the two
// comparisons the parameter classes call are restated here (tools.hh:1006-1044)
struct Tools {
  template <class T, int size>
  static bool lessOrEqualThan(const Dune::FieldVector<T, size>& v1, const Dune::FieldVector<T, size>& v2, T eps = 0.0) {
    bool result = true;
    for (int i = 0; i < size; i++) result = result && v1[i] - eps <= v2[i];
    return result;
  }
  template <class T, int size>
  static bool greaterOrEqualThan(const Dune::FieldVector<T, size>& v1, const Dune::FieldVector<T, size>& v2, T eps = 0.0) {
    bool result = true;
    for (int i = 0; i < size; i++) result = result && v1[i] + eps >= v2[i];
    return result;
  }
};

// ---- configuration stand-in (the reference's Acme2CylParameters wraps a Dune::ParameterTree): look-ups only
struct Tree {
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;
  double get(const std::string& k, double def) const { auto it = num.find(k); return it == num.end() ? def : it->second; }
  bool get(const std::string& k, bool def) const { auto it = num.find(k); return it == num.end() ? def : it->second != 0.0; }
  int get(const std::string& k, int def) const { auto it = num.find(k); return it == num.end() ? def : (int)it->second; }
  std::string get(const std::string& k, const std::string& def) const { auto it = str.find(k); return it == str.end() ? def : it->second; }
  std::string get(const std::string& k, const char* def) const { return get(k, std::string(def)); }
  std::vector<std::string> get(const std::string&, const std::vector<std::string>& def) const { return def; }
};
struct Acme2CylParameters {
  Tree general, boundary, stimulation, solution_ex, solution_in, equilibration, membrane;
  bool stim = false, volscale = true;
  double t0 = 0, t1 = 0, pot_scale = 1.0;
  double xmin = 0, xmax = 1, ymin = 0, ymax = 1, ymemb = 500.0, dmemb = 5.0;
#define AX1_BC(name) bool name() const { return false; }
  AX1_BC(isTopBoundaryDirichlet_Potential) AX1_BC(isBottomBoundaryDirichlet_Potential)
  AX1_BC(isLeftCytosolBoundaryDirichlet_Potential) AX1_BC(isRightCytosolBoundaryDirichlet_Potential)
  AX1_BC(isLeftDebyeBoundaryDirichlet_Potential) AX1_BC(isRightDebyeBoundaryDirichlet_Potential)
  AX1_BC(isLeftExtracellularBoundaryDirichlet_Potential) AX1_BC(isRightExtracellularBoundaryDirichlet_Potential)
  AX1_BC(isTopBoundaryDirichlet_Concentration) AX1_BC(isBottomBoundaryDirichlet_Concentration)
  AX1_BC(isLeftCytosolBoundaryDirichlet_Concentration) AX1_BC(isRightCytosolBoundaryDirichlet_Concentration)
  AX1_BC(isLeftDebyeBoundaryDirichlet_Concentration) AX1_BC(isRightDebyeBoundaryDirichlet_Concentration)
  AX1_BC(isLeftExtracellularBoundaryDirichlet_Concentration) AX1_BC(isRightExtracellularBoundaryDirichlet_Concentration)
#undef AX1_BC
  std::vector<double> yMemb() const { return std::vector<double>(1, ymemb); }
  double dMemb() const { return dmemb; }
  bool useMembrane() const { return true; }
  double xMin() const { return xmin; }
  double xMax() const { return xmax; }
  double yMin() const { return ymin; }
  double yMax() const { return ymax; }
  int nMembranes() const { return 1; }
  bool doStimulation() const { return stim; }
  double tInj_start() const { return t0; }
  double tInj_end() const { return t1; }
  double getPotentialResidualScalingFactor() const { return pot_scale; }
  bool doVolumeScaling() const { return volscale; }
  bool memb_neumann = false, use_mori = false;
  bool isBoundaryDirichlet_Potential(const std::string&) const { return false; }
  bool isBoundaryDirichlet_Concentration(const std::string&) const { return false; }
  bool isMembraneBoundaryNeumann_Potential() const { return memb_neumann; }
  bool useMori() const { return use_mori; }
  bool useMoriCorrection() const { return false; }
  bool equil = false;
  bool doEquilibration() const { return equil; }
  std::vector<bool> isMembraneActive() const { return std::vector<bool>(1, true); }
};

struct Output {   // printing helper of the reference (dune/ax1/common/output.hh needs DUNE); used by channelset.hh
  template <class V>
  static void printVector(const V&) {}
};

#include <dune/ax1/common/constants.hh>
#include <dune/ax1/common/power_parameters.hh>
#include <dune/ax1/acme2_cyl/common/acme2_cyl_geometrywrapper.hh>
#include <dune/ax1/acme2_cyl/configurations/default/default_poisson_parameters.hh>
#include <dune/ax1/acme2_cyl/configurations/default/default_nernst_planck_parameters.hh>
#include <dune/ax1/acme2_cyl/operator/acme2_cyl_operator_fully_coupled.hh>
#include <dune/ax1/channels/channel.hh>
#include <dune/ax1/channels/channelset.hh>
#include <dune/ax1/common/electrolyte.hh>
// The flux class keeps its clock (`time`) private and sets it in updateState(time, dt), which also walks all
// membrane intersections of the grid; the driver sets the clock directly instead.
#define private public
#include <dune/ax1/common/membranefluxgridfunction_implicit.hh>
#undef private
#ifdef AX1_REF_MORI
#include <fstream>
#include <iostream>
#include <dune/common/ios_state.hh>
// the electroneutral (Mori) model: charge layer (its init / timeStep are private, the driver calls them directly
// instead of walking the grid's membrane elements in updateState), parameter classes, the two split operators
#include <dune/ax1/common/ax1_lfs_tools.hh>
template <typename A, typename B, typename C, typename D> class ChargeLayerMori;   // CRTP tag only (charge_layer_mori.hh)
#define private public
#include <dune/ax1/common/charge_layer_mori_implicit.hh>
#undef private
#include <dune/ax1/acme2_cyl/configurations/mori/mori_poisson_parameters.hh>
#include <dune/ax1/acme2_cyl/configurations/mori/mori_nernst_planck_parameters.hh>
#include <dune/ax1/acme2_cyl/operator/acme2_cyl_mori_potential_operator.hh>
#include <dune/ax1/acme2_cyl/operator/acme2_cyl_mori_concentration_operator.hh>
#endif
#include <dune/ax1/acme2_cyl/operator/acme2_cyl_toperator.hh>
#include <dune/ax1/acme2_cyl/operator/convectiondiffusionfem.hh>
#include <istl_standins.hh>
#include <dune/ax1/common/rownorm_preconditioner.hh>

namespace {

typedef Dune::FieldVector<double, 2> Coord;

// ---- grid stand-ins: one axis-aligned cell ----------------------------------------------------
struct CellGeo {
  typedef double ctype;
  static const int dimension = 2, dimensionworld = 2, mydimension = 2, coorddimension = 2;
  double x0, x1, y0, y1;
  Dune::GeometryType type() const { return Dune::GeometryType(2); }
  int corners() const { return 4; }
  Coord corner(int i) const { Coord c; c[0] = (i & 1) ? x1 : x0; c[1] = (i & 2) ? y1 : y0; return c; }
  Coord global(const Coord& l) const { Coord g; g[0] = x0 + l[0] * (x1 - x0); g[1] = y0 + l[1] * (y1 - y0); return g; }
  Coord center() const { Coord l(0.5); return global(l); }
  double volume() const { return (x1 - x0) * (y1 - y0); }
  double integrationElement(const Coord&) const { return (x1 - x0) * (y1 - y0); }
  Dune::FieldMatrix<double, 2, 2> jacobianInverseTransposed(const Coord&) const {
    Dune::FieldMatrix<double, 2, 2> m(0.0);
    m[0][0] = 1.0 / (x1 - x0); m[1][1] = 1.0 / (y1 - y0);
    return m;
  }
};
struct Entity {
  typedef CellGeo Geometry;
  CellGeo geo;
  int index, subdomain;
  const CellGeo& geometry() const { return geo; }
};
// a membrane-interface face of the cell: the edge y = const between (x0, y) and (x1, y)
struct FaceGeo {
  typedef double ctype;
  static const int dimension = 2, dimensionworld = 2, mydimension = 1, coorddimension = 2;
  typedef Dune::FieldVector<double, 1> Local;
  double x0, x1, y;
  Dune::GeometryType type() const { return Dune::GeometryType(1); }
  int corners() const { return 2; }
  Coord corner(int i) const { Coord c; c[0] = i ? x1 : x0; c[1] = y; return c; }
  Coord global(const Local& l) const { Coord g; g[0] = x0 + l[0] * (x1 - x0); g[1] = y; return g; }
  Coord center() const { Coord g; g[0] = 0.5 * (x0 + x1); g[1] = y; return g; }
  double volume() const { return x1 - x0; }
  double integrationElement(const Local&) const { return x1 - x0; }
};
struct FaceInInside {   // the face in the local coordinates of the inside cell: its top (eta = 1) or bottom (eta = 0) edge
  double eta;
  Dune::GeometryType type() const { return Dune::GeometryType(1); }
  Coord global(const Dune::FieldVector<double, 1>& l) const { Coord g; g[0] = l[0]; g[1] = eta; return g; }
  Coord center() const { Coord g; g[0] = 0.5; g[1] = eta; return g; }
};
struct Intersection {
  typedef FaceGeo Geometry;
  typedef FaceInInside LocalGeometry;
  typedef ::Entity Entity;
  typedef const ::Entity* EntityPointer;
  FaceGeo geo;
  FaceInInside gii;
  const ::Entity* in;
  double ny;                          // outer normal (0, ny)
  // what the Physics look-ups return for this face (the test computes them as the grid-bound code would)
  double y_opposite, pot_opposite, con_opposite[3], ext_surface;
  double pot_old_this = 0.0, pot_old_opposite = 0.0, memb_perm = 2.0;   // Mori charge layer: previous time step
  int index;
  const ::Entity* outside() const { return in; }
  const FaceGeo& geometry() const { return geo; }
  const FaceInInside& geometryInInside() const { return gii; }
  const ::Entity* inside() const { return in; }
  Coord centerUnitOuterNormal() const { Coord n; n[0] = 0.0; n[1] = ny; return n; }
  Coord unitOuterNormal(const Dune::FieldVector<double, 1>&) const { return centerUnitOuterNormal(); }
};
struct IntersectionGeometry {       // PDELab's IntersectionGeometry wrapper
  typedef FaceGeo Geometry;
  typedef FaceInInside LocalGeometry;
  typedef ::Entity Entity;
  typedef const ::Entity* EntityPointer;
  typedef double ctype;
  enum { dimension = 2, dimensionworld = 2 };
  const ::Intersection* is;
  const ::Intersection& intersection() const { return *is; }
  const ::Entity* inside() const { return is->in; }
  const FaceGeo& geometry() const { return is->geo; }
  const FaceInInside& geometryInInside() const { return is->gii; }
  Coord unitOuterNormal(const Dune::FieldVector<double, 1>& x) const { return is->unitOuterNormal(x); }
};
struct GV {
  enum { dimension = 2, dimensionworld = 2 };
  typedef double ctype;
  typedef int IndexSet;
  typedef int IntersectionIterator;
  template <int c> struct Codim { typedef ::Entity Entity; typedef const ::Entity* Iterator; };
  typedef ::Intersection Intersection;
  struct Grid { enum { dimension = 2 }; typedef double ctype; };
  struct Traits { typedef GV::Grid Grid; };
  typedef GV GridViewType;
};
struct ElementGeometry {
  typedef ::Entity Entity;
  typedef CellGeo Geometry;
  const ::Entity* e;
  const ::Entity& entity() const { return *e; }
  const CellGeo& geometry() const { return e->geo; }
};

// ---- Physics stand-in: look-ups only ---------------------------------------------------------------
struct ChVec : std::vector<double> {   // vector type of the channel templates
  ChVec() {}
  explicit ChVec(int n) : std::vector<double>((size_t)n) {}
  ChVec& operator=(double v) { std::fill(begin(), end(), v); return *this; }
};
struct Physics {
  typedef double FieldType;
  typedef GV GridView;
  typedef DynamicChannelSet<double, ChVec> ChannelSet;
  typedef const Intersection* ElementIntersectionIterator;
  ChannelSet channels;
  Electrolyte<double> electro = Electrolyte<double>(80.0, 279.45, con_mol, 1e-9);
  ChannelSet& getChannelSet() { return channels; }
  const Electrolyte<double>& getElectrolyte() const { return electro; }
  // membrane look-ups of acme2_cyl_physics.hh (they need the grid's interface maps): answered from the numbers the
  // test put on the face. Restated: getMembranePotentialJump :666-721, getMembraneConcentrationRatio :927-1013
  bool isMembraneInterface(const Intersection&) const { return true; }
  int getMembraneIndex(const Intersection&) const { return 0; }
  int getMembraneNumber(const Intersection&) const { return 0; }
  ElementIntersectionIterator getPrimaryMembraneIntersection(const Intersection& is) const { return &is; }
  int getIntersectionIndex(const Intersection& is) const { return is.index; }
  double getMembraneSurfaceArea(const Intersection& is) const { return is.ext_surface; }
  template <class DGF>
  void getMembranePotentialJump(const Intersection& is, DGF&, typename DGF::Traits::RangeType& potJump,
                                const typename DGF::Traits::RangeType& uPotInside) const {
    typename DGF::Traits::RangeType pot1(uPotInside), pot2(is.pot_opposite);
    potJump = pot2;
    potJump -= pot1;
    const int ySign = Dune::sign(is.geo.y - is.y_opposite);
    potJump *= ySign;
  }
  template <class DGF>
  void getMembraneConcentrationRatio(const Intersection& is, DGF&, typename DGF::Traits::RangeType& conRatio,
                                     const typename DGF::Traits::RangeType& uConInside) const {
    typename DGF::Traits::RangeType con1(uConInside), con2;
    for (int j = 0; j < 3; ++j) con2[j] = is.con_opposite[j];
    conRatio = 0.0;
    const int ySign = Dune::sign(is.geo.y - is.y_opposite);
    if (ySign > 0) { conRatio = con2; for (int j = 0; j < 3; ++j) conRatio[j] /= con1[j]; }
    if (ySign < 0) { conRatio = con1; for (int j = 0; j < 3; ++j) conRatio[j] /= con2[j]; }
  }
  // the overload without the inside value (face-centre values on both sides, :603-660): the coupled path never
  // reaches it with fullyImplicitMembraneFlux = yes; the Mori charge layer uses it for the OLD potential jump, which
  // the test supplies on the face
  template <class DGF>
  void getMembranePotentialJump(const Intersection& is, DGF&, typename DGF::Traits::RangeType& potJump) const {
    typename DGF::Traits::RangeType pot1(is.pot_old_this), pot2(is.pot_old_opposite);
    potJump = pot2;
    potJump -= pot1;
    const int ySign = Dune::sign(is.geo.y - is.y_opposite);
    potJump *= ySign;
  }
  double getMembranePermittivity(const Intersection& is) const { return is.memb_perm; }
  // only named by the "boundary values loaded from a file" branches of the Mori parameter classes (not reached)
  typedef ::Entity Element;
  typedef const ::Entity* ElementPointer;
  int nElements(int) const { return 1; }
  int getSubDomainNumber(const Entity& e) const { return e.subdomain; }
  std::string getGroupName(const Entity&) const { return std::string("default"); }
  int getSubdomainIndex(const Entity& e) const { return e.subdomain; }
  bool mV = false;   // convertTo_mV: identity for the coupled pins (debug output only), real for the Mori charge layer
  template <class DGF>
  void getMembraneConcentrationRatio(const Intersection&, DGF&, typename DGF::Traits::RangeType&) const {
    DUNE_THROW(Dune::NotImplemented, "explicit membrane flux");
  }
  template <class V> V convertTo_mV(const V& v) const {
    if (!mV) return v;
    V o(v);
    o *= con_k * electro.getTemperature() / con_e * 1.0e3;
    return o;
  }
  Acme2CylParameters prm;
  double PC = 0, time_scale = 1e-6, length_scale = 1e-9, eps = 80.0, D[3] = {0, 0, 0};
  int z[3] = {1, 1, -1};
  std::vector<std::pair<Coord, double> > nodevol;     // node volume by vertex position
  const Acme2CylParameters& getParams() const { return prm; }
  double getPoissonConstant() const { return PC; }
  double convertFrom_mV(double v) const { return v; }
  int getElementIndex(const Entity& e) const { return e.index; }
  double getPermittivity(const Entity&) const { return eps; }
  bool isMembrane(const Entity& e) const { return e.subdomain == MEMBRANE; }
  int getValence(int j) const { return z[j]; }
  double getDiffCoeff(int j, int) const { return D[j]; }
  double getTimeScale() const { return time_scale; }
  double getLengthScale() const { return length_scale; }
  int numOfSpecies() const { return 3; }
  double getNodeVolume(const Coord& x) const {
    for (auto& nv : nodevol)
      if (std::fabs(nv.first[0] - x[0]) < 1e-6 && std::fabs(nv.first[1] - x[1]) < 1e-6) return nv.second;
    DUNE_THROW(Dune::Exception, "no node volume for this position");
  }
};
struct InitialGF {
  struct Traits { typedef Dune::FieldVector<double, 3> RangeType; typedef Coord DomainType; };
  void setTime(double) {}
  void evaluateGlobal(const Coord&, Traits::RangeType&) const {}
};
struct NoFlux {   // membrane flux grid function: not reached by the volume terms
  void setTime(double) {}
  template <class... A> void updateState(A...) {}
  template <class... A> void updateChannels(A...) {}
  template <class... A> void updateFlux(A...) {}
  template <class... A> void acceptTimeStep(A...) {}
};

// ---- Q1 local function space tree: [ power<3>(con) , pot ] -----------------------------------------
struct Q1Basis {
  struct Traits {
    typedef double DomainFieldType;
    typedef double RangeFieldType;
    typedef Coord DomainType;
    typedef Dune::FieldVector<double, 1> RangeType;
    typedef Dune::FieldMatrix<double, 1, 2> JacobianType;
  };
  unsigned int order() const { return 1; }
  unsigned int size() const { return 4; }
  void evaluateFunction(const Coord& x, std::vector<Traits::RangeType>& out) const {
    out.resize(4);
    out[0] = (1 - x[0]) * (1 - x[1]); out[1] = x[0] * (1 - x[1]); out[2] = (1 - x[0]) * x[1]; out[3] = x[0] * x[1];
  }
  void evaluateJacobian(const Coord& x, std::vector<Traits::JacobianType>& out) const {
    out.resize(4);
    out[0][0][0] = -(1 - x[1]); out[0][0][1] = -(1 - x[0]);
    out[1][0][0] = (1 - x[1]);  out[1][0][1] = -x[0];
    out[2][0][0] = -x[1];       out[2][0][1] = (1 - x[0]);
    out[3][0][0] = x[1];        out[3][0][1] = x[0];
  }
};
struct Q1Interpolation {
  template <class F, class C>
  void interpolate(const F& f, std::vector<C>& out) const {
    out.resize(4);
    for (int i = 0; i < 4; ++i) {
      Coord x; x[0] = (i & 1) ? 1.0 : 0.0; x[1] = (i & 2) ? 1.0 : 0.0;
      C y;
      f.evaluate(x, y);
      out[i] = y;
    }
  }
};
struct Q1FE {
  struct Traits { typedef Q1Basis LocalBasisType; typedef Q1Interpolation LocalInterpolationType; };
  Q1Basis b; Q1Interpolation ip;
  const Q1Basis& localBasis() const { return b; }
  const Q1Interpolation& localInterpolation() const { return ip; }
};
struct FEM { struct Traits { typedef Q1FE FiniteElementType; }; };
struct Root;
struct Leaf {
  struct Traits { typedef Q1FE FiniteElementType; typedef std::size_t SizeType; };
  typedef Root MultiDomainLocalFunctionSpace;          // the split (Mori) operators reach the other space through it
  const Root* root = nullptr;
  const Root& multiDomainLocalFunctionSpace() const { return *root; }
  int comp; Q1FE fe;
  std::size_t size() const { return 4; }
  const Q1FE& finiteElement() const { return fe; }
};
struct ConSpace {
  struct Traits { typedef std::size_t SizeType; };
  template <int k> struct Child { typedef Leaf Type; };
  typedef Root MultiDomainLocalFunctionSpace;
  const Root* root = nullptr;
  const Root& multiDomainLocalFunctionSpace() const { return *root; }
  Leaf c[3];
  const Leaf& child(int j) const { return c[j]; }
};
struct Root {
  ConSpace con; Leaf pot;
  std::size_t size() const { return 16; }
  std::size_t maxSize() const { return 16; }
  template <int k, int dummy = 0> struct Child { typedef ConSpace Type; };
  template <int dummy> struct Child<1, dummy> { typedef Leaf Type; };
  template <int k> const typename Child<k>::Type& child() const;
};
template <> const ConSpace& Root::child<0>() const { return con; }
template <> const Leaf& Root::child<1>() const { return pot; }
struct XView {
  const double* v;
  std::vector<double> own;                              // X xPot(multiLFS.maxSize()) of the split operators
  XView(const double* p) : v(p) {}
  explicit XView(std::size_t n) : v(nullptr), own(n, 0.0) { v = own.data(); }
  XView(const XView&) = delete;
  double operator()(const Leaf& l, std::size_t i) const { return v[l.comp * 4 + i]; }
};
struct RView {
  double* v;
  void accumulate(const Leaf& l, std::size_t i, double val) { v[l.comp * 4 + i] += val; }
};

struct MView {   // local matrix, row = test DOF, column = trial DOF
  double* m;
  void accumulate(const Leaf& lv, std::size_t i, const Leaf& lu, std::size_t j, double val) {
    m[(lv.comp * 4 + i) * 16 + lu.comp * 4 + j] += val;
  }
};

// discrete grid functions / solution container: the flux class constructs them but, in the implicit branch, only hands
// them to the Physics look-ups above
struct DGFBase { template <class A, class B> DGFBase(const A&, const B&) {} };
struct DGFCon {
  struct Traits { typedef GV GridViewType; typedef double RangeFieldType; typedef Dune::FieldVector<double, 3> RangeType; typedef Coord DomainType; };
  typedef DGFBase BASE_DGF;
  GV gv;
  DGFCon() {}
  DGFCon(const GV&, DGFBase&, Physics&) {}
  const GV& getGridView() const { return gv; }
};
struct DGFPot {
  struct Traits { typedef GV GridViewType; typedef double RangeFieldType; typedef Dune::FieldVector<double, 1> RangeType; typedef Coord DomainType; };
  DGFPot() {}
  template <class A, class B> DGFPot(const A&, const B&) {}
};
// the solution vectors the split operators read through SolutionContainer: on the one-cell harness a "vector" is the
// 16 local coefficients of that cell, a local view copies them
struct LocalVec {
  const double* xl = nullptr;
  template <class Cache>
  struct ConstLocalView {
    const LocalVec& u;
    explicit ConstLocalView(const LocalVec& u_) : u(u_) {}
    void bind(const Cache&) {}
    template <class X> void read(X& x) const { for (int k = 0; k < 16; ++k) x.own[k] = u.xl[k]; }
    void unbind() {}
  };
};
struct SolutionContainer {
  typedef LocalVec U;
  LocalVec potNew, conNew, conOld, potOld;
  int getGfsPot() const { return 0; }
  int getGfsCon() const { return 0; }
  const LocalVec* getSolutionPotNew() const { return &potNew; }
  const LocalVec* getSolutionConNew() const { return &conNew; }
  const LocalVec* getSolutionConOld() const { return &conOld; }
  const LocalVec* getSolutionPotOld() const { return &potOld; }
};
typedef MembraneFluxGridFunctionImplicit<DGFCon, DGFPot, Physics, SolutionContainer> MembFlux;

typedef DefaultPoissonParameters<GV, double, Physics> ParamPot;
typedef DefaultNernstPlanckParameters<GV, double, Physics, InitialGF, MembFlux, NoFlux> ParamNP;
typedef PowerParameters<ParamNP, 3> ParamCon;
typedef Acme2CylOperatorFullyCoupled<ParamCon, ParamPot, FEM, FEM> LopElec;
typedef NernstPlanckTimeLocalOperator<Physics, FEM> LopTime;
typedef Dune::PDELab::ConvectionDiffusionFEM<ParamPot, FEM> LopMemb;

// everything one call needs, built from the flat argument arrays
struct Setup {
  Physics ph;
  Entity e;
  GV gv;
  InitialGF igf;
  NoFlux nf;
  Root lfs;
  DGFCon dgfCon;
  DGFPot dgfPot;
  SolutionContainer sol;
  Dune::shared_ptr<MembFlux> flux;
  Setup(const double* cell, const double* phys, const int* valence, const double* nodevol4, int do_volume_scaling,
        const double* stim, int subdomain) {
    ph.eps = phys[0]; ph.D[0] = phys[1]; ph.D[1] = phys[2]; ph.D[2] = phys[3]; ph.PC = phys[4];
    ph.prm.pot_scale = phys[5]; ph.time_scale = phys[6];
    for (int j = 0; j < 3; ++j) ph.z[j] = valence[j];
    ph.prm.volscale = do_volume_scaling != 0;
    ph.prm.stim = stim[0] != 0.0; ph.prm.t0 = stim[2]; ph.prm.t1 = stim[3];
    ph.prm.stimulation.num["position_x"] = stim[4]; ph.prm.stimulation.num["position_y"] = stim[5];
    ph.prm.stimulation.num["I_inj"] = stim[6];
    ph.prm.general.num["useElecOperatorJacobianVolume"] = 1;
    ph.prm.general.num["useMembOperatorJacobianVolume"] = 1;
    ph.prm.general.num["useTimeOperatorJacobianVolume"] = 1;
    e.geo.x0 = cell[0]; e.geo.x1 = cell[1]; e.geo.y0 = cell[2]; e.geo.y1 = cell[3];
    e.index = 0; e.subdomain = subdomain;
    for (int i = 0; i < 4; ++i) ph.nodevol.push_back(std::make_pair(e.geo.corner(i), nodevol4[i]));
    for (int j = 0; j < 3; ++j) { lfs.con.c[j].comp = j; lfs.con.c[j].root = &lfs; }
    lfs.pot.comp = 3; lfs.pot.root = &lfs; lfs.con.root = &lfs;
    ph.prm.boundary.num["fullyImplicitMembraneFlux"] = 1;
    flux.reset(new MembFlux(dgfCon, dgfPot, ph, sol, 0.0));
  }
};

}  // namespace

extern "C" {

// argument conventions of all entry points:
//   cell  = {x0, x1, y0, y1}
//   xl    = 16 local coefficients, comp * 4 + vertex, comp = Na, K, Cl, pot, vertices x fastest
//   phys  = {eps, D_Na, D_K, D_Cl, poisson constant, potential residual scale, time scale}
//   nodevol4 = node volumes of the 4 vertices (used when do_volume_scaling)
//   stim  = {doStimulation, LOP time, t_inj_start, t_inj_end, position_x, position_y, I_inj}
//   results are the operator's contributions (not weighted by dt), DOF = comp * 4 + vertex
// which = 0: Acme2CylOperatorFullyCoupled (electrolyte cell), 1: NernstPlanckTimeLocalOperator, 2: ConvectionDiffusionFEM
// with DefaultPoissonParameters on a membrane cell (phys[0] = membrane permittivity).
// jac = 0: alpha_volume into out[0..15]; jac = 1: the operator's analytic jacobian_volume into out[0..255]
// (row = test DOF, column = trial DOF, DOF = comp * 4 + vertex). Arguments as ax1ref_elec_alpha_volume.
int ax1ref_local_operator(int which, int jac, const double* cell, const double* xl, const double* phys,
                          const int* valence, const double* nodevol4, int do_volume_scaling, const double* stim,
                          double* out) {
  try {
    Setup S(cell, phys, valence, nodevol4, do_volume_scaling, stim, which == 2 ? MEMBRANE : ES);
    ParamPot paramPot(S.gv, S.ph);
    ParamCon paramCon;
    for (int j = 0; j < 3; ++j)
      paramCon.push_back(Dune::shared_ptr<ParamNP>(new ParamNP(j, S.gv, S.ph, S.igf, *S.flux, S.nf, 0.0)));
    XView x(xl);
    ElementGeometry eg{&S.e};
    const int n = jac ? 256 : 16;
    for (int k = 0; k < n; ++k) out[k] = 0.0;
    RView r{out};
    MView m{out};
    if (which == 0) {
      LopElec lop(paramCon, paramPot, 4);
      lop.setTime(stim[1]);
      if (jac) lop.jacobian_volume(eg, S.lfs, x, S.lfs, m);
      else lop.alpha_volume(eg, S.lfs, x, S.lfs, r);
    } else if (which == 1) {
      LopTime lop(S.ph);
      if (jac) lop.jacobian_volume(eg, S.lfs.con, x, S.lfs.con, m);
      else lop.alpha_volume(eg, S.lfs.con, x, S.lfs.con, r);
    } else if (which == 2) {
      LopMemb lop(paramPot, 4);
      if (jac) lop.jacobian_volume(eg, S.lfs.pot, x, S.lfs.pot, m);
      else lop.alpha_volume(eg, S.lfs.pot, x, S.lfs.pot, r);
    } else {
      return 2;
    }
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

// Acme2CylOperatorFullyCoupled::alpha_boundary on the membrane-interface face of an electrolyte cell
// (from_cytosol: the face is the cell's top edge and the membrane lies above it; otherwise its bottom edge).
//   memb  = {y of the opposite interface, potential and Na, K, Cl at the opposite face centre, extracellular
//            membrane surface 2 pi r L of this column}
//   geff  = new effective conductances of the channel set [Na_leak, K_leak, Nav, Kv] of this membrane element
//   equil = {doEquilibration, flux-GF time, tEquilibrium, dummy_value}
//   fstim = {memb_flux_stimulation, ion (0 Na, 1 K, 2 Cl), memb_flux_inj}   (time window and position_x from `stim`)
int ax1ref_elec_alpha_boundary(const double* cell, int from_cytosol, const double* xl, const double* phys,
                               const int* valence, const double* nodevol4, int do_volume_scaling, const double* stim,
                               const double* memb, const double* geff, const double* equil, const double* fstim,
                               double temperature, double* out16) {
  try {
    Setup S(cell, phys, valence, nodevol4, do_volume_scaling, stim, from_cytosol ? CYTOSOL : ES);
    S.ph.electro = Electrolyte<double>(80.0, temperature, con_mol, 1e-9);
    S.ph.prm.equil = equil[0] != 0.0;
    S.ph.prm.general.num["dummy_value"] = equil[3];
    S.ph.prm.stimulation.num["memb_flux_stimulation"] = fstim[0];
    S.ph.prm.stimulation.str["memb_flux_ion"] = fstim[1] == 0 ? "Na" : (fstim[1] == 1 ? "K" : "Cl");
    S.ph.prm.stimulation.num["memb_flux_inj"] = fstim[2];
    // channel set of this membrane element with the given effective conductances: leak channels g * ratio with
    // ratio 1, voltage-gated channels gbar * prod(gates^exponent) with all gates 1
    typedef Physics::ChannelSet CS;
    CS& cs = S.ph.channels;
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(Na)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(K)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new NavChannel<double, ChVec>()));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new KvChannel<double, ChVec>()));
    const int elem = 7;
    cs.addMembraneElement(elem);
    cs.resize();
    for (int k = 0; k < cs.size(); ++k) cs.getChannelPtr(k)->init();
    ChVec ones(1); ones = 1.0;
    for (int k = 0; k < 4; ++k) {
      cs.setConductance(k, elem, geff[k]);
      if (k < 2) cs.setRatio(k, elem, 1.0);
      else for (int g = 0; g < cs.getChannel(k).numGatingParticles(); ++g) cs.getChannelPtr(k)->setGatingParticle(g, ones);
    }
    // the flux class: a new object sees the channel set above; its clock decides the equilibration branch
    S.flux.reset(new MembFlux(S.dgfCon, S.dgfPot, S.ph, S.sol, equil[2]));
    S.flux->time = equil[1];
    ParamPot paramPot(S.gv, S.ph);
    ParamCon paramCon;
    for (int j = 0; j < 3; ++j)
      paramCon.push_back(Dune::shared_ptr<ParamNP>(new ParamNP(j, S.gv, S.ph, S.igf, *S.flux, S.nf, equil[2])));
    LopElec lop(paramCon, paramPot, 4);
    lop.setTime(stim[1]);
    Intersection is;
    is.geo.x0 = cell[0]; is.geo.x1 = cell[1]; is.geo.y = from_cytosol ? cell[3] : cell[2];
    is.gii.eta = from_cytosol ? 1.0 : 0.0;
    is.in = &S.e;
    is.ny = from_cytosol ? 1.0 : -1.0;
    is.y_opposite = memb[0]; is.pot_opposite = memb[1];
    for (int j = 0; j < 3; ++j) is.con_opposite[j] = memb[2 + j];
    is.ext_surface = memb[5];
    is.index = elem;
    IntersectionGeometry ig{&is};
    XView x(xl);
    for (int k = 0; k < 16; ++k) out16[k] = 0.0;
    RView r{out16};
    lop.alpha_boundary(ig, S.lfs, x, S.lfs, r);
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

// MembraneFluxGridFunctionImplicit::evaluate (membranefluxgridfunction_implicit.hh:83-361) alone, for its five
// output terms (FluxTerms: 0 total, 1 diffusion term, 2 drift term, 3 leak channels, 4 voltage-gated channels) --
// what acme2_cyl_output.hh:1215-1410 writes as memb_flux / memb_flux_diff_term / ... per ion. The potential and the
// concentrations at this side's face centre are handed in (the `implicit` overload), the opposite side through
// `memb` as in ax1ref_elec_alpha_boundary. out15[term * 3 + ion].
int ax1ref_membrane_flux_terms(int from_cytosol, const double* upot_ucon4, const double* memb, const double* geff,
                               const double* equil, double temperature, double* out15) {
  try {
    const double cell[4] = {0.0, 1.0, 0.0, 1.0};
    const double phys[7] = {80.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1e-6};
    const int valence[3] = {1, 1, -1};
    const double nodevol4[4] = {1.0, 1.0, 1.0, 1.0};
    const double stim[7] = {0, 0, 0, 0, 0, 0, 0};
    Setup S(cell, phys, valence, nodevol4, 0, stim, from_cytosol ? CYTOSOL : ES);
    S.ph.electro = Electrolyte<double>(80.0, temperature, con_mol, 1e-9);
    S.ph.prm.equil = equil[0] != 0.0;
    S.ph.prm.general.num["dummy_value"] = equil[3];
    typedef Physics::ChannelSet CS;
    CS& cs = S.ph.channels;
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(Na)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(K)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new NavChannel<double, ChVec>()));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new KvChannel<double, ChVec>()));
    const int elem = 7;
    cs.addMembraneElement(elem);
    cs.resize();
    for (int k = 0; k < cs.size(); ++k) cs.getChannelPtr(k)->init();
    ChVec ones(1); ones = 1.0;
    for (int k = 0; k < 4; ++k) {
      cs.setConductance(k, elem, geff[k]);
      if (k < 2) cs.setRatio(k, elem, 1.0);
      else for (int g = 0; g < cs.getChannel(k).numGatingParticles(); ++g) cs.getChannelPtr(k)->setGatingParticle(g, ones);
    }
    S.flux.reset(new MembFlux(S.dgfCon, S.dgfPot, S.ph, S.sol, equil[2]));
    S.flux->time = equil[1];
    Intersection is;
    is.geo.x0 = cell[0]; is.geo.x1 = cell[1]; is.geo.y = from_cytosol ? cell[3] : cell[2];
    is.gii.eta = from_cytosol ? 1.0 : 0.0;
    is.in = &S.e;
    is.ny = from_cytosol ? 1.0 : -1.0;
    is.y_opposite = memb[0]; is.pot_opposite = memb[1];
    for (int j = 0; j < 3; ++j) is.con_opposite[j] = memb[2 + j];
    is.ext_surface = memb[5];
    is.index = elem;
    Dune::FieldVector<double, 1> uPot(upot_ucon4[0]);
    Dune::FieldVector<double, 3> uCon;
    for (int j = 0; j < 3; ++j) uCon[j] = upot_ucon4[1 + j];
    Dune::FieldVector<double, 1> xloc(0.5);
    for (int term = 0; term < 5; ++term) {
      Dune::FieldVector<double, 3> y(0.0);
      S.flux->evaluate(is, xloc, y, term, -1, true, uPot, uCon);
      for (int j = 0; j < 3; ++j) out15[3 * term + j] = y[j];
    }
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

// RowNormPreconditioner<GV, 4>(A, b) (rownorm_preconditioner.hh:65-155) on a block CSR matrix with 4 x 4 blocks given
// as rowptr / col / vals (16 doubles per block, row-major) and the right-hand side r (4 per block row); in place.
int ax1ref_rownorm(int n, const int* rowptr, const int* col, double* vals, double* r) {
  try {
    typedef Dune::FieldMatrix<double, 4, 4> Block;
    Dune::BCRSMatrix<Block> A;
    Dune::BlockVector<Dune::FieldVector<double, 4> > b;
    A.rows.resize(n); b.resize(n);
    for (int i = 0; i < n; ++i) {
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        Block B;
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) B[a][c] = vals[16 * (size_t)k + 4 * a + c];
        A.rows[i].cols.push_back(col[k]);
        A.rows[i].blocks.push_back(B);
      }
      for (int a = 0; a < 4; ++a) b[i][a] = r[4 * (size_t)i + a];
    }
    { RowNormPreconditioner<GV, 4> prec(A, b); }
    for (int i = 0; i < n; ++i) {
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) vals[16 * (size_t)k + 4 * a + c] = A.rows[i].blocks[k - rowptr[i]][a][c];
      for (int a = 0; a < 4; ++a) r[4 * (size_t)i + a] = b[i][a];
    }
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

}  // extern "C"
