// oracle/ref_stubs/dune_standins.hh -- TEST INFRASTRUCTURE ONLY (see README.md).
// Stand-ins for the dune-geometry / dune-localfunctions / dune-pdelab names the reference's local operators and
// parameter classes use, so that those classes can be compiled from /root/reference and run on single cells:
// Gauss-Legendre rules on the cube (2 points per direction for order <= 3), the reference element's centre, the Q1
// local basis / interpolation with dune's vertex numbering (x fastest), and the PDELab base classes as empty shells.
#pragma once
#include <cstddef>
#include <string>
#include <vector>
#include <dune/common/fvector.hh>
#include <dune/common/exceptions.hh>
#include <dune/common/shared_ptr.hh>

namespace Dune {

class GeometryType {
 public:
  explicit GeometryType(unsigned dim = 2) : dim_(dim) {}
  unsigned dim() const { return dim_; }
  bool isCube() const { return true; }
 private:
  unsigned dim_;
};

template <class ct, int dim>
class QuadraturePoint {
 public:
  QuadraturePoint(const FieldVector<ct, dim>& x, ct w) : x_(x), w_(w) {}
  const FieldVector<ct, dim>& position() const { return x_; }
  ct weight() const { return w_; }
 private:
  FieldVector<ct, dim> x_;
  ct w_;
};
template <class ct, int dim>
class QuadratureRule : public std::vector<QuadraturePoint<ct, dim> > {};

// dune-geometry's Gauss-Legendre points on [0,1] for 2 points (exact to order 3), tensorised, first coordinate fastest
template <class ct, int dim>
struct QuadratureRules {
  static const QuadratureRule<ct, dim>& rule(const GeometryType&, int order) {
    static QuadratureRule<ct, dim> r2;
    if (order > 3) DUNE_THROW(Dune::NotImplemented, "stand-in quadrature: order <= 3 only");
    if (r2.empty()) {
      const ct g[2] = {0.2113248654051871177454256097490212721762, 0.7886751345948128822545743902509787278238};
      const int n = dim == 1 ? 2 : 4;
      for (int k = 0; k < n; ++k) {
        FieldVector<ct, dim> x;
        ct w = 1.0;
        for (int d = 0; d < dim; ++d) { x[d] = g[(k >> d) & 1]; w *= 0.5; }
        r2.push_back(QuadraturePoint<ct, dim>(x, w));
      }
    }
    return r2;
  }
};

template <class ct, int dim>
struct ReferenceElement {
  FieldVector<ct, dim> position(int, int) const { return FieldVector<ct, dim>(0.5); }   // centre of the cube
};
template <class T> int sign(const T& v) { return (v > T(0)) - (v < T(0)); }

template <class ct, int dim>
struct ReferenceElements {
  static const ReferenceElement<ct, dim>& general(const GeometryType&) { static ReferenceElement<ct, dim> r; return r; }
};

namespace PDELab {

// PDELab's finite-difference Jacobians are not stood in for: the drivers only call the analytic jacobian_volume
template <class Imp> struct NumericalJacobianVolume {
  template <class... A> void jacobian_volume(A&&...) const { DUNE_THROW(Dune::NotImplemented, "numerical Jacobian"); }
};
template <class Imp> struct NumericalJacobianApplyVolume {};
template <class Imp> struct NumericalJacobianBoundary {};
template <class Imp> struct NumericalJacobianApplyBoundary {};
template <class Imp> struct NumericalJacobianSkeleton {};
template <class Imp> struct NumericalJacobianApplySkeleton {};
struct FullSkeletonPattern {};
struct FullVolumePattern {};
struct LocalOperatorDefaultFlags {};
template <class R>
class InstationaryLocalOperatorDefaultMethods {
 public:
  typedef R RealType;
  void setTime(R t) { time_ = t; }
  R getTime() const { return time_; }
 private:
  R time_ = R();
};
template <class LB>
class LocalBasisCache {};

// boundary grid functions (the membrane flux class derives from these)
template <class GV, class RF, int m, class R>
struct IntersectionGridFunctionTraits {
  typedef GV GridViewType;
  typedef typename GV::ctype DomainFieldType;
  enum { dimDomain = GV::dimension - 1 };
  typedef Dune::FieldVector<DomainFieldType, GV::dimension - 1> DomainType;
  typedef RF RangeFieldType;
  enum { dimRange = m };
  typedef R RangeType;
};
template <class T, class Imp>
struct IntersectionGridFunctionBase {
  typedef T Traits;
};

struct ConvectionDiffusionBoundaryConditions {
  enum Type { Dirichlet = 1, Neumann = -1, Outflow = -2, None = -3 };
};
template <class GV, class RF>
struct ConvectionDiffusionParameterTraits {
  typedef GV GridViewType;
  enum { dimDomain = GV::dimension };
  typedef typename GV::ctype DomainFieldType;
  typedef Dune::FieldVector<DomainFieldType, dimDomain> DomainType;
  typedef Dune::FieldVector<DomainFieldType, dimDomain - 1> IntersectionDomainType;
  typedef RF RangeFieldType;
  typedef Dune::FieldVector<RF, dimDomain> RangeType;
  typedef Dune::FieldMatrix<RF, dimDomain, dimDomain> PermTensorType;
  typedef typename GV::template Codim<0>::Entity ElementType;
  typedef typename GV::Intersection IntersectionType;
};

}  // namespace PDELab
}  // namespace Dune

namespace Dune { namespace PDELab {
// index cache of a local function space (the split Mori operators build one to read the other sub-problem's
// coefficients); nothing to cache on the one-cell harness
template <class LFS>
struct LFSIndexCache {
  const LFS& lfs;
  explicit LFSIndexCache(const LFS& l) : lfs(l) {}
  void update() {}
};
} }
