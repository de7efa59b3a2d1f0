// oracle/ref_stubs/newton_standins.hh -- TEST INFRASTRUCTURE ONLY (see README.md).
// Stand-ins for the PDELab 2.0 Newton building blocks the reference's Ax1Newton derives from (ax1_newton.hh:25-28):
// just the state its overrides read and write (u, res, prev_defect, reassembled, verbosity_level, the line-search
// maxit / damping factor) and NewtonSolver::defect = residual + norm + NaN check. The Newton loop itself
// (NewtonSolver::apply) lives in PDELab and is NOT restated here; only Ax1Newton's own overrides are exercised.
#pragma once
#include <cmath>
#include <ostream>
#include <string>
#include <dune/common/exceptions.hh>
#include <istl_standins.hh>

namespace Dune {

class Timer {
 public:
  double elapsed() const { return 0.0; }
};
template <class M> void writeMatrixToMatlab(const M&, const std::string&) {}
template <class V> void printvector(std::ostream&, const V&, const std::string&, const std::string&, int = 1, int = 10, int = 2) {}

namespace PDELab {

class FilenameHelper {
 public:
  explicit FilenameHelper(const std::string& b) : base_(b), i_(0) {}
  void increment() { ++i_; }
  std::string getName() const { return base_ + "-" + std::to_string(i_); }
 private:
  std::string base_;
  int i_;
};

namespace istl {
template <class T> struct raw_type { typedef typename T::BaseT type; };
template <class T> typename raw_type<T>::type& raw(T& t) { return t; }
}  // namespace istl

class NewtonError : public Dune::Exception { public: using Dune::Exception::Exception; };
class NewtonDefectError : public NewtonError { public: using NewtonError::NewtonError; };
class NewtonLinearSolverError : public NewtonError { public: using NewtonError::NewtonError; };
class NewtonLineSearchError : public NewtonError { public: using NewtonError::NewtonError; };
class NewtonNotConverged : public NewtonError { public: using NewtonError::NewtonError; };

struct NewtonResult {
  bool converged = false;
  double first_defect = 0, defect = 0, reduction = 1, conv_rate = 1, assembler_time = 0, linear_solver_time = 0, elapsed = 0;
  int iterations = 0, linear_solver_iterations = 0;
};

template <class GOS, class TrlV, class TstV>
class NewtonBase {
 public:
  typedef typename GOS::Traits::Jacobian Matrix;
  NewtonBase(GOS& go, TrlV& u_) : gridoperator(go), u(&u_) {}
  explicit NewtonBase(GOS& go) : gridoperator(go), u(0) {}
  virtual ~NewtonBase() {}
  void setVerbosityLevel(unsigned v) { verbosity_level = v; }
  GOS& gridoperator;
  TrlV* u;
  NewtonResult res;
  unsigned verbosity_level = 0;
  double prev_defect = 0, linear_reduction = 0;
  bool reassembled = false;
  virtual void defect(TstV& r) = 0;
  virtual void line_search(TrlV& z, TstV& r) = 0;
  virtual void prepare_step(Matrix& A, TstV& r) = 0;
};

template <class GOS, class S, class TrlV, class TstV>
class NewtonSolver : public virtual NewtonBase<GOS, TrlV, TstV> {
 public:
  NewtonSolver(GOS& go, TrlV& u_, S& s) : NewtonBase<GOS, TrlV, TstV>(go, u_), solver(s) {}
  NewtonSolver(GOS& go, S& s) : NewtonBase<GOS, TrlV, TstV>(go), solver(s) {}
  void apply() { DUNE_THROW(Dune::NotImplemented, "the PDELab Newton loop is not part of the stand-ins"); }
  virtual void defect(TstV& r) {
    r = 0.0;
    this->gridoperator.residual(*this->u, r);
    this->res.defect = solver.norm(r);
    if (!std::isfinite(this->res.defect)) DUNE_THROW(NewtonDefectError, "non-finite defect");
  }
 protected:
  S& solver;
};

template <class GOS, class TrlV, class TstV>
class NewtonTerminate : public virtual NewtonBase<GOS, TrlV, TstV> {
 public:
  NewtonTerminate(GOS& go, TrlV& u_) : NewtonBase<GOS, TrlV, TstV>(go, u_) {}
  explicit NewtonTerminate(GOS& go) : NewtonBase<GOS, TrlV, TstV>(go) {}
};

template <class GOS, class TrlV, class TstV>
class NewtonLineSearch : public virtual NewtonBase<GOS, TrlV, TstV> {
 public:
  NewtonLineSearch(GOS& go, TrlV& u_) : NewtonBase<GOS, TrlV, TstV>(go, u_) {}
  explicit NewtonLineSearch(GOS& go) : NewtonBase<GOS, TrlV, TstV>(go) {}
  void setLineSearchMaxIterations(unsigned m) { maxit = m; }
  void setLineSearchDampingFactor(double d) { damping_factor = d; }
 protected:
  unsigned maxit = 10;
  double damping_factor = 0.5;
};

template <class GOS, class TrlV, class TstV>
class NewtonPrepareStep : public virtual NewtonBase<GOS, TrlV, TstV> {
 public:
  typedef typename GOS::Traits::Jacobian Matrix;
  NewtonPrepareStep(GOS& go, TrlV& u_) : NewtonBase<GOS, TrlV, TstV>(go, u_) {}
  explicit NewtonPrepareStep(GOS& go) : NewtonBase<GOS, TrlV, TstV>(go) {}
  // the real one decides on reassembly and evaluates the Jacobian; here the caller supplies A and it counts as fresh
  virtual void prepare_step(Matrix&, TstV&) { this->reassembled = true; }
};

}  // namespace PDELab
}  // namespace Dune
