// include/ax1gpu_pdelab_shim.hh -- header-only C++ shim that models the three C++ concepts
// Acme2CylSetup::setup instantiates for the PNP hot path on top of the C ABI in ax1gpu.h:
//
//   ax1gpu::GridOperatorShim   models IGO   (residual, jacobian, preStep)   acme2_cyl_setup.hh:700-708
//   ax1gpu::SolverBackendShim  models LS    (apply, norm, result)           acme2_cyl_setup.hh:727-811
//   ax1gpu::NewtonShim         models PDESOLVER (apply, result, setters)    acme2_cyl_setup.hh:822-854
//   ax1gpu::MembraneFluxShim   models gfMembFlux (updateState, acceptTimeStep)
//                                                                           acme2_cyl_simulation.hh:457,921-927
//
// The shim is written against two tiny access traits instead of DUNE headers so that it compiles
// without DUNE (tests/cpp/test_shim.cpp uses plain std::vector containers); INTEGRATION.md gives
// the DUNE specialisations (istl::raw on BlockVector / BCRSMatrix). Errors never cross the C ABI
// as exceptions: the shim turns status codes into the exception types the reference's time loop
// catches (Dune::PDELab::NewtonError family, acme2_cyl_simulation.hh:593-692) via ErrorPolicy.
#pragma once
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>
#include "ax1gpu.h"

namespace ax1gpu {

// ---- access traits ---------------------------------------------------------------------------
// VectorAccess<V>::data(v) -> double* to 4*N contiguous doubles (vertex-blocked [Na,K,Cl,pot]).
// DUNE: &Dune::PDELab::istl::raw(v)[0][0]  (BlockVector<FieldVector<double,4>> is contiguous).
template <class V>
struct VectorAccess {
  static double* data(V& v) { return v.data(); }
  static const double* data(const V& v) { return v.data(); }
  static std::size_t size(const V& v) { return v.size(); }
};
// MatrixAccess<M>: copy BCRS-ordered 4x4 blocks (row-major, columns ascending per row) in/out.
// DUNE: iterate istl::raw(A) rows / columns in order and copy (*cit)[r][c].
template <class M>
struct MatrixAccess {
  static void from_blocks(M& A, const double* blocks, std::size_t nblocks) { A.assign(blocks, blocks + 16 * nblocks); }
  static const double* blocks(const M& A) { return A.data(); }
};

// ---- error policy ----------------------------------------------------------------------------
struct NewtonError : std::runtime_error { using std::runtime_error::runtime_error; };
struct NewtonNotConverged : NewtonError { using NewtonError::NewtonError; };
struct NewtonLinearSolverError : NewtonError { using NewtonError::NewtonError; };
struct NewtonDefectError : NewtonError { using NewtonError::NewtonError; };
struct NewtonLineSearchError : NewtonError { using NewtonError::NewtonError; };
// Default policy throws the shim's own types; with DUNE use a policy that does
// DUNE_THROW(Dune::PDELab::NewtonNotConverged, msg) etc.
struct DefaultErrorPolicy {
  static void library(int rc) {
    if (rc != AX1GPU_OK) throw std::runtime_error(std::string("libax1gpu: ") + ax1gpu_last_error());
  }
  static void newton(int status) {
    switch (status) {
      case AX1GPU_NEWTON_OK: return;
      case AX1GPU_NEWTON_NOT_CONVERGED: throw NewtonNotConverged("Newton did not converge");
      case AX1GPU_NEWTON_LINSOLVE_FAILED: throw NewtonLinearSolverError("linear solver did not converge");
      case AX1GPU_NEWTON_DEFECT_NAN: throw NewtonDefectError("non-linear defect is NaN or Inf");
      default: throw NewtonLineSearchError("line search failed");
    }
  }
};

// ---- owning handle ---------------------------------------------------------------------------
class Device {
 public:
  Device(const ax1gpu_grid& grid, const unsigned char* cell_subdomain, const double* memb_perm,
         const double* node_volume, const unsigned char* dirichlet_mask, const ax1gpu_params& params,
         const double* g_leak, const double* g_nav, const double* g_kv, int device = 0, void* stream = nullptr) {
    DefaultErrorPolicy::library(ax1gpu_create(&grid, cell_subdomain, memb_perm, node_volume, dirichlet_mask, &params,
                                              g_leak, g_nav, g_kv, device, stream, &h_));
    nv_ = (std::size_t)(grid.nx + 1) * (grid.ny + 1);
    ax1gpu_pattern(h_, nullptr, nullptr, &nblocks_);
  }
  Device(const Device&) = delete;
  Device& operator=(const Device&) = delete;
  ~Device() { ax1gpu_destroy(h_); }
  ax1gpu_handle handle() const { return h_; }
  std::size_t vertices() const { return nv_; }
  std::size_t blocks() const { return (std::size_t)nblocks_; }

 private:
  ax1gpu_handle h_ = nullptr;
  std::size_t nv_ = 0;
  int nblocks_ = 0;
};

// ---- gfMembFlux --------------------------------------------------------------------------------
template <class U, class EP = DefaultErrorPolicy>
class MembraneFluxShim {
 public:
  explicit MembraneFluxShim(Device& d) : d_(d) {}
  void updateState(double time, double dt, const U& u) {  // acme2_cyl_simulation.hh:457
    EP::library(ax1gpu_set_time(d_.handle(), time, dt));
    EP::library(ax1gpu_update_state(d_.handle(), VectorAccess<U>::data(u)));
  }
  void acceptTimeStep() { EP::library(ax1gpu_accept_state(d_.handle())); }  // :921-927
  // Output side of the grid function (acme2_cyl_output.hh:1215-1410 evaluates it per membrane element and ion for
  // FluxTerms TOTAL_FLUX .. VOLTAGE_GATED): all five terms of the device iterate in one call,
  // terms[(fluxTerm * 3 + ion) * nx + membraneColumn]
  enum FluxTerms { TOTAL_FLUX = 0, DIFF_TERM = 1, DRIFT_TERM = 2, LEAK = 3, VOLTAGE_GATED = 4 };
  void evaluateTerms(std::vector<double>& terms, int nx) const {
    terms.resize(15 * static_cast<std::size_t>(nx));
    EP::library(ax1gpu_membrane_flux_terms(d_.handle(), nullptr, terms.data()));
  }

 private:
  Device& d_;
};

// ---- IGO ---------------------------------------------------------------------------------------
template <class U, class M, class EP = DefaultErrorPolicy>
class GridOperatorShim {
 public:
  struct Traits { typedef U Domain; typedef U Range; typedef M Jacobian; };
  explicit GridOperatorShim(Device& d) : d_(d) {}
  // OneStepGridOperator::preStep/preStage for implicit Euler: stage time = time+dt, r0 = -M(uold)
  void preStep(double time, double dt, const U& uold) {
    EP::library(ax1gpu_set_time(d_.handle(), time, dt));
    EP::library(ax1gpu_set_uold(d_.handle(), VectorAccess<U>::data(uold)));
  }
  void residual(const U& x, U& r) const {
    EP::library(ax1gpu_residual(d_.handle(), VectorAccess<U>::data(x), VectorAccess<U>::data(r), nullptr));
  }
  void jacobian(const U& x, M& A) const {
    scratch_.resize(16 * d_.blocks());
    EP::library(ax1gpu_jacobian(d_.handle(), VectorAccess<U>::data(x), scratch_.data()));
    MatrixAccess<M>::from_blocks(A, scratch_.data(), d_.blocks());
  }
  Device& device() const { return d_; }

 private:
  Device& d_;
  mutable std::vector<double> scratch_;
};

// ---- LS ----------------------------------------------------------------------------------------
struct LinearSolverResult {  // Dune::PDELab::LinearSolverResult
  bool converged = false;
  unsigned int iterations = 0;
  double elapsed = 0, reduction = 0, conv_rate = 0;
};

template <class U, class M, class EP = DefaultErrorPolicy>
class SolverBackendShim {
 public:
  // n = level of fill of SeqILUn as in ISTLBackend_SEQ_BCGS_ILUn(n, w, maxiter, verbose)
  // (acme2_cyl_setup.hh:797-798 passes n = 1); n < 0 keeps the library default (block ILU0)
  SolverBackendShim(Device& d, unsigned maxiter = 5000, int slab_w = 0, int n = -1)
      : d_(d), maxiter_(maxiter), slab_w_(slab_w), n_(n) {}
  double norm(const U& v) const {  // sequential backend: v.two_norm()
    const double* p = VectorAccess<U>::data(v);
    double s = 0;
    for (std::size_t k = 0; k < VectorAccess<U>::size(v); ++k) s += p[k] * p[k];
    return std::sqrt(s);
  }
  // z = A^-1 r to relative reduction `reduction`; A is re-uploaded because PDELab hands the backend a
  // host matrix that Newton may have modified (row-norm equilibration, ax1_newton.hh:255-259)
  void apply(M& A, U& z, U& r, double reduction) {
    EP::library(ax1gpu_set_jacobian(d_.handle(), MatrixAccess<M>::blocks(A)));
    ax1gpu_lin_result lr;
    int slab_w = slab_w_;
    if (n_ >= 0) {  // the backend builds SeqILUn(A, n, w) inside apply
      EP::library(ax1gpu_ilu_factor_n(d_.handle(), slab_w_, n_));
      slab_w = 0;
    }
    EP::library(ax1gpu_solve(d_.handle(), VectorAccess<U>::data(r), reduction, (int)maxiter_, slab_w,
                             VectorAccess<U>::data(z), &lr));
    res_.converged = lr.converged != 0;
    res_.iterations = (unsigned)lr.iterations;
    res_.reduction = lr.reduction;
    res_.conv_rate = lr.it_half > 0 ? std::pow(lr.reduction, 1.0 / lr.it_half) : 0.0;
  }
  const LinearSolverResult& result() const { return res_; }

 private:
  Device& d_;
  unsigned maxiter_;
  int slab_w_, n_;
  LinearSolverResult res_;
};

// LS of the myelin build: ISTLBackend_GMRES_AMG_ILU0(gfs, maxiter, verbose, reuse, usesuperlu)
// (ax1_solverbackend.hh:279-302; instantiated at acme2_cyl_setup.hh:749-760) -- RestartedGMRes(100)
// preconditioned by one multigrid cycle with ILU0 smoothers
template <class U, class M, class EP = DefaultErrorPolicy>
class GMResAMGBackendShim {
 public:
  GMResAMGBackendShim(Device& d, unsigned maxiter = 5000, int slab_w = 0, int restart = 100, int columns_per_aggregate = 4)
      : d_(d), maxiter_(maxiter), slab_w_(slab_w), restart_(restart), cx_(columns_per_aggregate) {}
  double norm(const U& v) const {
    const double* p = VectorAccess<U>::data(v);
    double s = 0;
    for (std::size_t k = 0; k < VectorAccess<U>::size(v); ++k) s += p[k] * p[k];
    return std::sqrt(s);
  }
  void apply(M& A, U& z, U& r, double reduction) {
    EP::library(ax1gpu_set_jacobian(d_.handle(), MatrixAccess<M>::blocks(A)));
    EP::library(ax1gpu_set_amg(d_.handle(), 1, cx_));
    EP::library(ax1gpu_ilu_factor_n(d_.handle(), slab_w_, 0));   // smoother SeqILU0 + Galerkin hierarchy
    ax1gpu_lin_result lr;
    const int rc = ax1gpu_solve_gmres(d_.handle(), VectorAccess<U>::data(r), reduction, restart_, (int)maxiter_, 0,
                                      VectorAccess<U>::data(z), &lr);
    ax1gpu_set_amg(d_.handle(), 0, 0);
    EP::library(rc);
    res_.converged = lr.converged != 0;
    res_.iterations = (unsigned)lr.iterations;
    res_.reduction = lr.reduction;
    res_.conv_rate = lr.iterations > 0 ? std::pow(lr.reduction, 1.0 / lr.iterations) : 0.0;
  }
  const LinearSolverResult& result() const { return res_; }

 private:
  Device& d_;
  unsigned maxiter_;
  int slab_w_, restart_, cx_;
  LinearSolverResult res_;
};

// ---- PDESOLVER -----------------------------------------------------------------------------------
struct NewtonResult {  // Dune::PDELab::NewtonResult + the keys Ax1Newton copies to diagnostics.dat
  bool converged = false;
  unsigned int iterations = 0, linear_solver_iterations = 0;
  double first_defect = 0, defect = 0, reduction = 0, conv_rate = 0;
  double assembler_time = 0, linear_solver_time = 0, elapsed = 0;
};

template <class U, class EP = DefaultErrorPolicy>
class NewtonShim {
 public:
  explicit NewtonShim(Device& d) : d_(d) { ax1gpu_newton_opts_default(&o_); }
  void setReduction(double v) { o_.reduction = v; }                       // acme2_cyl_setup.hh:828
  void setAbsoluteLimit(double v) { o_.abs_limit = v; }                   // :829
  void setMinLinearReduction(double v) { o_.min_lin_reduction = v; }      // :830
  void setFixedLinearReduction(bool v) { o_.fixed_lin_reduction = v; }    // :831
  void setMaxIterations(unsigned v) { o_.maxit = (int)v; }                // :832
  void setForceIteration(bool v) { o_.force_iteration = v; }              // :833
  void setLineSearchMaxIterations(unsigned v) { o_.line_search_maxit = (int)v; }  // :841
  void setLinearSolverMaxIterations(unsigned v) { o_.lin_maxit = (int)v; }
  void disableLineSearch(bool v) { o_.disable_line_search = v; }          // :838-839
  void setSlabWidth(int v) { o_.slab_w = v; }
  void setIluFill(int n) { o_.ilu_fill = n; }  // n of SeqILUn in the linear solver backend (:764-765, :797-798)
  void setAmg(bool on) { o_.amg = on; }        // AMG(ILU0) preconditioner of the myelin build (:749-760)
  void setGMRes(bool on, int restart = 100) { o_.krylov = on; o_.restart = restart; }  // ax1_solverbackend.hh:201-204
  // Ax1Newton::apply(u): the whole damped Newton loop runs on the device; u in/out
  void apply(U& u) {
    ax1gpu_newton_result r;
    EP::library(ax1gpu_newton(d_.handle(), &o_, VectorAccess<U>::data(u), &r));
    res_.converged = r.converged != 0;
    res_.iterations = (unsigned)r.iterations;
    res_.linear_solver_iterations = (unsigned)r.lin_iterations;
    res_.first_defect = r.first_defect; res_.defect = r.defect; res_.reduction = r.reduction;
    res_.conv_rate = r.conv_rate; res_.assembler_time = r.t_assembler; res_.linear_solver_time = r.t_lin_solv;
    res_.elapsed = r.t_elapsed;
    EP::newton(r.status);
  }
  const NewtonResult& result() const { return res_; }

 private:
  Device& d_;
  ax1gpu_newton_opts o_;
  NewtonResult res_;
};

// ---- time loop -----------------------------------------------------------------------------------
// The body of the while(time < tend) loop of Acme2CylSimulation::run (acme2_cyl_simulation.hh:236-932)
// executed by the library with the solution vectors resident on the device: dt control (:255-439),
// updateState -> solver.apply -> uold = unew -> acceptTimeStep (:457, :589, :916-927), halve-and-retry on a
// NewtonError (:593-638). u is uploaded once (load) and read back only when the caller wants output (store).
template <class U, class EP = DefaultErrorPolicy>
class TimeLoopShim {
 public:
  TimeLoopShim(Device& d, double time0, double dtstart) : d_(d) {
    ax1gpu_timeloop_opts_default(&tl_);
    ax1gpu_newton_opts_default(&no_);
    tl_.dtstart = dtstart;
    st_.time = time0; st_.dt = dtstart; st_.its = st_.last_its = st_.steps = 0;
  }
  ax1gpu_timeloop_opts& options() { return tl_; }        // [timeStepping] keys
  ax1gpu_newton_opts& newtonOptions() { return no_; }    // [solver] keys
  void load(const U& u) { EP::library(ax1gpu_upload_u(d_.handle(), VectorAccess<U>::data(u))); }
  void store(U& u) const { EP::library(ax1gpu_download_u(d_.handle(), VectorAccess<U>::data(u))); }
  double time() const { return st_.time; }
  double dt() const { return st_.dt; }
  // one accepted time step; throws the NewtonError of the last attempt when all restarts failed
  const ax1gpu_step_result& step() {
    EP::library(ax1gpu_time_step(d_.handle(), &tl_, &no_, &st_, &res_));
    EP::newton(res_.newton_status);
    return res_;
  }

 private:
  Device& d_;
  ax1gpu_timeloop_opts tl_;
  ax1gpu_newton_opts no_;
  ax1gpu_timeloop_state st_;
  ax1gpu_step_result res_;
};

// The step body of the electroneutral (Mori) build, Acme2CylMoriSimulation::run (acme2_cyl_mori_simulation.hh:562-600):
//   ulast = unew; solverPot.apply(unew); gfMoriFlux.updateState(time, dt); solverCon.apply(time, dt, uold, unew)
// i.e. the two Ax1LinearSubProblemSolver objects of acme2_cyl_mori_setup.hh:866-893 and the charge-layer grid function
// behind one call. The caller keeps driving gfMembFlux (MembraneFluxShim::updateState / acceptTimeStep, which also
// accepts the charge-layer states).
template <class U, class EP = DefaultErrorPolicy>
class MoriStepShim {
 public:
  explicit MoriStepShim(Device& d) : d_(d) { ax1gpu_newton_opts_default(&lin_); lin_.ilu_fill = 0; }
  void setReduction(double r) { lin_.reduction = r; }            // params.getReduction()
  void setMinDefect(double m) { lin_.abs_limit = m; }            // params.getAbsLimit()
  void setLinearSolverMaxIterations(int n) { lin_.lin_maxit = n; }
  ax1gpu_newton_opts& options() { return lin_; }
  // one operator-split step from unew (= uold on entry of a fresh step); time / dt / uold as set on the device
  void apply(double time, double dt, const U& uold, U& unew) {
    EP::library(ax1gpu_set_time(d_.handle(), time, dt));
    EP::library(ax1gpu_set_uold(d_.handle(), VectorAccess<U>::data(uold)));
    EP::library(ax1gpu_mori_step(d_.handle(), &lin_, VectorAccess<U>::data(unew), &res_));
  }
  const ax1gpu_mori_result& result() const { return res_; }     // the two StationaryLinearProblemSolverResults

 private:
  Device& d_;
  ax1gpu_newton_opts lin_;
  ax1gpu_mori_result res_;
};

}  // namespace ax1gpu
