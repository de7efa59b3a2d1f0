// oracle/ref_newton_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// C entry points around the reference's Newton class, compiled FROM /root/reference (never copied):
//   dune/ax1/common/ax1_newton.hh            Ax1Newton::line_search (:308-503), ::defect (:194-199), ::prepare_step (:216-293)
//   dune/ax1/common/ax1_solution_container.hh, rownorm_preconditioner.hh (as included by it)
// The PDELab Newton bases Ax1Newton derives from are stand-ins (ref_stubs/newton_standins.hh) that carry only the
// state those overrides touch; the grid operator is a stand-in whose residual() calls back into the test. What this
// pins is therefore the Ax1-specific part of the Newton solver: the Hackbusch-Reusken search with the
// accept-best / accept-full-step-anyway fallbacks, its NaN handling, and that prepare_step equilibrates (A, r) with
// RowNormPreconditioner. Built into oracle/_ref/libax1refnewton.so by `make -C oracle _ref`.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <map>
#include <string>
#include <vector>

#include <dune/common/exceptions.hh>
#include <dune/common/fvector.hh>
#include <dune/common/shared_ptr.hh>
#include <dune_standins.hh>
#include <istl_standins.hh>
#include <dune/ax1/common/constants.hh>
#include <dune/ax1/common/ax1_newton.hh>

namespace {

typedef double (*defect_cb)(void* ctx, const double* u, int n);

struct Comm { int rank() const { return 0; } };
struct GV {   // the typedefs RowNormPreconditioner and the reordering helper mention
  enum { dimension = 2, dimensionworld = 2 };
  typedef double ctype;
  typedef int IndexSet;
  typedef int IntersectionIterator;
  template <int c> struct Codim { typedef int Entity; typedef const int* Iterator; };
  struct Grid { enum { dimension = 2 }; typedef double ctype; };
  struct Traits { typedef GV::Grid Grid; typedef int IndexSet; };
  Comm comm() const { return Comm(); }
};
struct GFS {
  struct Traits { typedef GV GridViewType; };
  int n;
  GV gv;
  int size() const { return n; }
  template <int k> const GFS& child() const { return *this; }
  const GV& gridView() const { return gv; }
};

typedef Dune::FieldVector<double, 4> VBlock;
struct Vec : Dune::BlockVector<VBlock> {
  typedef Dune::BlockVector<VBlock> BaseT;
  typedef double ElementType;
  Vec() {}
  explicit Vec(const GFS& g) { resize(g.n / 4); *this = 0.0; }
  Vec& operator=(double v) { for (VBlock& b : *this) for (int a = 0; a < 4; ++a) b[a] = v; return *this; }
  void axpy(double a, const Vec& z) { for (size_t i = 0; i < size(); ++i) for (int c = 0; c < 4; ++c) (*this)[i][c] += a * z[i][c]; }
  int flatsize() const { return 4 * (int)size(); }
  BaseT& base() { return *this; }
};

struct GridOp;
struct Mat : Dune::BCRSMatrix<Dune::FieldMatrix<double, 4, 4> > {
  typedef Dune::BCRSMatrix<Dune::FieldMatrix<double, 4, 4> > BaseT;
  Mat() {}
  explicit Mat(const GridOp&) {}
  BaseT& base() { return *this; }
};

// grid operator: the residual is a single number per trial u, delivered as r = (defect, 0, 0, ...)
struct GridOp {
  struct Traits {
    typedef GFS TrialGridFunctionSpace;
    typedef GFS TestGridFunctionSpace;
    typedef Mat Jacobian;
  };
  GFS gfs;
  defect_cb cb;
  void* ctx;
  std::vector<double>* trace;   // step lengths are recovered by the test from u; here we count evaluations
  int evals = 0;
  const GFS& testGridFunctionSpace() const { return gfs; }
  const GFS& trialGridFunctionSpace() const { return gfs; }
  void residual(const Vec& u, Vec& r) {
    std::vector<double> flat(gfs.n);
    for (size_t i = 0; i < u.size(); ++i) for (int c = 0; c < 4; ++c) flat[4 * i + c] = u[i][c];
    if (trace) trace->insert(trace->end(), flat.begin(), flat.end());
    r[0][0] = cb(ctx, flat.data(), gfs.n);
    ++evals;
  }
};
struct Solver {
  double norm(const Vec& r) const { return r[0][0]; }   // NaN passes through as the real two-norm would
};
struct Lop {
  int pre = 0;
  void preAssembly() { ++pre; }
};
struct DiagInfo {
  void registerDebugData(const std::string&, double) {}
  void setDebugData(const std::string&, double) {}
};
struct SubGfs { SubGfs() {} explicit SubGfs(const GFS&) {} };
struct Dgf { Dgf(const SubGfs&, const Vec&) {} };
struct AcmeOut {
  typedef SubGfs GFS_POT;
  typedef SubGfs GFS_CON;
  struct Traits {
    typedef SubGfs GFS_POT_SUB;
    typedef SubGfs GFS_CON_SUB;
    typedef Dgf DGF_POT;
    typedef Dgf DGF_CON;
  };
  DiagInfo di;
  DiagInfo& getDiagInfo() { return di; }
  void initGnuplotFile(const std::string&, const std::string&) {}
  template <class G> void writeGFToGnuplot(const G&, const std::string&, const std::string&) {}
};

typedef Ax1Newton<AcmeOut, Lop, GridOp, Solver, Vec, Vec> Newton;
typedef Ax1SolutionContainer<Vec, SubGfs, SubGfs> Container;   // = Newton's (private) SolutionContainer typedef

}  // namespace

extern "C" {

// One Ax1Newton::line_search call. u (n = 4 * blocks, in/out), z the Newton correction, cur_defect / prev_defect the
// state NewtonSolver::apply would hold on entry, strategy as in Ax1Newton::LineSearchStrategy (0 none, 1 HR,
// 2 accept best, 3 accept best no throw). u_trace (cap_evals * n doubles) receives every u the defect was
// evaluated at; out = {accepted defect, number of defect evaluations, number of preAssembly calls}.
// Returns 0, 1 if NewtonLineSearchError was thrown, 3 if NewtonDefectError escaped (u / out are still filled),
// 2 on any other exception.
int ax1ref_line_search(int n, double* u, const double* z, double cur_defect, double prev_defect, int maxit,
                       double damping, int strategy, int fully_implicit, defect_cb cb, void* ctx, double* u_trace,
                       int cap_evals, double* out) {
  int rc = 0;
  std::vector<double> trace;
  try {
    GridOp go;
    go.gfs.n = n; go.cb = cb; go.ctx = ctx; go.trace = &trace;
    Vec uu(go.gfs), zz(go.gfs), rr(go.gfs);
    for (int i = 0; i < n; ++i) { uu[i / 4][i % 4] = u[i]; zz[i / 4][i % 4] = z[i]; }
    Solver s;
    Lop lop;
    AcmeOut ao;
    SubGfs sp, sc;
    Container container(sp, sc);
    Newton newton(ao, lop, go, uu, s, "ref", container);
    newton.setLineSearchStrategy((Newton::LineSearchStrategy)strategy);
    newton.setLineSearchMaxIterations(maxit);
    newton.setLineSearchDampingFactor(damping);
    newton.setFullyImplicit(fully_implicit != 0);
    newton.res.defect = cur_defect;
    newton.res.first_defect = cur_defect;
    newton.prev_defect = prev_defect;
    try {
      newton.line_search(zz, rr);
    } catch (Dune::PDELab::NewtonLineSearchError&) {
      rc = 1;
    } catch (Dune::PDELab::NewtonDefectError&) {
      rc = 3;   // a NaN defect at the accepted u (outside the search's own try block) ends the Newton solve
    }
    for (int i = 0; i < n; ++i) u[i] = uu[i / 4][i % 4];
    out[0] = newton.res.defect;
    out[1] = go.evals;
    out[2] = lop.pre;
    // the solution container must point at the vector line_search moved (:346-347)
    out[3] = (container.getSolutionConNew() == &uu && container.getSolutionPotNew() == &uu) ? 1.0 : 0.0;
    int ne = std::min(go.evals, cap_evals);
    if (u_trace) std::memcpy(u_trace, trace.data(), sizeof(double) * (size_t)ne * n);
  } catch (std::exception&) {
    return 2;
  }
  return rc;
}

// Ax1Newton::prepare_step on a caller-supplied block CSR Jacobian (4 x 4 blocks, 16 doubles each, row-major) and
// residual: with row_preconditioner set it must leave (A, r) equilibrated exactly as RowNormPreconditioner does.
int ax1ref_prepare_step(int nb, const int* rowptr, const int* col, double* vals, double* r, int row_preconditioner) {
  try {
    GridOp go;
    go.gfs.n = 4 * nb; go.cb = 0; go.ctx = 0; go.trace = 0;
    Vec uu(go.gfs), rr(go.gfs);
    Mat A;
    A.rows.resize(nb);
    for (int i = 0; i < nb; ++i) {
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        Dune::FieldMatrix<double, 4, 4> B;
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) B[a][c] = vals[16 * (size_t)k + 4 * a + c];
        A.rows[i].cols.push_back(col[k]);
        A.rows[i].blocks.push_back(B);
      }
      for (int a = 0; a < 4; ++a) rr[i][a] = r[4 * (size_t)i + a];
    }
    Solver s;
    Lop lop;
    AcmeOut ao;
    SubGfs sp, sc;
    Container container(sp, sc);
    Newton newton(ao, lop, go, uu, s, "ref", container);
    newton.setLineSearchStrategy(Newton::hackbuschReuskenAcceptBestNoThrow);
    newton.setRowPreconditioner(row_preconditioner != 0);
    newton.prepare_step(A, rr);
    for (int i = 0; i < nb; ++i) {
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) vals[16 * (size_t)k + 4 * a + c] = A.rows[i].blocks[k - rowptr[i]][a][c];
      for (int a = 0; a < 4; ++a) r[4 * (size_t)i + a] = rr[i][a];
    }
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

}  // extern "C"
