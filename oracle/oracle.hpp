// oracle/oracle.hpp -- TEST INFRASTRUCTURE ONLY (CPU restatement of the dune-ax1 PNP hot path).
//
// This is a plain C++17 restatement of the reference algorithm, written from the reference
// sources cited at each function (paths relative to the reference repository). It is NOT part of
// the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it. The product path (dune_ax1_b200/) never calls into oracle/.
//
// PARITY STATUS. The reference application cannot be built here (needs DUNE 2.3 + MPI + autotools, see
// DESIGN.md) and its own tests hold no golden vector for residual / Jacobian / solve. What pins this oracle:
// (1) the shipped equilibrium state files (tests/golden/, tests/test_oracle_golden.py) -- grid vector bit-exact,
//     channel steady states, steady-state residual cancellation;
// (2) the parts of the reference that compile from its own headers against the stand-ins of oracle/ref_stubs
//     (oracle/_ref, tests/test_reference_pins.py): grid-vector primitives, HH channel classes, electrolyte,
//     cylinder geometry wrapper, the local operators (alpha_volume, analytic jacobian_volume, alpha_boundary with
//     the implicit membrane flux) with their parameter classes, RowNormPreconditioner, Ax1Newton::line_search and
//     ::prepare_step.
// Still "parity unpinned" (restated from the pinned PDELab / ISTL versions, which are not in the tree): the
// assembler (element loop, constraints, implicit-Euler combination, scatter), the PDELab Newton loop and its
// linear-reduction rule, BiCGStab / GMRES / ILU(n) / AMG.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <vector>
#include <string>
#include "ax1_oracle.h"

namespace ax1o {

// dune/ax1/common/constants.hh:64-74
constexpr double con_pi   = 3.1415926535897932384626433;
constexpr double con_e    = 1.602176487e-19;
constexpr double con_k    = 1.380650e-23;
constexpr double con_eps0 = 8.854187817e-12;
constexpr double con_mol  = 6.02214129e23;

enum { Na = 0, K = 1, Cl = 2, NSPEC = 3, POT = 3, BS = 4 };
enum { CYTOSOL = 0, ES = 1, MEMBRANE = 2 };  // constants.hh (subdomain enum)
enum { CH_NA_LEAK = 0, CH_K_LEAK = 1, CH_NAV = 2, CH_KV = 3, NCHAN = 4 };

// Flat configuration (the INI keys the hot path reads, acme2_cyl_parametertree.hh); the POD
// lives in ax1_oracle.h so that ctypes can fill it.
struct Config : ax1o_config {
  Config() { ax1o_config_default(this); }
  Config(const ax1o_config& c) : ax1o_config(c) {}
};

// One rank's local problem (interior + overlap cells), cf. ax1_gridgenerator.hh:601-631.
struct Problem {
  Config cfg;
  int nx = 0, ny = 0;              // local cells
  std::vector<double> x, y;        // local coordinates
  int nvx = 0, nvy = 0, nv = 0;    // vertices
  bool left_proc_boundary = false, right_proc_boundary = false;
  int own_lo = 0, own_hi = 0;      // owned local vertex columns [own_lo, own_hi]
  // derived constants
  double PC = 0, kT_e = 0, D[NSPEC] = {0, 0, 0};
  // Physics maps
  std::vector<uint8_t> cell_sub;    // nx*ny subdomain tag
  int jm = -1;                      // membrane cell row
  std::vector<double> eps_memb;     // per membrane cell column
  std::vector<double> node_volume;  // per vertex
  std::vector<uint8_t> dirichlet;   // per vertex, bit c = component c constrained
  // channels (per membrane column = ES-side membrane edge, physics.hh:2041-2061)
  std::vector<double> gbar[3];      // leak, Nav, Kv max conductances [mS/cm^2]
  std::vector<double> ratio[2];     // Na_leak, K_leak ratios
  std::vector<double> gate[3], gate_new[3];  // m, h, n accepted / new
  double v_rest = -65.0;
  bool channels_initialized = false, channels_partially_initialized = false;
  // time
  double time = 0, dt = 0;          // step start time and dt (flux GF); LOP time = time+dt
  std::vector<double> uold, r0;     // previous step and constant residual part -M(uold)
  // radial split of the ILU sub-slabs (additive Schwarz in y as well): rows j < ilu_jcut and j >= ilu_jcut of a
  // sub-slab are factorised separately, couplings across the cut are dropped from the preconditioner. 0 = off
  // (the reference: one ILU per rank). The device places the cut in the uniform far field, where it costs no
  // Krylov iterations and halves the dependency depth of the sweeps.
  int ilu_jcut = 0;

  // charge layer of the electroneutral (Mori) model: gamma_j per membrane column, cytosol (int) / ES (ext) side,
  // accepted (old) / new (charge_layer_mori_implicit.hh); empty until the first mori_update_state
  std::vector<double> mori_gamma_int_old[NSPEC], mori_gamma_int_new[NSPEC];
  std::vector<double> mori_gamma_ext_old[NSPEC], mori_gamma_ext_new[NSPEC];
  bool mori_initialized = false, mori_partially_initialized = false;

  int vid(int i, int j) const { return i + nvx * j; }
  int cid(int i, int j) const { return i + nx * j; }
};

// block CSR matrix with 4x4 blocks (Dune::BCRSMatrix<FieldMatrix<double,4,4>>)
struct BCRS {
  int n = 0;
  std::vector<int> rowptr, col;
  std::vector<double> val;  // 16 per block, row-major
  double* block(int k) { return &val[16 * (size_t)k]; }
  const double* block(int k) const { return &val[16 * (size_t)k]; }
  int find(int r, int c) const {
    for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) if (col[k] == c) return k;
    return -1;
  }
};

// ---- grid.cpp
std::vector<double> generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell,
                               double dy_cell_min, int n_dy_cell_min, double dy, double dy_min,
                               int n_dy_min);
std::vector<double> equidistant_x(double xmin, double xmax, double dx);
std::vector<double> gridvector_op(int op, double start, int n, double a, double b, double c);

// ---- model.cpp
void setup_problem(Problem& p);           // tags, node volumes, constraints, constants
void build_pattern(const Problem& p, BCRS& A);
void channels_static_init(Problem& p);    // channelset.hh:133-182
void update_state(Problem& p, const double* u);   // membranefluxgridfunction_implicit.hh:429-501
void accept_state(Problem& p);            // :524-546
void membrane_potential_mV(const Problem& p, const double* u, double* vm);  // physics.hh:562-603

// ---- assembly.cpp
void set_uold(Problem& p, const double* uold);    // r0 = -M(uold)
void residual(const Problem& p, const double* u, double* r, double* abs_terms = nullptr);
void jacobian(const Problem& p, const double* u, BCRS& A);
void rownorm_scale(BCRS& A, double* r);           // rownorm_preconditioner.hh:65-155
void cell_jacobian(const Problem& p, int i, int j, int which, const double* xl, double weight, double* out256);
void cell_kernel(const Problem& p, int i, int j, int which, const double* xl, const double* uglob, double weight,
                 double* out);

// ---- solver.cpp
struct ILU {
  BCRS LU;                  // factor with fill pattern; diagonal blocks stored inverted
  std::vector<int> diag;    // index of diagonal block per row
};
// sub-slab additive-Schwarz decomposition inside one rank: vertex columns grouped into slabs of
// width slab_w (<=0: one slab = plain ILU on the local matrix as the reference does)
void ilu_factor(const Problem& p, const BCRS& A, int fill_level, int slab_w, ILU& f);
void ilu_factor_nvx(int nvx, const BCRS& A, int fill_level, int slab_w, ILU& f, int jcut = 0);  // vertex v lies in column v % nvx, row v / nvx
void ilu_apply(const ILU& f, const double* d, double* v);
void spmv(const BCRS& A, const double* x, double* y);

struct MoriResult : ax1o_mori_result { MoriResult() : ax1o_mori_result{0, 0, 0, 0.0, 0.0, 0.0, 0.0} {} };
struct LinResult : ax1o_lin_result { LinResult() : ax1o_lin_result{0, 0, 0.0, 0.0} {} };
struct Comm {  // collectives for the overlapping multi-rank path (identity for one rank)
  void* ctx = nullptr;
  void (*halo_add)(void* ctx, int rank, double* v) = nullptr;
  double (*allreduce_sum)(void* ctx, int rank, double local) = nullptr;
  int rank = 0;
};
using Prec = std::function<void(const double* d, double* v)>;  // rank-local v = M^-1 d
void bicgstab(const Problem& p, const BCRS& A, const ILU& M, double* x, double* b, double reduction,
              int maxit, const Comm* comm, LinResult& res);
void bicgstab(const Problem& p, const BCRS& A, const Prec& M, double* x, double* b, double reduction,
              int maxit, const Comm* comm, LinResult& res);
void zero_constrained(const Problem& p, double* v);
void op_apply(const Problem& p, const BCRS& A, const double* x, double* y, const Comm* comm);
void prec_apply(const Problem& p, const Prec& M, const double* d, double* v, std::vector<double>& dd,
                const Comm* comm);
double masked_dot(const Problem& p, const double* a, const double* b, const Comm* comm);

// ---- gmres_amg.cpp
// Dune::RestartedGMResSolver (dune-istl 2.3 solvers.hh) as driven by ISTLBackend_GMRES_AMG
// (dune/ax1/common/ax1_solverbackend.hh:196-213: restart = 100, recalcDefect = false)          [ext]
void gmres(const Problem& p, const BCRS& A, const Prec& M, double* x, double* b, double reduction,
           int restart, int maxit, const Comm* comm, LinResult& res);
// x-aggregation multigrid with block-ILU smoothers: the CHECKER of the device cycle (csrc/amg.cu).
// Dune::Amg's greedy aggregation [ext, dune-istl aggregates.hh] is not in /root/reference and is not
// reproducible bit for bit; a preconditioner changes iteration counts only, never converged answers.
struct AMGx {
  int cx = 4, nvy = 0;
  struct Level {
    int nvx = 0;
    BCRS A;            // level 0 holds a copy of the fine matrix
    ILU M;
    std::vector<double> b, x;
  };
  std::vector<Level> lv;
  std::vector<uint8_t> mask;   // fine-level constrained DOFs (bit c of vertex v)
  std::vector<double> t, s;
};
void amg_build(const Problem& p, const BCRS& A, int fill_level, int slab_w, int cx, AMGx& H);
void amg_vcycle(AMGx& H, const double* d, double* v);

struct NewtonOpts : ax1o_newton_opts {
  NewtonOpts() { ax1o_newton_opts_default(this); }
  NewtonOpts(const ax1o_newton_opts& o) : ax1o_newton_opts(o) {}
};
struct NewtonResult : ax1o_newton_result {
  NewtonResult() : ax1o_newton_result{} {}
};
enum { NEWTON_OK = 0, NEWTON_NOT_CONVERGED = 1, NEWTON_LINSOLVE_FAILED = 2, NEWTON_DEFECT_NAN = 3,
       NEWTON_LINESEARCH_FAILED = 4 };
void membrane_flux_terms(const Problem& p, int i, const double* uglob, double* out15);
void newton(Problem& p, double* u, const NewtonOpts& o, const Comm* comm, NewtonResult& res);

// ---- mori.cpp: electroneutral (Mori) operator-split model
void mori_update_state(Problem& p, const double* u);
void mori_accept_state(Problem& p);
void mori_potential_assemble(const Problem& p, const double* u, const double* ulast, double* r, BCRS* A);
void mori_concentration_assemble(const Problem& p, const double* u, const double* ulast, double* r, BCRS* A);
void mori_step(Problem& p, double* u, const NewtonOpts& o, MoriResult& res);
double line_search(size_t n, double* u, const double* z, double* prev_u, double cur_defect, double prev_defect,
                   int maxit, const std::function<double()>& eval, std::vector<double>* trace);

}  // namespace ax1o
