/* oracle/ax1_oracle.h -- C API of the CPU oracle. TEST INFRASTRUCTURE ONLY (see oracle.hpp).
 * Loaded with ctypes from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs; never from the product path. */
#ifndef AX1_ORACLE_H
#define AX1_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ax1o_config {
  double temperature, eps_elec, length_scale, time_scale;
  double diff_water[3];
  int valence[3];
  double pot_residual_scale;
  double ymemb, dmemb, xmin_global, xmax_global, ymin, ymax;
  int do_volume_scaling;
  double volume_scaling_threshold;
  int top_dirichlet[2], bottom_dirichlet[2];
  int left_cytosol_dirichlet[2], right_cytosol_dirichlet[2];
  int left_debye_dirichlet[2], right_debye_dirichlet[2];
  int left_extra_dirichlet[2], right_extra_dirichlet[2];
  double debye_layer_width;
  int do_stimulation;
  double tinj_start, tinj_end, stim_x, stim_y, I_inj;
  int do_equilibration;
  double t_equilibrium, dummy_value;
  int membrane_active, automatic_channel_init;
  double channel_resting_potential;
  /* bit 0/1/2 = use{Elec,Memb,Time}OperatorJacobianVolume (analytic jacobian_volume; a cleared bit means
   * PDELab's NumericalJacobianVolume, the reference default); 7 = all analytic, 0 = all finite differences */
  int analytic_volume_jacobian, use_rownorm_preconditioner;
  double leak_ratio_na, leak_ratio_k;
  /* [stimulation] memb_flux_stimulation / memb_flux_ion / memb_flux_inj: cosine-ramped trans-membrane flux
   * added to one species on the interface faces above stim_x (default_nernst_planck_parameters.hh:439-475) */
  int do_memb_flux_stimulation, memb_flux_ion;
  double memb_flux_inj;
  /* electroneutral (Mori) model only: [boundary] membrane potential Neumann (isMembraneBoundaryNeumann_Potential) --
   * the membrane interface carries the ionic + capacitive current as a Neumann condition of the potential problem */
  int mori_membrane_neumann;
} ax1o_config;

typedef struct ax1o_newton_opts {
  double reduction, abs_limit, min_lin_reduction;
  int fixed_lin_reduction, maxit, line_search_maxit, lin_maxit;
  int force_iteration, disable_line_search;
  int ilu_fill, slab_w, fixed_work;
  int krylov;   /* 0 BiCGSTABSolver, 1 RestartedGMResSolver (ISTLBackend_GMRES_AMG, ax1_solverbackend.hh:196-213) */
  int restart;  /* GMRES restart length (reference: 100) */
  int amg;      /* 1: x-aggregation multigrid V(1,1) with block-ILU smoothers instead of the ILU alone */
  int amg_cx;   /* vertex columns per aggregate (<= 0: 4) */
} ax1o_newton_opts;

typedef struct ax1o_newton_result {
  int converged, iterations, lin_iterations, status, residual_evals;
  double first_defect, defect, reduction, conv_rate;
  double t_assembler, t_lin_solv, t_elapsed;
} ax1o_newton_result;

typedef struct ax1o_lin_result {
  int converged, iterations;
  double reduction, it_half;
} ax1o_lin_result;

typedef struct ax1o_problem ax1o_problem;

void ax1o_config_default(ax1o_config* c);
void ax1o_newton_opts_default(ax1o_newton_opts* o);
int ax1o_sizeof_config(void);

/* radial grid vector (returns the number of points; out may be NULL to query) */
int ax1o_generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell,
                    double dy_cell_min, int n_dy_cell_min, double dy, double dy_min, int n_dy_min,
                    double* out, int cap);

/* one GridVector primitive appended to a vector holding `start` (ops: oracle/ref_driver.cpp) */
int ax1o_gridvector(int op, double start, int n, double a, double b, double c, double* out, int cap);
/* The Newton line search alone, on a caller-supplied defect function (defect_cb sees the trial u and
   returns its defect norm, NaN allowed). Writes the trial step lengths to trace (<= cap), their
   count to *ntrace; returns the accepted defect. Used to pin the search against the reference's
   compiled Ax1Newton::line_search (tests/test_reference_pins.py). */
typedef double (*ax1o_defect_cb)(void* ctx, const double* u, int n);
double ax1o_line_search(int n, double* u, const double* z, double cur_defect, double prev_defect, int maxit,
                        ax1o_defect_cb cb, void* ctx, double* trace, int cap, int* ntrace);

ax1o_problem* ax1o_create(const ax1o_config* cfg, int nx, int ny, const double* x, const double* y,
                          int left_proc_boundary, int right_proc_boundary, const double* eps_memb,
                          const double* g_leak, const double* g_nav, const double* g_kv);
void ax1o_destroy(ax1o_problem* p);
const char* ax1o_last_error(void);

/* radial split of the ILU sub-slabs for all later factorisations of this problem: rows j < jcut and j >= jcut are
 * factorised separately (oracle.hpp: Problem::ilu_jcut); 0 = off */
void ax1o_set_ilu_split(ax1o_problem* p, int jcut);
int ax1o_num_vertices(const ax1o_problem* p);
int ax1o_membrane_row(const ax1o_problem* p);
void ax1o_get_maps(const ax1o_problem* p, unsigned char* cell_sub, double* node_volume,
                   unsigned char* dirichlet);
int ax1o_pattern(const ax1o_problem* p, int* rowptr, int* col); /* returns #blocks */
void ax1o_constants(const ax1o_problem* p, double* out /* PC, kT_e, D0, D1, D2 */);

void ax1o_set_time(ax1o_problem* p, double t, double dt);
void ax1o_set_uold(ax1o_problem* p, const double* uold);
int ax1o_update_state(ax1o_problem* p, const double* u);
void ax1o_accept_state(ax1o_problem* p);
/* serialized like DynamicChannelSet::serializeChannelStates: [ratioNa|ratioK|m|h|n] x nx, accepted
 * values; newstate != 0 returns the p_new values instead */
void ax1o_get_channel_state(const ax1o_problem* p, double* s, int newstate);
void ax1o_set_channel_state(ax1o_problem* p, const double* s, double v_rest, int initialized);
double ax1o_get_v_rest(const ax1o_problem* p);
void ax1o_membrane_potential(const ax1o_problem* p, const double* u, double* vm_mV);

/* one local kernel of cell (i, j) on given local coefficients xl[16] (comp*4 + vertex): which = 0 elec spatial
 * alpha_volume, 1 time operator, 2 membrane Poisson, 3 membrane-interface alpha_boundary (u supplies the other side) */
int ax1o_cell_kernel(const ax1o_problem* p, int i, int j, int which, const double* xl, const double* u, double weight,
                     double* out16);
/* analytic local Jacobian (16 x 16, row/column = comp*4 + vertex): which = 0 elec spatial, 1 time operator, 2 membrane */
int ax1o_cell_jacobian(const ax1o_problem* p, int i, int j, int which, const double* xl, double weight, double* out256);
/* output terms of the membrane flux per column: out[(term * 3 + ion) * nx + i], term = 0 total, 1 diffusion term,
 * 2 drift term, 3 leak channels, 4 voltage-gated channels (membranefluxgridfunction_implicit.hh:42) */
int ax1o_membrane_flux_terms(const ax1o_problem* p, const double* u, double* out);
int ax1o_residual(const ax1o_problem* p, const double* u, double* r, double* abs_terms);
int ax1o_jacobian(const ax1o_problem* p, const double* u, double* vals /* BCRS order */);
void ax1o_rownorm(const ax1o_problem* p, double* vals, double* r);
void ax1o_spmv(const ax1o_problem* p, const double* vals, const double* x, double* y);
/* ILU(fill) with sub-slab width slab_w of `vals`, then one application v = M^-1 d */
int ax1o_ilu_apply(const ax1o_problem* p, const double* vals, int ilu_fill, int slab_w,
                   const double* d, double* v);
int ax1o_ilu_num_blocks(const ax1o_problem* p, int ilu_fill, int slab_w);
/* BiCGStab(ILU) solve A z = r from z = 0; r is overwritten by the final residual */
int ax1o_solve(const ax1o_problem* p, const double* vals, double* r, double* z, double reduction,
               int maxit, int ilu_fill, int slab_w, ax1o_lin_result* res);
/* general Krylov solve: krylov 0 BiCGStab / 1 restarted GMRES; amg 0 ILU(fill) / 1 multigrid cycle */
int ax1o_solve_ex(const ax1o_problem* p, const double* vals, double* r, double* z, double reduction,
                  int maxit, int ilu_fill, int slab_w, int krylov, int restart, int amg, int amg_cx,
                  ax1o_lin_result* res);
/* one application of the multigrid cycle to d */
int ax1o_amg_apply(const ax1o_problem* p, const double* vals, int ilu_fill, int slab_w, int amg_cx,
                   const double* d, double* v);
int ax1o_newton(ax1o_problem* p, double* u, const ax1o_newton_opts* o, ax1o_newton_result* res);

/* overlapping multi-rank emulation: one thread per rank (stand-in for MPI ranks) */
int ax1o_newton_mt(ax1o_problem** ranks, int nranks, double** u, const ax1o_newton_opts* o,
                   ax1o_newton_result* res /* per rank */);
int ax1o_solve_mt(ax1o_problem** ranks, int nranks, double** vals, double** r, double** z,
                  double reduction, int maxit, int ilu_fill, int slab_w, ax1o_lin_result* res);

/* ---- electroneutral (Mori) operator-split model, oracle/mori.cpp ---- */
typedef struct ax1o_mori_result {
  int status, pot_iterations, con_iterations;
  double pot_first_defect, pot_defect, con_first_defect, con_defect;
} ax1o_mori_result;
/* which = 0: potential sub-problem, 1: concentration sub-problem. r (4 per vertex, the other sub-problem's rows 0) and,
 * if vals != NULL, the Jacobian in BCRS order (identity on the other sub-problem's rows) at u; ulast = the solution of
 * the last iteration the split lags (concentrations of the potential problem; drift / membrane terms of the
 * concentration problem). uold / time / dt / channel state as set on the problem. */
int ax1o_mori_assemble(ax1o_problem* p, int which, const double* u, const double* ulast, double* r, double* vals);
/* charge-layer states gamma: out[((side * 2 + new) * 3 + ion) * nx + column], side 0 = cytosol (int), 1 = ES (ext) */
void ax1o_mori_get_gamma(const ax1o_problem* p, double* out);
void ax1o_mori_update_state(ax1o_problem* p, const double* u);   /* ChargeLayerMoriImplicit::updateState */
/* one operator-split time step (potential solve, gamma update, concentration solve) from u = uold; uses
 * o->reduction, abs_limit, lin_maxit, ilu_fill, slab_w. ax1o_accept_state also accepts the gamma states. */
int ax1o_mori_step(ax1o_problem* p, double* u, const ax1o_newton_opts* o, ax1o_mori_result* res);

/* external collectives (e.g. torch.distributed/gloo from Python): callbacks for one rank */
typedef void (*ax1o_halo_add_fn)(void* ctx, int rank, double* v);
typedef double (*ax1o_allreduce_fn)(void* ctx, int rank, double local);
int ax1o_newton_ext(ax1o_problem* p, double* u, const ax1o_newton_opts* o, ax1o_newton_result* res,
                    int rank, ax1o_halo_add_fn halo_add, ax1o_allreduce_fn allreduce, void* ctx);

#ifdef __cplusplus
}
#endif
#endif
