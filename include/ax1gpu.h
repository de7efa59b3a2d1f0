/* include/ax1gpu.h -- C ABI of libax1gpu.so: the B200-native PNP Newton hot path of dune-ax1.
 *
 * The reference (jurgispods/dune-ax1) has no FFI boundary; the seam this library replaces is the
 * set of C++ template concepts instantiated in Acme2CylSetup::setup
 * (dune/ax1/acme2_cyl/common/acme2_cyl_setup.hh:226-1018):
 *   GridOperator  IGO::residual(x,r) / IGO::jacobian(x,A)     acme2_cyl_setup.hh:700-708
 *   Solver backend LS::apply(A,z,r,reduction) / norm / result  acme2_cyl_setup.hh:727-811
 *   PDE solver    Ax1Newton::apply(u), result()                acme2_cyl_setup.hh:822-854,
 *                                                              dune/ax1/common/ax1_newton.hh:541-617
 *   membrane flux gfMembFlux.updateState / acceptTimeStep      dune/ax1/common/
 *                                                              membranefluxgridfunction_implicit.hh:402-546
 * A header-only C++ shim that models those concepts on top of this ABI is in
 * include/ax1gpu_pdelab_shim.hh; INTEGRATION.md shows the lines a maintainer changes.
 *
 * Conventions: every function returns 0 on success, a negative AX1GPU_E* code on failure and
 * never throws across the boundary; ax1gpu_last_error() gives the message. Vectors are
 * vertex-blocked [Na,K,Cl,pot], vertices x-fastest (Dune::BlockVector<FieldVector<double,4>>,
 * istl::raw(v)); the Jacobian is block-CSR with row-major 4x4 blocks, columns ascending
 * (Dune::BCRSMatrix<FieldMatrix<double,4,4>>). All pointers are HOST pointers unless the
 * function name ends in _dev. There is no CPU fallback: without a CUDA device every compute
 * entry point fails with AX1GPU_ENODEVICE.
 */
#ifndef AX1GPU_H
#define AX1GPU_H
#ifdef __cplusplus
extern "C" {
#endif

#define AX1GPU_OK 0
#define AX1GPU_EINVAL (-1)
#define AX1GPU_ENODEVICE (-2)
#define AX1GPU_ECUDA (-3)
#define AX1GPU_ENCCL (-4)
#define AX1GPU_ESTATE (-5)

/* Newton status codes; the shim re-throws the matching Dune::PDELab::NewtonError subclass
 * (caught by the time loop, acme2_cyl_simulation.hh:593-692). */
#define AX1GPU_NEWTON_OK 0
#define AX1GPU_NEWTON_NOT_CONVERGED 1   /* NewtonNotConverged      */
#define AX1GPU_NEWTON_LINSOLVE_FAILED 2 /* NewtonLinearSolverError */
#define AX1GPU_NEWTON_DEFECT_NAN 3      /* NewtonDefectError       */
#define AX1GPU_NEWTON_LINESEARCH_FAILED 4

typedef struct ax1gpu_ctx* ax1gpu_handle;

/* Local tensor grid of this rank: interior + one overlap cell column per interior side
 * (dune/ax1/common/ax1_gridgenerator.hh:601-631). */
typedef struct ax1gpu_grid {
  int nx, ny;          /* local cells */
  const double* x;     /* nx+1 axial coordinates [nm]  */
  const double* y;     /* ny+1 radial coordinates [nm] */
  int left_proc_boundary, right_proc_boundary; /* side is a processor boundary (overlap) */
} ax1gpu_grid;

/* Physical constants and run-time switches (Physics / Acme2CylParameters getters used on the path). */
typedef struct ax1gpu_params {
  double temperature;        /* electrolyte.hh: 279.45 K                                        */
  double eps_elec;           /* default_config.hh:71  (80)                                      */
  double length_scale, time_scale; /* default_config.hh:113-116                                 */
  double diff_water[3];      /* default_config.hh:119-123 [m^2/s]                               */
  int valence[3];            /* default_config.hh:77,83,89                                      */
  double pot_residual_scale; /* acme2_cyl_parametertree.hh:300-303                              */
  int do_volume_scaling;     /* use node_volume (acme2_cyl_operator_fully_coupled.hh:142-156)   */
  int do_stimulation;        /* default_nernst_planck_parameters.hh:168-216                     */
  double tinj_start, tinj_end, stim_x, stim_y, I_inj;
  int do_equilibration;      /* membranefluxgridfunction_implicit.hh:223-265                    */
  double t_equilibrium, dummy_value;
  int membrane_active, automatic_channel_init;
  double channel_resting_potential;
  int use_rownorm_preconditioner; /* rownorm_preconditioner.hh:65-155                           */
  double leak_ratio_na, leak_ratio_k; /* [equilibration] Na_leak / K_leak                       */
  /* [general] useElecOperatorJacobianVolume / useMembOperatorJacobianVolume / useTimeOperatorJacobianVolume
   * (acme2_cyl_operator_fully_coupled.hh:79, convectiondiffusionfem.hh:78, acme2_cyl_toperator.hh:51):
   * 1 = the operator's analytic jacobian_volume, 0 = PDELab's NumericalJacobianVolume (forward difference,
   * delta = 1e-7 (1 + |u_j|)), which is the reference's default. ax1gpu_params_default sets all three to 1. */
  int use_elec_jacobian_volume, use_memb_jacobian_volume, use_time_jacobian_volume;
  /* [stimulation] memb_flux_stimulation, memb_flux_ion (0 Na, 1 K, 2 Cl; default Cl), memb_flux_inj: cosine-
   * ramped flux 0.5 (1 - cos(2 pi (t - tinj_start) / tinj_end)) memb_flux_inj / |face| added to one species on
   * the membrane-interface faces above stim_x (default_nernst_planck_parameters.hh:439-475)              */
  int do_memb_flux_stimulation, memb_flux_ion;
  double memb_flux_inj;
  /* electroneutral (Mori) model only: [boundary] membrane potential Neumann (isMembraneBoundaryNeumann_Potential):
   * the membrane interface carries the ionic + capacitive current as Neumann condition of the potential problem */
  int mori_membrane_neumann;
} ax1gpu_params;

typedef struct ax1gpu_newton_opts {
  double reduction, abs_limit, min_lin_reduction; /* acme2_cyl_setup.hh:828-833 */
  int fixed_lin_reduction, maxit, line_search_maxit, lin_maxit;
  int force_iteration, disable_line_search;
  int slab_w;      /* vertex columns per ILU0 sub-slab (<=0: library default)                  */
  int fixed_work;  /* benchmark mode: exactly maxit Newton x lin_maxit BiCGStab iterations     */
  int ilu_fill;    /* n of SeqILUn: 0 = block ILU0, 1 = block ILU(1) (reference default,
                      acme2_cyl_setup.hh:764-765: ISTLBackend_OVLP_BCGS_ILUn(multigfs,cc,1,...))  */
  int amg;         /* 1: multigrid preconditioner (x-aggregation AMG, V(1,1), block-ILU smoothers) instead of
                      the plain ILU -- the device counterpart of ISTLBackend_GMRES_AMG_ILU0's Dune::Amg::AMG
                      (ax1_solverbackend.hh:73-302, selected by the myelin build acme2_cyl_setup.hh:749-760)   */
  int krylov;      /* 0: Dune::BiCGSTABSolver (default backends), 1: Dune::RestartedGMResSolver as in
                      ISTLBackend_GMRES_AMG::apply (ax1_solverbackend.hh:196-213)                               */
  int restart;     /* GMRES restart length (reference: 100, ax1_solverbackend.hh:201)                          */
} ax1gpu_newton_opts;

typedef struct ax1gpu_newton_result { /* PDELab NewtonResult + ax1_newton.hh:600-607 */
  int converged, iterations, lin_iterations, status, residual_evals;
  double first_defect, defect, reduction, conv_rate;
  double t_assembler, t_lin_solv, t_elapsed; /* seconds, device time from CUDA events */
} ax1gpu_newton_result;

/* result of one operator-split step of the electroneutral (Mori) model: the two Ax1LinearSubProblemSolver results */
typedef struct ax1gpu_mori_result {
  int status, pot_iterations, con_iterations;
  double pot_first_defect, pot_defect, con_first_defect, con_defect;
} ax1gpu_mori_result;

typedef struct ax1gpu_lin_result { /* Dune::InverseOperatorResult */
  int converged, iterations;
  double reduction, it_half;
} ax1gpu_lin_result;

const char* ax1gpu_last_error(void);
void ax1gpu_params_default(ax1gpu_params* p);
void ax1gpu_newton_opts_default(ax1gpu_newton_opts* o);
int ax1gpu_sizeof_params(void);
int ax1gpu_device_count(void);

/* Create the device-resident problem. Arrays are the Physics maps of the reference
 * (acme2_cyl_physics.hh): cell_subdomain nx*ny {0 CYTOSOL,1 ES,2 MEMBRANE}; memb_perm nx
 * (initPermittivities :2424-2606); node_volume (nx+1)*(ny+1) (calcNodeVolumes :2286-2397);
 * dirichlet_mask (nx+1)*(ny+1), bit c = component c constrained (setup.hh:646-665; rows become
 * unit rows, columns are kept as in PDELab's default non-symmetric constraints handling);
 * g_leak/g_nav/g_kv nx channel densities per membrane column [mS/cm^2] (channelset.hh:133-182).
 * stream: a cudaStream_t to run on (NULL: the library creates one). */
int ax1gpu_create(const ax1gpu_grid* grid, const unsigned char* cell_subdomain, const double* memb_perm,
                  const double* node_volume, const unsigned char* dirichlet_mask, const ax1gpu_params* params,
                  const double* g_leak, const double* g_nav, const double* g_kv, int device, void* stream,
                  ax1gpu_handle* out);
int ax1gpu_destroy(ax1gpu_handle h);

/* BCRS sparsity pattern (FullVolumePattern on Q1 vertices). Pass NULLs to query *nblocks. */
int ax1gpu_pattern(ax1gpu_handle h, int* rowptr, int* colidx, int* nblocks);

/* OneStepMethod::apply(time,dt,uold,unew): flux GF time = t, LOP time = t+dt; builds r0=-M(uold). */
int ax1gpu_set_time(ax1gpu_handle h, double t, double dt);
int ax1gpu_set_uold(ax1gpu_handle h, const double* uold);
/* gfMembFlux.updateState() / acceptTimeStep(); channel state layout =
 * DynamicChannelSet::serializeChannelStates: [ratioNa|ratioK|m|h|n] each nx long. */
int ax1gpu_update_state(ax1gpu_handle h, const double* u);
int ax1gpu_accept_state(ax1gpu_handle h);
int ax1gpu_get_channel_state(ax1gpu_handle h, double* s, int newstate, double* v_rest);
int ax1gpu_set_channel_state(ax1gpu_handle h, const double* s, double v_rest, int initialized);
int ax1gpu_membrane_potential(ax1gpu_handle h, const double* u, double* vm_mV);

/* IGO::residual + LS::norm. u == NULL: use the device-resident iterate. r/norm may be NULL. */
int ax1gpu_residual(ax1gpu_handle h, const double* u, double* r, double* norm);
/* IGO::jacobian; blocks (BCRS order, 16 doubles per block) may be NULL (stay on device). */
int ax1gpu_jacobian(ax1gpu_handle h, const double* u, double* blocks);
/* overwrite the device Jacobian with host BCRS values (used by the shim when A was modified on host) */
int ax1gpu_set_jacobian(ax1gpu_handle h, const double* blocks);
int ax1gpu_get_jacobian(ax1gpu_handle h, double* blocks);
/* LS::apply(A,z,r,reduction): block-ILU0(sub-slab) BiCGStab on the device Jacobian.
 * r == NULL: use the device residual of the last ax1gpu_residual. r (if given) is overwritten by
 * the final linear residual as ISTL does. */
int ax1gpu_solve(ax1gpu_handle h, double* r, double reduction, int maxit, int slab_w, double* z,
                 ax1gpu_lin_result* res);
/* LS::apply of ISTLBackend_GMRES_AMG (ax1_solverbackend.hh:166-213): left-preconditioned restarted GMRES
 * (modified Gram-Schmidt, Givens rotations, convergence on the preconditioned residual) with the handle's
 * current preconditioner (ILU, or the multigrid cycle after ax1gpu_set_amg); arguments as ax1gpu_solve. */
int ax1gpu_solve_gmres(ax1gpu_handle h, double* r, double reduction, int restart, int maxit, int slab_w, double* z,
                       ax1gpu_lin_result* res);
/* Ax1Newton::apply(u): u in/out. */
int ax1gpu_newton(ax1gpu_handle h, const ax1gpu_newton_opts* opts, double* u, ax1gpu_newton_result* res);

/* ---- device-resident variants (value metric: no host<->device traffic in the call) ---- */
int ax1gpu_upload_u(ax1gpu_handle h, const double* u);
int ax1gpu_download_u(ax1gpu_handle h, double* u);
int ax1gpu_set_uold_dev(ax1gpu_handle h); /* uold := device iterate */
int ax1gpu_newton_dev(ax1gpu_handle h, const ax1gpu_newton_opts* opts, ax1gpu_newton_result* res);

/* ---- the caller of the path with all state resident on the device ----
 * One adaptive time step of Acme2CylSimulation::run (acme2_cyl_simulation.hh:255-439 dt control, :457/:589/
 * :916-927 updateState -> solver.apply -> uold = unew -> acceptTimeStep, :593-638 halve-and-retry). Times in
 * the units of the grid operator (micro seconds with the default time_scale). */
typedef struct ax1gpu_timeloop_opts {
  double dtstart;                 /* command line dt0; re-used when the stimulation starts (:262-266)        */
  double max_dt, min_dt;          /* [timeStepping] maxTimeStep / minTimeStep, already divided by time_scale */
  double max_dt_ap;               /* maxTimeStepAP / time_scale (:378-383)                                   */
  double vm_ap_threshold;         /* -50 mV                                                                  */
  int lower_limit, upper_limit;   /* timeStepLowerLimitNewtonIt (3) / timeStepUpperLimitNewtonIt (5)         */
  int max_restarts;               /* maxNumberNewtonRestarts                                                 */
  int adaptive;                   /* useAdaptiveTimeStep                                                     */
  int restart_from_uold;          /* 0 (reference, :563-566): a retry starts from the iterate the failed Newton solve
                                     left behind, except after a NaN defect; 1: every retry starts from uold       */
} ax1gpu_timeloop_opts;
typedef struct ax1gpu_timeloop_state { /* carried by the caller between steps */
  double time, dt;
  int its, last_its, steps;
} ax1gpu_timeloop_state;
typedef struct ax1gpu_step_result {
  double time, dt;                /* time reached and the step that was accepted                             */
  int newton_iterations, lin_iterations, restarts;
  int newton_status;              /* != 0: the step failed max_restarts + 1 times; iterate reset to uold     */
  double vm_min, vm_max;          /* membrane potential extrema [mV] over all ranks after the step           */
  double defect, first_defect, t_assembler, t_lin_solv;
} ax1gpu_step_result;
void ax1gpu_timeloop_opts_default(ax1gpu_timeloop_opts* o);
/* advances the device-resident iterate (ax1gpu_upload_u / ax1gpu_download_u) by one time step */
int ax1gpu_time_step(ax1gpu_handle h, const ax1gpu_timeloop_opts* opts, const ax1gpu_newton_opts* newton,
                     ax1gpu_timeloop_state* state, ax1gpu_step_result* res);
/* min / max of the membrane potential [mV] of the device iterate (all ranks) */
int ax1gpu_membrane_potential_range(ax1gpu_handle h, double* vm_min, double* vm_max);
/* membrane potential [mV] per membrane column of the device iterate */
int ax1gpu_membrane_potential_dev(ax1gpu_handle h, double* vm_mV);

/* Output terms of the trans-membrane flux per ion and membrane column of the iterate u (u == NULL: the device
 * iterate), as MembraneFluxGridFunctionImplicit::evaluate returns them for FluxTerms TOTAL_FLUX, DIFF_TERM,
 * DRIFT_TERM, LEAK, VOLTAGE_GATED (membranefluxgridfunction_implicit.hh:42, 83-361) on the cytosol-side interface
 * (outer normal +y), with the new channel state: what acme2_cyl_output.hh:1215-1410 writes as memb_flux* per ion.
 * terms[(term * 3 + ion) * nx + column], 15 * nx doubles. */
int ax1gpu_membrane_flux_terms(ax1gpu_handle h, const double* u, double* terms);

/* ---- electroneutral (Mori) operator-split model (BASELINE configs[2]; csrc/mori.cu) -------------------------------
 * replaces, for that model, Acme2CylMoriSimulation::run's step body (acme2_cyl_mori_simulation.hh:562-600):
 *   solverPot.apply(unew)              Ax1LinearSubProblemSolver over Acme2CylMoriPotentialOperator
 *   gfMoriFlux.updateState(time, dt)   ChargeLayerMoriImplicit (charge_layer_mori_implicit.hh)
 *   solverCon.apply(time,dt,uold,unew) OneStepMethod / implicit Euler over Acme2CylMoriConcentrationOperator
 * The caller drives it like the coupled path: ax1gpu_set_time, ax1gpu_upload_u + ax1gpu_set_uold_dev (or
 * ax1gpu_set_uold), ax1gpu_update_state (channel gating), ax1gpu_mori_step, ax1gpu_accept_state (which also accepts
 * the charge-layer states). Uses opts->reduction, abs_limit (min_defect), lin_maxit, slab_w, ilu_fill. u != NULL:
 * host iterate in / out; NULL: the device iterate. Single rank. */
int ax1gpu_mori_step(ax1gpu_handle h, const ax1gpu_newton_opts* opts, double* u, ax1gpu_mori_result* res);
/* residual (4 per vertex; the other sub-problem's rows are 0) and optionally the Jacobian (BCRS order, identity on the
 * other sub-problem's rows) of sub-problem `which` (0 potential, 1 concentrations) at u, with ulast = the solution of
 * the last iteration the split lags; the device Jacobian is left in place for ax1gpu_solve */
int ax1gpu_mori_assemble(ax1gpu_handle h, int which, const double* u, const double* ulast, double* r, double* blocks);
int ax1gpu_mori_update_state(ax1gpu_handle h, const double* u);   /* gfMoriFlux.updateState on the host iterate u */
/* charge-layer states gamma[((side * 2 + new) * 3 + ion) * nx + column], side 0 cytosol / 1 ES: 12 * nx doubles */
int ax1gpu_mori_get_gamma(ax1gpu_handle h, double* gamma);

/* ---- building blocks exposed for parity tests and profiling ---- */
int ax1gpu_spmv(ax1gpu_handle h, const double* x, double* y);
int ax1gpu_ilu_factor(ax1gpu_handle h, int slab_w);   /* keeps the handle's current fill level (0 at creation) */
/* SeqILUn(A, n, 1.0): fill = 0 or 1; also selects the preconditioner of later ax1gpu_solve calls */
int ax1gpu_ilu_factor_n(ax1gpu_handle h, int slab_w, int fill);
/* switch the preconditioner of ax1gpu_ilu_factor / ax1gpu_ilu_apply / ax1gpu_solve between the ILU alone
 * (on = 0) and one V(1,1) cycle of the aggregation multigrid with `cx` vertex columns per aggregate
 * (cx <= 0: keep, default 4); the hierarchy is rebuilt by the next factorisation */
int ax1gpu_set_amg(ax1gpu_handle h, int on, int cx);
/* Radial split of the block-ILU0 sub-slabs (additive Schwarz in y as well as in x): rows below and from grid row
 * `jcut` on are factorised and swept separately, couplings across the cut are dropped from the preconditioner.
 * jcut = -1 (default): automatic -- when the library also chooses the sub-slab width (slab_w <= 0) and that width is
 * one warp's (grids of up to 9,472 vertex columns: a sweep costs its dependency depth there), it cuts in the uniform
 * far field of the radial grid, which halves the depth of a sweep at no Krylov iterations on the paper grid;
 * 0: never; > 0: cut at this row. Takes effect with the next factorisation. *active (optional) returns the cut the
 * last factorisation used. */
int ax1gpu_set_ilu_split(ax1gpu_handle h, int jcut);
int ax1gpu_get_ilu_split(ax1gpu_handle h, int* active);
/* Operand of the solver's matrix-vector products: mode -1 (default) automatic -- grids of >= 2^20 vertices (HBM-bound)
 * multiply with a copy of the Jacobian in 64-byte blocks (the potential row's three concentration entries are one
 * number times the valences, acme2_cyl_operator_fully_coupled.hh:541-547, so 8 doubles instead of 10 carry a block),
 * built by each factorisation and bit-checked against the Jacobian: a matrix whose potential rows do not have that
 * shape (finite-difference Jacobians) is multiplied as it is; 0: never; 1: on every grid. *active (optional) returns
 * whether the last factorisation's copy is in use. */
int ax1gpu_set_spmv_compact(ax1gpu_handle h, int mode);
int ax1gpu_get_spmv_compact(ax1gpu_handle h, int* active);
/* one application of the current preconditioner (ILU sweep or AMG cycle) incl. the halo sum */
int ax1gpu_ilu_apply(ax1gpu_handle h, const double* d, double* v);
int ax1gpu_rownorm(ax1gpu_handle h, double* r); /* scales device Jacobian and r */
/* run `reps` launches of one named kernel on device-resident data and return the mean device time
 * [ms] from CUDA events ("residual","jacobian","spmv","spmv_dot1","spmv_dot2" = the two dot-fused SpMV variants
 * BiCGStab launches,"ilu_factor","ilu_apply","dot","axpy","pupdate") */
int ax1gpu_time_kernel(ax1gpu_handle h, const char* name, int reps, double* ms);
long long ax1gpu_launch_count(ax1gpu_handle h); /* kernels launched so far by this handle */
void* ax1gpu_stream(ax1gpu_handle h);

/* ---- multi-GPU (z-slabs along the axon, one process per GPU, NCCL) ----
 * Replaces Dune::CollectiveCommunication / YaspGrid::communicate on the path
 * (AddDataHandle halo sum + allreduce, ax1_solverbackend.hh:346-381,413-437). */
int ax1gpu_nccl_unique_id(const char* libnccl_path, char* id128);
int ax1gpu_comm_init(ax1gpu_handle h, const char* libnccl_path, const char* id128, int rank, int nranks);
/* how the two exchange steps run: 0 single rank, 1 NCCL send/recv + allreduce, 2 direct stores into the
 * neighbours' memory over NVLink (CUDA IPC peer windows; default when every rank can map every peer,
 * AX1_P2P=0 forces mode 1) */
/* Peer windows without NCCL (the data plane never needs it): ax1gpu_p2p_export allocates this rank's window and
 * returns its CUDA IPC handle (64 bytes); the caller gathers the handles of all ranks (MPI_Allgather in a DUNE build)
 * and hands the nranks x 64 bytes to ax1gpu_p2p_attach on every rank. Equivalent to ax1gpu_comm_init with AX1_P2P=1
 * except that the time loop's V_m min/max over ranks (ncclMax) is not available. Ranks may share one GPU (tests). */
int ax1gpu_p2p_export(ax1gpu_handle h, int rank, int nranks, char* handle64);
int ax1gpu_p2p_attach(ax1gpu_handle h, const char* all_handles);
int ax1gpu_comm_mode(ax1gpu_handle h);

/* ---- host model: restatement of the reference's host-side Physics precomputations, so that the
 * library can be driven without DUNE (bench, tests, examples). No GPU needed. ---- */
typedef struct ax1host_bc { /* [boundary.concentration] / [boundary.potential]; index 0 con, 1 pot */
  int top[2], bottom[2], left_cytosol[2], right_cytosol[2], left_debye[2], right_debye[2],
      left_extra[2], right_extra[2];
  double debye_layer_width;
} ax1host_bc;
int ax1host_generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell,
                       double dy_cell_min, int n_dy_cell_min, double dy, double dy_min, int n_dy_min,
                       double* out, int cap);
int ax1host_physics_maps(const ax1gpu_grid* grid, double ymemb, double dmemb, double xmin_global,
                         double xmax_global, double ymin, double ymax, int do_volume_scaling,
                         double volume_scaling_threshold, const ax1host_bc* bc,
                         unsigned char* cell_subdomain, double* node_volume, unsigned char* dirichlet_mask);
/* z-slab partition of nx_global cell columns over nranks (Ax1YaspPartition, px = nranks, py = 1):
 * interior cells [c0,c1) and local cells [l0,l1) incl. one overlap column per interior side. */
int ax1host_slab(int nx_global, int nranks, int rank, int* c0, int* c1, int* l0, int* l1);

/* ---- membrane groups ([membrane.<group>] sections: myelin / nodes of Ranvier) ---- */
typedef struct ax1host_membrane_group {
  double start, width, stride;   /* -1: key absent (a group without position is only used as default group) */
  double dx;                     /* <= 0: [general] dx                                                      */
  int smooth_transition, geometric_refinement;
  double dx_transition;          /* <= 0: min(width of this interval, width of the neighbouring interval)   */
  int n_transition;              /* default 10                                                              */
  double permittivity, d_memb;   /* <= 0: membrane default (2.0) / [general] d_memb                         */
  double g_leak, g_nav, g_kv;    /* channel densities [mS/cm^2]                                             */
} ax1host_membrane_group;
/* Ax1GridGenerator::setupXPartition (ax1_gridgenerator.hh:1238-1312); returns the number of intervals */
int ax1host_group_partition(const ax1host_membrane_group* groups, int ngroups, int default_group, double xmax,
                            int* ival_group, double* ival_start, double* ival_end, int cap);
/* Ax1GridGenerator::intervalVectorToGridVector (:1021-1234); returns the number of x points */
int ax1host_generate_x_groups(const ax1host_membrane_group* groups, int ngroups, int nival, const int* ival_group,
                              const double* ival_start, const double* ival_end, double dx_default, double xmax,
                              double* out, int cap);
/* group index (:868-895), channel densities and Physics::initPermittivities (acme2_cyl_physics.hh:2424-2606)
 * for the nx membrane cell columns of the local grid x[0..nx] */
int ax1host_membrane_arrays(const ax1host_membrane_group* groups, int ngroups, int nival, const int* ival_group,
                            const double* ival_start, const double* ival_end, int smooth_permittivities,
                            double default_permittivity, double default_dmemb, int nx, const double* x,
                            double* eps_memb, double* g_leak, double* g_nav, double* g_kv, int* cell_group);
/* processor-boundary guard of makeMultidomainGrid (:903-942) */
int ax1host_check_processor_boundary(const ax1host_membrane_group* groups, int nival, const int* ival_group,
                                     const double* ival_start, const double* ival_end, int forbidden_group,
                                     double dx_default, double xmax_rank, double xmax_global);

/* ---- simulation-state files ("next" row 8(f)1): Ax1SimulationState ASCII format
 * (dune/ax1/common/ax1_simulationstate.hh:30-57, acme2_cyl_output.hh:937-979) and the x-extrapolation
 * of a loaded state onto the simulation grid (doExtrapolateXData, acme2_cyl_simloader.hh:364). ---- */
typedef struct ax1host_state ax1host_state;
int ax1host_state_load(const char* path, ax1host_state** out);
void ax1host_state_free(ax1host_state* s);
int ax1host_state_info(const ax1host_state* s, int* nelements, double* time, double* dt, int* timestep,
                       int* mpi_rank, int* mpi_size, int* nvectors);
int ax1host_state_vector(const ax1host_state* s, const char* name, double* out, int cap); /* returns length */
int ax1host_state_save(const char* path, int nelements, int timestep, double time, double dt, int mpi_rank,
                       int mpi_size, int nvec, const char* const* names, const double* const* data, const int* sizes);
int ax1host_state_extrapolate(int nxs, const double* xs, int nys, const double* ys, const double* us, int nx,
                              const double* x, int ny, const double* y, double* u);

#ifdef __cplusplus
}
#endif
#endif
