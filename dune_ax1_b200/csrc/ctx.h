// ctx.h -- the device-resident problem behind an ax1gpu_handle.
#pragma once
#include "kernels.h"
#include "nccl_dyn.h"

struct ax1gpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  ax1::Dev g{};
  ax1gpu_params prm{};
  // host copies
  std::vector<double> hx, hy;
  std::vector<int> h_rowptr, h_col;
  int nblocks = 0;
  bool left_pb = false, right_pb = false;
  int own_lo = 0, own_hi = 0;
  // time
  double time = 0.0, dt = 0.0;
  // device arrays
  double *d_x = nullptr, *d_y = nullptr, *d_nodevol = nullptr, *d_epsm = nullptr;
  uint8_t* d_dmask = nullptr;
  int* d_rowptr = nullptr;
  double *d_u = nullptr, *d_uold = nullptr, *d_r0 = nullptr, *d_r = nullptr, *d_z = nullptr, *d_uprev = nullptr;
  double *d_A = nullptr, *d_LU = nullptr;
  double* d_Ac = nullptr;      // 64-byte-block copy of d_A for the solver's SpMV (driver.cu: drv_factor), lazily allocated
  bool ac_valid = false;
  int spmv_compact = -1;       // -1 automatic (nv >= 2^20), 0 never, 1 always (ax1gpu_set_spmv_compact)
  double *d_rt = nullptr, *d_p = nullptr, *d_v = nullptr, *d_y2 = nullptr, *d_t = nullptr, *d_tmp = nullptr;
  double* d_y1 = nullptr;      // y = M^-1 p of the current BiCGStab iteration (its x-update is applied with the second one)
  double* d_bcrs = nullptr;  // staging for Jacobian up/download (allocated lazily)
  // channels
  ax1::ChannelDev ch{};
  double *d_gbar = nullptr, *d_chstate = nullptr, *d_geff = nullptr, *d_vrest = nullptr, *d_vm = nullptr;
  double* d_fluxterms = nullptr;  // 15 x nx output terms of the membrane flux (allocated on first use)
  int* d_err = nullptr;
  // electroneutral (Mori) model: charge-layer states [side][old/new][ion][nx] (mori.cu), allocated on first use
  double* d_mori_gamma = nullptr;
  bool mori_initialized = false, mori_partially_initialized = false;
  bool channels_initialized = false, channels_partially_initialized = false;
  // reductions / Krylov scalars
  double* d_partials = nullptr;   // 2 * nred_blocks
  double* d_spartials = nullptr;  // 2 * spmv blocks (dot products fused into the SpMV)
  double* d_sums = nullptr;       // [4]
  ax1::Scalars* d_sc = nullptr;
  ax1::Scalars* h_sc = nullptr;   // pinned
  double* h_pin = nullptr;        // pinned scratch (8 doubles)
  int* h_err = nullptr;           // pinned copy of d_err[0..3]
  int nred_blocks = 0;
  // ILU
  int slab_w = 0;
  bool slab_auto = false;      // slab_w was chosen by the library (re-derived when fill or multigrid change)
  int ilu_fill = 0;            // 0: block ILU0 (solver.cu), 1: block ILU(1), the reference default (ilu1.cu)
  int lvl_slab_w = 0;          // slab width / fill level the level tables were built for
  int lvl_fill = -1;
  int lvl_jc = -1;
  int ilu_split = -1;          // radial split of the ILU0 sub-slabs: -1 automatic, 0 off, > 0 cut at this grid row
  int ilu_jc = 0;              // the cut in effect (0: none), chosen by select_preconditioner
  int* d_lvl = nullptr;        // level prefix tables (solver.cu / ilu1.cu)
  size_t lvl_cap = 0;          // ints allocated for d_lvl
  size_t lu_cap = 0;           // doubles allocated for d_LU (144 per vertex for ILU0, 208 for ILU(1))
  bool factored = false;
  // aggregation multigrid on top of the ILU smoother (amg.cu); level 0 = (g, d_A, d_LU)
  struct AmgLevel {
    ax1::Dev g{};
    int slab_w = 0;
    double *A = nullptr, *LU = nullptr, *b = nullptr, *x = nullptr;
    int* lvl = nullptr;
  };
  int amg = 0;                 // 0: ILU only, 1: V(1,1) cycle of x-aggregation AMG with block-ILU smoothers
  int amg_cx = 4;              // vertex columns per aggregate
  int amg_slab_w = 0;          // level-0 slab width the hierarchy was built for
  std::vector<AmgLevel> amg_levels;   // coarse levels 1..L-1
  double* d_amg_t = nullptr;   // level-0 scratch (defect / smoother output)
  // restarted GMRES (driver.cu: drv_solve_gmres), allocated by the first GMRES solve
  double* d_gmV = nullptr;     // Krylov basis, gm_cap + 1 vectors
  double* d_gm = nullptr;      // Hessenberg matrix, rotations, scalars
  int gm_cap = 0;              // restart length the buffers hold
  // multi-GPU
  ax1::NcclApi nccl;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  double *d_halo_send = nullptr, *d_halo_recv = nullptr;  // 2 sides x 3 columns x nvy x 4
  // peer-memory windows over NVLink (p2p.cuh); off: the NCCL send/recv + allreduce path is used
  struct {
    bool on = false;
    double* win = nullptr;                 // own window (cudaMalloc, exported with cudaIpcGetMemHandle)
    double* peer[AX1_MAX_RANKS] = {};      // mapped windows of all ranks, peer[rank] = win
    unsigned int* cnt = nullptr;           // local last-CTA counters [2]
    unsigned long long halo_seq = 0, red_seq = 0;
    size_t halo_len = 0;                   // doubles per halo slot = 3 * nvy * 4
    long long timeout_cycles = 0;          // wait limit of one exchange in SM cycles (AX1_P2P_TIMEOUT_S, default 600 s)
    bool broken = false;                   // an exchange timed out: sequence counters disagree, handle unusable
  } p2p;
  // accounting
  long long launches = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

namespace ax1 {
// driver.cu
int drv_refresh_dev(ax1gpu_ctx* c);
int drv_set_uold(ax1gpu_ctx* c);                       // r0 = -M(d_uold)
int drv_update_state(ax1gpu_ctx* c, const double* d_u);
int drv_accept_state(ax1gpu_ctx* c);
int drv_residual(ax1gpu_ctx* c, const double* d_u, double* d_r, double* norm_host);
int drv_jacobian(ax1gpu_ctx* c, const double* d_u);
int drv_factor(ax1gpu_ctx* c, int slab_w, int fill = -1, int amg = -1);  // < 0: keep the handle's setting
int drv_prec_apply(ax1gpu_ctx* c, const double* d, double* v);
void drv_amg_free(ax1gpu_ctx* c);
int drv_op_apply(ax1gpu_ctx* c, const double* x, double* y);
int drv_solve(ax1gpu_ctx* c, double* d_rhs, double* d_zout, double reduction, int maxit, ax1gpu_lin_result* res);
int drv_solve_gmres(ax1gpu_ctx* c, double* d_rhs, double* d_zout, double reduction, int restart, int maxit,
                    ax1gpu_lin_result* res);
int drv_newton(ax1gpu_ctx* c, const ax1gpu_newton_opts* o, ax1gpu_newton_result* res);
// timeloop.cu
int drv_vm_minmax(ax1gpu_ctx* c, double* vmin, double* vmax);
int drv_time_step(ax1gpu_ctx* c, const ax1gpu_timeloop_opts* o, const ax1gpu_newton_opts* no, ax1gpu_timeloop_state* s,
                  ax1gpu_step_result* out);
int drv_norm(ax1gpu_ctx* c, const double* a, double* out_host);
int drv_time_kernel(ax1gpu_ctx* c, const char* name, int reps, double* ms);
int drv_check_p2p(ax1gpu_ctx* c);
// mori.cu
int drv_mori_update_state(ax1gpu_ctx* c, const double* d_u);
int drv_mori_accept_state(ax1gpu_ctx* c);
int drv_mori_assemble(ax1gpu_ctx* c, int which, const double* d_u, const double* d_ulast);   // -> d_r, d_A
int drv_mori_step(ax1gpu_ctx* c, const ax1gpu_newton_opts* o, ax1gpu_mori_result* res);   // after a host sync: AX1GPU_ENCCL if a peer-window exchange timed out
}  // namespace ax1
