// kernels.h -- launch wrappers shared between the translation units of libax1gpu.
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace ax1 {

struct ChannelDev {
  const double* gbar[3];  // leak, Nav, Kv
  double* ratio[2];       // Na_leak, K_leak
  double* gate[3];        // m, h, n accepted
  double* gate_new[3];
  double* geff;           // 4*nx
  double* v_rest;         // device scalar
};

// assembly.cu
void launch_residual(const Dev& g, const double* u, const double* r0, double* r, int mode, cudaStream_t st);
void launch_jacobian_volume(const Dev& g, const double* u, double* A, cudaStream_t st);
void launch_channels(const Dev& g, const ChannelDev& ch, const double* u, int mode, double dt_ms, int automatic,
                     double v_fixed, int* err, cudaStream_t st);
void launch_channel_accept(int nx, const ChannelDev& ch, cudaStream_t st);
void launch_memb_pot(const Dev& g, const double* u, double* vm, cudaStream_t st);
void launch_memb_flux_terms(const Dev& g, const double* u, double* out, cudaStream_t st);
// boundary_fd.cu (compiled with --fmad=false)
void launch_jacobian_boundary_fd(const Dev& g, const double* u, double* A, cudaStream_t st);
void launch_jacobian_volume_fd(const Dev& g, const double* u, double* A, cudaStream_t st);  // adds the g.fd_mask parts

// solver.cu
struct Scalars {  // device-resident Krylov scalars
  double rho, rho_new, alpha, omega, beta, h, norm, norm0, tr, tt, it_half, reduction, rho_next, pad0;
  int done, converged, breakdown, x_pending;   // x_pending: converged at the half step, x += alpha y still to be applied
};
// Ac != null: multiply with the 64-byte-block copy of A (launch_compact_rows) instead of A itself
void launch_spmv(const Dev& g, const double* A, const double* x, double* y, int zero_constrained, cudaStream_t st,
                 const double* Ac = nullptr);
void launch_compact_rows(const Dev& g, const double* A, double* Ac, int* flag, cudaStream_t st);
void launch_spmv_residual(const Dev& g, const double* A, const double* x, const double* b, double* r,
                          int zero_constrained, cudaStream_t st);
int spmv_blocks(const Dev& g);
// SpMV with fused dot products: dot_mode 1: partial <w, y>; 2: partial <y, w> and <y, y>; one partial per block
void launch_spmv_dot(const Dev& g, const double* A, const double* x, double* y, int zero_constrained, int dot_mode,
                     const double* w, int own_lo, int own_hi, int masked, const Scalars* sc, double* partials,
                     cudaStream_t st, const double* Ac = nullptr);
// jc > 0 (everywhere below): radial split of the sub-slabs at grid row jc, see solver.cu: ypart_select
void ilu_level_tables(const Dev& g, int slab_w, std::vector<int>& tab, int jc = 0);
void launch_ilu_factor(const Dev& g, const double* A, double* LU, int slab_w, const int* lvl_tab, cudaStream_t st,
                       int jc = 0);
// hp != null: multi-GPU with peer windows, the sweep also pushes the halo columns to the neighbours (p2p.cuh)
void launch_ilu_apply(const Dev& g, const double* LU, const double* d, double* v, int slab_w, const int* lvl_tab,
                      cudaStream_t st, const HaloPush* hp = nullptr, int jc = 0);
void halo_contributors(const Dev& g, int slab_w, int* left, int* right, int jc = 0);
// ilu1.cu: block ILU(1), 13 blocks per row, wavefronts il + 3 j
int ilu1_max_slab_w(const Dev& g);
void ilu1_level_tables(const Dev& g, int slab_w, std::vector<int>& tab);
void launch_ilu1_factor(const Dev& g, const double* A, double* LU, int slab_w, const int* lvl_tab, cudaStream_t st);
void launch_ilu1_apply(const Dev& g, const double* LU, const double* d, double* v, int slab_w, const int* lvl_tab,
                       cudaStream_t st, const HaloPush* hp = nullptr);
// amg.cu: x-aggregation multigrid (piecewise-constant aggregates of cx vertex columns, Galerkin coarse
// operators in the same 9-slot arrow format) -- transfer kernels; the cycle itself is in driver.cu
void launch_galerkin_x(int nvx_f, int nvx_c, int nvy, int cx, const double* Af, double* Ac, const uint8_t* dmask_f,
                       cudaStream_t st);
void launch_restrict_x(int nvx_f, int nvx_c, int nvy, int cx, const double* rf, double* bc, const uint8_t* dmask_f,
                       cudaStream_t st);
void launch_prolong_add_x(int nvx_f, int nvx_c, int nvy, int cx, const double* xc, double* xf, const uint8_t* dmask_f,
                          cudaStream_t st);
void launch_vec_add(int n, double* x, const double* t, cudaStream_t st);   // x += t
void launch_halo_push(const Dev& g, const double* v, const HaloPush& hp, cudaStream_t st);
void launch_rownorm(const Dev& g, double* A, double* r, cudaStream_t st);
void launch_compact_bcrs(const Dev& g, const double* A9, double* blocks, const int* rowptr, cudaStream_t st);
void launch_expand_bcrs(const Dev& g, const double* blocks, double* A9, const int* rowptr, int* viol, cudaStream_t st);

}  // namespace ax1
