// common.cuh -- shared device-side definitions of libax1gpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/ax1gpu.h"

namespace ax1 {

// dune/ax1/common/constants.hh:64-74
constexpr double con_pi = 3.1415926535897932384626433;
constexpr double con_e = 1.602176487e-19;
constexpr double con_k = 1.380650e-23;
constexpr double con_eps0 = 8.854187817e-12;
constexpr double con_mol = 6.02214129e23;

enum { NA = 0, KK = 1, CL = 2, NSPEC = 3, POT = 3 };
enum { SD_CYTOSOL = 0, SD_ES = 1, SD_MEMBRANE = 2 };

// 2-point Gauss-Legendre on [0,1] (quadrature order 2, operator_fully_coupled.hh:161)
#define AX1_GP0 0.2113248654051871177454256097490212721762
#define AX1_GP1 0.7886751345948128822545743902509787278238

// Everything a kernel needs to know about the problem; passed by value.
struct Dev {
  int nx, ny, nvx, nvy, nv, jm;
  const double* x;
  const double* y;
  const double* node_vol;     // nv
  const uint8_t* dmask;       // nv
  const double* eps_memb;     // nx
  const double* geff;         // 4*nx: Na_leak, K_leak, Nav, Kv new effective conductances
  double eps_elec, PC, D[3], S_pot, kT_e, temperature;
  double time_scale, length_scale;
  int valence[3];
  int do_volume_scaling, membrane_active;
  // stimulation
  int stim_on;                // evaluated on host for the current LOP time
  double stim_x, stim_y, I_inj;
  // equilibration
  int equil;                  // do_equilibration && time < t_equilibrium (host evaluated)
  double dummy_value;
  double dt;
  // volume Jacobian by forward differences instead of the analytic blocks: bit 0 electrolyte operator,
  // bit 1 membrane operator, bit 2 time operator (the use*OperatorJacobianVolume = no default of the reference)
  int fd_mask;
  // trans-membrane flux stimulation (host evaluated per LOP time): species, and memb_flux_inj times the cosine ramp
  int fstim_on, fstim_ion;
  double fstim_amp, fstim_inj;
};

// Matrix layout: 9 stencil slots per vertex row, slot s = (dj+1)*3 + (di+1). Within a 4x4 block the
// PNP Jacobian is an "arrow": concentration row i couples only to its own species and to the
// potential, the potential row couples to everything (acme2_cyl_operator_fully_coupled.hh:516-561;
// the FD boundary blocks and the row-norm scaling keep that shape). A block is therefore stored as
// 10 doubles (80 B instead of 128 B):
//     [ d0 c0 | d1 c1 | d2 c2 | p0 p1 p2 p3 ]     d_i = (con_i,con_i), c_i = (con_i,pot), p_k = (pot,k)
// A[(v*9 + s)*10 + ...]; lane i<3 of a quad reads its 16 B pair, lane 3 its 32 B row.
// Missing neighbours (domain boundary) hold zero blocks. The ILU factor blocks are dense (solver.cu).
#define AX1_BLK 10
#define AX1_ROW 90
// the solver's SpMV operand: 8-double blocks [d0 c0 | d1 c1 | d2 c2 | q p3], p_j = z_j q, stored as
// [group of 8 vertices][slot][vertex in group][8] (solver.cu: compact_rows_kernel, spmv_compact_kernel)
#define AX1_CBLK 8
#define AX1_CROW 72
__host__ __device__ inline int slot_di(int s) { return s % 3 - 1; }
__host__ __device__ inline int slot_dj(int s) { return s / 3 - 1; }

#define AX1_CUDA(call)                                                                         \
  do {                                                                                         \
    cudaError_t e_ = (call);                                                                   \
    if (e_ != cudaSuccess) {                                                                   \
      ax1::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
      return AX1GPU_ECUDA;                                                                     \
    }                                                                                          \
  } while (0)

void set_error(const std::string& s);

}  // namespace ax1
