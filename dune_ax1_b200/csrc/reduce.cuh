// reduce.cuh -- deterministic block reductions and the owner mask shared by solver.cu / driver.cu.
#pragma once
#include "kernels.h"

namespace ax1 {

// owner mask of the overlapping decomposition: only vertex columns [lo, hi] count in dot products
// (OVLPScalarProduct, ax1_solverbackend.hh:413-437)
struct Own { int lo, hi, nvx, masked; };

__device__ __forceinline__ bool owned(const Own& o, int v) {
  if (!o.masked) return true;
  const int i = v % o.nvx;
  return i >= o.lo && i <= o.hi;
}

// warp-shuffle + shared-memory block reduction of NS running sums; thread 0 of the block stores
// partials[k * gridDim.x + blockIdx.x]. Fixed order -> bitwise reproducible.
// Two sums (every user today): the first shuffle swaps halves -- lanes 0-15 go on with sum 0, lanes 16-31 with
// sum 1 -- so a warp spends 5 double shuffles instead of 10, and the second stage of a block of <= 16 warps is packed
// the same way (the epilogue of the dot-fused SpMV cost 0.13 ms per launch at full size with the two-pass form).
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&s)[NS], double* partials) {
  __shared__ double sh[NS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  if (NS == 2 && nwarps <= 16) {
    const bool hi = lane & 16;
    double m = (hi ? s[1] : s[0]) + __shfl_xor_sync(0xffffffffu, hi ? s[0] : s[1], 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if ((lane & 15) == 0) sh[hi][w] = m;
    __syncthreads();
    if (w == 0) {
      double t = (lane & 15) < nwarps ? sh[hi][lane & 15] : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if ((lane & 15) == 0) partials[(hi ? 1 : 0) * gridDim.x + blockIdx.x] = t;
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < NS; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) sh[k][w] = s[k];
  }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      double t = lane < nw ? sh[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0) partials[k * gridDim.x + blockIdx.x] = t;
    }
  }
}

}  // namespace ax1
