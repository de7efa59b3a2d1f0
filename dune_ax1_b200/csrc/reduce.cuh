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
template <int NS>
__device__ __forceinline__ void block_reduce_store(double (&s)[NS], double* partials) {
  __shared__ double sh[NS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
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
