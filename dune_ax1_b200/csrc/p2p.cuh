// p2p.cuh -- the two exchange steps of the overlapping Krylov solve done with direct stores into the
// neighbour GPU's memory over NVLink (CUDA IPC windows, one process per GPU) instead of NCCL calls:
//   halo sum     OverlappingWrappedPreconditioner::apply -> gv.communicate(AddDataHandle, All_All_Interface)
//                (restated in dune/ax1/common/ax1_solverbackend.hh:346-381): the sweep kernel writes the
//                three shared vertex columns straight into the neighbour's window while it produces them;
//                the last contributing CTA releases a sequence flag; the receiver acquires it and adds.
//   dot products OVLPScalarProduct (ax1_solverbackend.hh:413-437): every rank stores its owner-masked
//                partial sums into every peer's window and sums the nranks slots in rank order, inside the
//                kernel that also runs the Krylov scalar recurrence -- one launch per reduction.
// Windows are double-buffered by the parity of the sequence number; a rank can be at most one exchange
// ahead of a neighbour because every exchange waits for the neighbour's contribution of the same number.
#pragma once
#include <cuda_runtime.h>

namespace ax1 {

#define AX1_MAX_RANKS 16
#define AX1_P2P_RED_OFF 0                      // [2 parity][AX1_MAX_RANKS][8 doubles]: s0 s1 . . flag . . .
#define AX1_P2P_FLAG_OFF (2 * AX1_MAX_RANKS * 8)  // [2 sides] u64 halo sequence flags (+6 padding)
#define AX1_P2P_HALO_OFF (AX1_P2P_FLAG_OFF + 8)   // [2 sides][2 parity][3 * nvy * 4 doubles]
// A lost peer must not hang the GPU for ever, but ordinary rank skew (a rank writing output, a lazy cudaMalloc) must
// never trip the limit: default 600 s (AX1_P2P_TIMEOUT_S), converted to SM cycles at comm_init. After a timeout the
// exchange sets Scalars::done with NaN sums, the host reports AX1GPU_ENCCL at its next poll and the handle's
// communication is marked broken (sequence counters of the ranks no longer agree) -- see driver.cu: check_p2p.
#define AX1_P2P_TIMEOUT_DEFAULT_S 600.0

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool wait_seq(const unsigned long long* flag, unsigned long long seq, long long timeout) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < seq)
    if (clock64() - t0 > timeout) return false;
  return true;
}

// what a sweep kernel needs to push its halo columns (passed by value)
struct HaloPush {
  double* dst[2];                 // neighbour window slot for my left / right 3 vertex columns (null: no neighbour)
  unsigned long long* flag[2];    // neighbour's sequence flag for that slot
  unsigned int* cnt;              // local arrival counters [2] (last-CTA detection)
  int ncontrib[2];                // sub-slab CTAs that own part of the left / right halo columns
  unsigned long long seq;
};

// called where the sweep stores a final solution value: column ig (local vertex column), row j, component r
__device__ __forceinline__ void halo_push_value(const HaloPush& hp, int nvx, int nvy, int ig, int j, int r, double x) {
  if (ig < 3) {
    if (hp.dst[0]) hp.dst[0][((size_t)ig * nvy + j) * 4 + r] = x;
  }
  if (ig >= nvx - 3) {
    if (hp.dst[1]) hp.dst[1][((size_t)(ig - (nvx - 3)) * nvy + j) * 4 + r] = x;
  }
}
// called by every thread of a sweep CTA after its last store; i0, wc = vertex-column range of the CTA
__device__ __forceinline__ void halo_push_finish(const HaloPush& hp, int nvx, int i0, int wc) {
  const bool touch_l = hp.dst[0] && i0 < 3, touch_r = hp.dst[1] && (i0 + wc > nvx - 3);
  if (!touch_l && !touch_r) return;
  __threadfence_system();  // my remote stores are ordered before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    if (touch_l && atomicAdd(&hp.cnt[0], 1u) + 1u == (unsigned)hp.ncontrib[0]) {
      hp.cnt[0] = 0;
      st_release_sys(hp.flag[0], hp.seq);
    }
    if (touch_r && atomicAdd(&hp.cnt[1], 1u) + 1u == (unsigned)hp.ncontrib[1]) {
      hp.cnt[1] = 0;
      st_release_sys(hp.flag[1], hp.seq);
    }
  }
}

// one-shot all-reduce of two doubles through the peer windows (see reduce_final_kernel in driver.cu)
struct RedP2P {
  double* peer[AX1_MAX_RANKS];  // window base of every rank (peer[rank] = own window)
  int rank, nranks;
  unsigned long long seq;
  long long timeout;            // SM cycles a rank waits for a peer before it gives up (ctx.p2p.timeout_cycles)
};

}  // namespace ax1
