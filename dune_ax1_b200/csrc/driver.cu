// driver.cu -- host-side orchestration of the hot path: gating update, residual, Jacobian,
// BiCGStab(block ILU0) and the damped Newton loop, all on device-resident data.
//
//   BiCGStab   Dune::BiCGSTABSolver (dune-istl 2.3 solvers.hh) as driven by
//              ISTLBackend_{SEQ,OVLP}_BCGS_ILUn (acme2_cyl_setup.hh:764-765,797-798)     [ext]
//   overlap    OverlappingOperator / OverlappingWrappedPreconditioner / OVLPScalarProduct,
//              restated in dune/ax1/common/ax1_solverbackend.hh:306-443
//   Newton     PDELab NewtonSolver::apply [ext] + Ax1Newton::{defect,prepare_step,line_search}
//              dune/ax1/common/ax1_newton.hh:194-205, 216-293, 308-503
// Krylov scalars (rho, alpha, omega, norms, convergence flag) live on the device; the host only
// polls the flag, so a solve is a stream of kernel launches without per-iteration round trips.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "ctx.h"
#include "reduce.cuh"

namespace ax1 {

#define AX1_LAUNCH_CHECK(c)                                  \
  do {                                                       \
    (c)->launches++;                                         \
    cudaError_t e_ = cudaGetLastError();                     \
    if (e_ != cudaSuccess) {                                 \
      set_error(std::string("kernel launch: ") + cudaGetErrorString(e_)); \
      return AX1GPU_ECUDA;                                   \
    }                                                        \
  } while (0)

// ------------------------------------------------------------------------------------------
// vector kernels: one thread per vertex block (4 doubles), grid-stride
__device__ __forceinline__ void ld4v(const double* p, double* o) {
  const double2 lo = *reinterpret_cast<const double2*>(p);
  const double2 hi = *reinterpret_cast<const double2*>(p + 2);
  o[0] = lo.x; o[1] = lo.y; o[2] = hi.x; o[3] = hi.y;
}
__device__ __forceinline__ void st4v(double* p, const double* v) {
  reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
  reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

// sums[0] = <a,b>, sums[1] = <c,d> (if two)
__global__ void __launch_bounds__(256) dot_kernel(int nv, const double* __restrict__ a, const double* __restrict__ b,
                                                  const double* __restrict__ c, const double* __restrict__ d, int two,
                                                  Own own, const Scalars* sc, double* partials) {
  double s[2] = {0.0, 0.0};
  if (!(sc && sc->done)) {
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
      if (!owned(own, v)) continue;
      double x[4], y[4];
      ld4v(a + 4 * (size_t)v, x); ld4v(b + 4 * (size_t)v, y);
      s[0] += x[0] * y[0]; s[0] += x[1] * y[1]; s[0] += x[2] * y[2]; s[0] += x[3] * y[3];
      if (two) {
        ld4v(c + 4 * (size_t)v, x); ld4v(d + 4 * (size_t)v, y);
        s[1] += x[0] * y[0]; s[1] += x[1] * y[1]; s[1] += x[2] * y[2]; s[1] += x[3] * y[3];
      }
    }
  }
  block_reduce_store<2>(s, partials);
}

// p = r (first) or p = (p - omega v) beta + r
__global__ void __launch_bounds__(256) pupdate_kernel(int nv, double* p, const double* __restrict__ v,
                                                      const double* __restrict__ r, const Scalars* sc, int first) {
  if (sc->done) return;
  const double omega = sc->omega, beta = sc->beta;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
    double pr[4], rr[4];
    ld4v(r + 4 * (size_t)k, rr);
    if (first) {
      st4v(p + 4 * (size_t)k, rr);
    } else {
      double vv[4];
      ld4v(p + 4 * (size_t)k, pr); ld4v(v + 4 * (size_t)k, vv);
#pragma unroll
      for (int c = 0; c < 4; ++c) { double t = pr[c] + (-omega) * vv[c]; t *= beta; t += rr[c]; pr[c] = t; }
      st4v(p + 4 * (size_t)k, pr);
    }
  }
}

// The two vector updates of a BiCGStab iteration. x is read by nobody inside the iteration, so its first update
// (x += alpha y) is applied together with the second one -- the same two fused multiply-adds in the same order, i.e.
// bit-identical x, for one read and one write of x less per iteration (64 of 416 B per vertex):
//   first half:   r -= alpha v ;                         partial <r,r>
//   second half:  x = (x + alpha y) + omega z ; r -= omega t ; partial <r,r>, <rt,r> (the next rho)
// A solve that converges at the half step leaves x += alpha y pending (Scalars::x_pending, x_fixup_kernel).
__global__ void __launch_bounds__(256) r_update_kernel(int nv, double* r, const double* __restrict__ v, const Scalars* sc,
                                                       Own own, double* partials) {
  double s[2] = {0.0, 0.0};
  if (!sc->done) {
    const double a = sc->alpha;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
      double rv[4], wv[4];
      ld4v(r + 4 * (size_t)k, rv); ld4v(v + 4 * (size_t)k, wv);
#pragma unroll
      for (int c = 0; c < 4; ++c) rv[c] += (-a) * wv[c];
      st4v(r + 4 * (size_t)k, rv);
      if (owned(own, k)) { s[0] += rv[0] * rv[0]; s[0] += rv[1] * rv[1]; s[0] += rv[2] * rv[2]; s[0] += rv[3] * rv[3]; }
    }
  }
  block_reduce_store<2>(s, partials);
}
__global__ void __launch_bounds__(256) xr_update_kernel(int nv, double* x, const double* __restrict__ y, const double* __restrict__ z,
                                                        double* r, const double* __restrict__ t, const double* __restrict__ rt,
                                                        const Scalars* sc, Own own, double* partials) {
  double s[2] = {0.0, 0.0};
  if (!sc->done) {
    const double a = sc->alpha, om = sc->omega;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
      double xv[4], yv[4], zv[4], rv[4], wv[4], tv[4];
      ld4v(x + 4 * (size_t)k, xv); ld4v(y + 4 * (size_t)k, yv); ld4v(z + 4 * (size_t)k, zv);
      ld4v(r + 4 * (size_t)k, rv); ld4v(t + 4 * (size_t)k, wv);
#pragma unroll
      for (int c = 0; c < 4; ++c) { xv[c] += a * yv[c]; xv[c] += om * zv[c]; rv[c] += (-om) * wv[c]; }
      st4v(x + 4 * (size_t)k, xv); st4v(r + 4 * (size_t)k, rv);
      if (owned(own, k)) {
        s[0] += rv[0] * rv[0]; s[0] += rv[1] * rv[1]; s[0] += rv[2] * rv[2]; s[0] += rv[3] * rv[3];
        ld4v(rt + 4 * (size_t)k, tv);
        s[1] += tv[0] * rv[0]; s[1] += tv[1] * rv[1]; s[1] += tv[2] * rv[2]; s[1] += tv[3] * rv[3];
      }
    }
  }
  block_reduce_store<2>(s, partials);
}
// x += alpha y after a solve that converged at the half step
__global__ void __launch_bounds__(256) x_fixup_kernel(int nv, double* x, const double* __restrict__ y, const Scalars* sc) {
  const double a = sc->alpha;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
    double xv[4], yv[4];
    ld4v(x + 4 * (size_t)k, xv); ld4v(y + 4 * (size_t)k, yv);
#pragma unroll
    for (int c = 0; c < 4; ++c) xv[c] += a * yv[c];
    st4v(x + 4 * (size_t)k, xv);
  }
}

// u = uprev - lambda z
__global__ void __launch_bounds__(256) newton_update_kernel(int nv, double* u, const double* __restrict__ uprev,
                                                            const double* __restrict__ z, double lambda) {
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
    double a[4], b[4];
    ld4v(uprev + 4 * (size_t)k, a); ld4v(z + 4 * (size_t)k, b);
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] += (-lambda) * b[c];
    st4v(u + 4 * (size_t)k, a);
  }
}

// ---- restarted GMRES vector kernels (modified Gram-Schmidt) ----
// w -= gm[GM_SCALE] * vprev (if vprev), then the partial of <w, vk> (vk == nullptr: <w, w>), owner-masked
__global__ void __launch_bounds__(256) gmres_mgs_kernel(int nv, double* w, const double* __restrict__ vprev,
                                                        const double* __restrict__ vk, const double* gm,
                                                        const Scalars* sc, Own own, double* partials) {
  double s[2] = {0.0, 0.0};
  if (!sc->done) {
    const double h = vprev ? gm[1] : 0.0;   // GM_SCALE
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
      double wv[4];
      ld4v(w + 4 * (size_t)k, wv);
      if (vprev) {
        double pv[4];
        ld4v(vprev + 4 * (size_t)k, pv);
#pragma unroll
        for (int c = 0; c < 4; ++c) wv[c] += (-h) * pv[c];
        st4v(w + 4 * (size_t)k, wv);
      }
      if (owned(own, k)) {
        if (vk) {
          double kv[4];
          ld4v(vk + 4 * (size_t)k, kv);
          s[0] += wv[0] * kv[0]; s[0] += wv[1] * kv[1]; s[0] += wv[2] * kv[2]; s[0] += wv[3] * kv[3];
        } else {
          s[0] += wv[0] * wv[0]; s[0] += wv[1] * wv[1]; s[0] += wv[2] * wv[2]; s[0] += wv[3] * wv[3];
        }
      }
    }
  }
  block_reduce_store<2>(s, partials);
}
// dst = src * gm[GM_SCALE]
__global__ void __launch_bounds__(256) gmres_scale_kernel(int nv, double* dst, const double* src, const double* gm,
                                                          const Scalars* sc) {
  if (sc->done) return;
  const double a = gm[1];
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
    double v[4];
    ld4v(src + 4 * (size_t)k, v);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] *= a;
    st4v(dst + 4 * (size_t)k, v);
  }
}
// RestartedGMResSolver::update: back substitution y = H^-1 s over the cnt columns of this cycle
__global__ void gmres_ysolve_kernel(double* gm) {
  const int m = (int)gm[3], cnt = (int)gm[2];
  const double* H = gm + 8;
  const double* sv = H + (size_t)(m + 1) * m + 2 * (size_t)m;
  double* y = gm + 8 + (size_t)(m + 1) * m + 3 * (size_t)m + 1;
  for (int a = 0; a < cnt; ++a) y[a] = sv[a];
  for (int a = cnt - 1; a >= 0; --a) {
    y[a] /= H[(size_t)a * m + a];
    for (int c = a - 1; c >= 0; --c) y[c] -= H[(size_t)c * m + a] * y[a];
  }
}
// w = sum_{c < cnt} y[c] V_c ; x += w
__global__ void __launch_bounds__(256) gmres_combine_kernel(int nv, const double* __restrict__ V, size_t ldv,
                                                            const double* gm, double* w, double* x) {
  const int m = (int)gm[3], cnt = (int)gm[2];
  const double* y = gm + 8 + (size_t)(m + 1) * m + 3 * (size_t)m + 1;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nv; k += gridDim.x * blockDim.x) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < cnt; ++c) {
      double v[4];
      ld4v(V + (size_t)c * ldv + 4 * (size_t)k, v);
      const double yc = y[c];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += yc * v[q];
    }
    double xv[4];
    ld4v(x + 4 * (size_t)k, xv);
#pragma unroll
    for (int q = 0; q < 4; ++q) xv[q] += acc[q];
    st4v(w + 4 * (size_t)k, acc);
    st4v(x + 4 * (size_t)k, xv);
  }
}

enum { OP_NONE = -1, OP_NORM0 = 0, OP_H, OP_NORM_A, OP_OMEGA, OP_NORM_B, OP_PLAIN_NORM, OP_G_BETA, OP_G_H, OP_G_NORMW };

// device-resident state of the restarted GMRES (one flat double array `gm`): scalars, then the Hessenberg
// matrix H[(m+1) x m] (row-major), the Givens rotations cs[m], sn[m], the rotated rhs s[m+1] and y[m]
enum { GM_BETA = 0, GM_SCALE = 1, GM_CNT = 2, GM_M = 3, GM_H = 8 };
__host__ __device__ inline size_t gm_len(int m) { return GM_H + (size_t)(m + 1) * m + 4 * (size_t)m + 8; }

// Dune::RestartedGMResSolver::generatePlaneRotation / applyPlaneRotation (dune-istl 2.3 solvers.hh) [ext]
__device__ inline void gm_generate_rotation(double dx, double dy, double& cs, double& sn) {
  if (dy == 0.0) { cs = 1.0; sn = 0.0; }
  else if (fabs(dy) > fabs(dx)) { const double t = dx / dy; sn = 1.0 / sqrt(1.0 + t * t); cs = t * sn; }
  else { const double t = dy / dx; cs = 1.0 / sqrt(1.0 + t * t); sn = t * cs; }
}
__device__ inline void gm_apply_rotation(double& dx, double& dy, double cs, double sn) {
  const double t = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = t;
}

// the Krylov scalar recurrences of Dune::BiCGSTABSolver / Dune::RestartedGMResSolver, executed by one
// thread on the device. GMRES ops: out = the gm array, aux = step indices.
__device__ void scalar_step(int op, Scalars* sc, const double* sums, double reduction, double* out, int aux) {
  const double EPSILON = 1e-80;
  if (op == OP_PLAIN_NORM) { out[0] = sqrt(sums[0]); return; }
  if (sc->done) return;
  switch (op) {
    case OP_G_BETA: {   // beta = |M^-1 (b - A x)| at the start (aux = 1) / after a restart (aux = 0)
      double* gm = out;
      const int m = (int)gm[GM_M];
      double* sv = gm + GM_H + (size_t)(m + 1) * m + 2 * (size_t)m;
      const double beta = sqrt(sums[0]);
      gm[GM_BETA] = beta;
      if (aux) { sc->norm0 = beta == 0.0 ? 1.0 : beta; sc->it_half = 0.0; sc->converged = 0; sc->breakdown = 0; }
      sc->norm = beta;
      if (aux ? beta <= reduction * sc->norm0 : beta < reduction * sc->norm0) { sc->converged = 1; sc->done = 1; break; }
      gm[GM_SCALE] = 1.0 / beta;
      gm[GM_CNT] = 0.0;
      sv[0] = beta;
      for (int i = 1; i <= m; ++i) sv[i] = 0.0;
      break;
    }
    case OP_G_H: {      // H[k][i] = <w, v_k>, aux = k * 1024 + i
      double* gm = out;
      const int m = (int)gm[GM_M], k = aux >> 10, i = aux & 1023;
      gm[GM_H + (size_t)k * m + i] = sums[0];
      gm[GM_SCALE] = sums[0];
      break;
    }
    case OP_G_NORMW: {  // H[i+1][i] = |w|, Givens update of column i, residual estimate |s[i+1]|
      double* gm = out;
      const int m = (int)gm[GM_M], i = aux;
      double* H = gm + GM_H;
      double* cs = H + (size_t)(m + 1) * m;
      double* sn = cs + m;
      double* sv = sn + m;
      const double hn = sqrt(sums[0]);
      H[(size_t)(i + 1) * m + i] = hn;
      if (hn == 0.0) { sc->breakdown = 1; sc->done = 1; break; }
      gm[GM_SCALE] = 1.0 / hn;
      for (int k = 0; k < i; ++k) gm_apply_rotation(H[(size_t)k * m + i], H[(size_t)(k + 1) * m + i], cs[k], sn[k]);
      gm_generate_rotation(H[(size_t)i * m + i], H[(size_t)(i + 1) * m + i], cs[i], sn[i]);
      gm_apply_rotation(H[(size_t)i * m + i], H[(size_t)(i + 1) * m + i], cs[i], sn[i]);
      gm_apply_rotation(sv[i], sv[i + 1], cs[i], sn[i]);
      sc->norm = fabs(sv[i + 1]);
      sc->it_half += 1.0;
      gm[GM_CNT] = (double)(i + 1);
      if (sc->norm < reduction * sc->norm0) { sc->converged = 1; sc->done = 1; }
      break;
    }
    case OP_NORM0:
      sc->norm = sc->norm0 = sqrt(sums[0]);
      sc->rho = sc->alpha = sc->omega = 1.0;
      sc->rho_new = sums[0];  // rt = r  ->  <rt, r> = <r, r>
      sc->it_half = 0.5; sc->converged = 0; sc->breakdown = 0; sc->x_pending = 0;
      if (sc->norm < reduction * sc->norm0 || sc->norm < 1e-30) { sc->converged = 1; sc->done = 1; sc->it_half = 0.0; }
      break;
    case OP_H:
      sc->h = sums[0];
      if (fabs(sc->h) < EPSILON) { sc->breakdown = 1; sc->done = 1; break; }
      sc->alpha = sc->rho_new / sc->h;
      break;
    case OP_NORM_A:
      sc->norm = sqrt(sums[0]);
      if (sc->norm < reduction * sc->norm0) { sc->converged = 1; sc->done = 1; sc->x_pending = 1; break; }
      sc->it_half += 0.5;
      break;
    case OP_OMEGA:
      sc->tr = sums[0]; sc->tt = sums[1];
      sc->omega = sc->tr / sc->tt;
      break;
    case OP_NORM_B:
      sc->rho = sc->rho_new;
      sc->norm = sqrt(sums[0]);
      if (sc->norm < reduction * sc->norm0 || sc->norm < 1e-30) { sc->converged = 1; sc->done = 1; break; }
      sc->it_half += 0.5;
      // top of the next iteration: rho_new = <rt, r> (fused into the x,r update), breakdown tests, beta
      sc->rho_new = sums[1];
      if (fabs(sc->rho) <= EPSILON || fabs(sc->omega) <= EPSILON) { sc->breakdown = 1; sc->done = 1; break; }
      sc->beta = (sc->rho_new / sc->rho) * (sc->alpha / sc->omega);
      break;
  }
}

// partials -> sums (fixed order) and, on one GPU, the scalar recurrence in the same launch
__global__ void __launch_bounds__(1024) reduce_final_kernel(const double* partials, int nblocks, double* sums, int op,
                                                            Scalars* sc, double reduction, double* out, int aux) {
  __shared__ double sh[2][32];
  double s[2] = {0.0, 0.0};
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { s[0] += partials[k]; s[1] += partials[nblocks + k]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    if (lane == 0) sh[q][w] = s[q];
  }
  __syncthreads();
  if (w == 0) {
    double t0 = sh[0][lane], t1 = sh[1][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
    if (lane == 0) {
      sums[0] = t0; sums[1] = t1;
      if (op != OP_NONE) scalar_step(op, sc, sums, reduction, out, aux);
    }
  }
}

// multi-GPU with peer windows: block reduction, one-shot all-reduce over NVLink and the scalar recurrence
// in ONE launch. Every rank stores its two partial sums into slot [parity][rank] of every peer's window,
// releases the slot's sequence flag, then acquires the nranks slots of its own window and adds them in
// rank order -> all ranks obtain bitwise identical sums (they must take identical convergence decisions).
__global__ void __launch_bounds__(1024) reduce_final_p2p_kernel(const double* partials, int nblocks, double* sums, int op,
                                                                Scalars* sc, double reduction, double* out, RedP2P rp,
                                                                int* err, int aux) {
  __shared__ double sh[2][32];
  double s[2] = {0.0, 0.0};
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { s[0] += partials[k]; s[1] += partials[nblocks + k]; }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    if (lane == 0) sh[q][w] = s[q];
  }
  __syncthreads();
  if (w == 0) {
    double t0 = sh[0][lane], t1 = sh[1][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
    const size_t slot = AX1_P2P_RED_OFF + (size_t)(rp.seq & 1) * AX1_MAX_RANKS * 8;
    double a0 = 0.0, a1 = 0.0;
    bool ok = true;
    if (lane < rp.nranks) {
      volatile double* p = rp.peer[lane] + slot + (size_t)rp.rank * 8;  // my slot in rank `lane`'s window
      p[0] = t0; p[1] = t1;
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(rp.peer[lane] + slot + (size_t)rp.rank * 8 + 4), rp.seq);
      const double* q = rp.peer[rp.rank] + slot + (size_t)lane * 8;       // rank `lane`'s slot in my window
      ok = wait_seq(reinterpret_cast<const unsigned long long*>(q + 4), rp.seq, rp.timeout);
      a0 = reinterpret_cast<const volatile double*>(q)[0];
      a1 = reinterpret_cast<const volatile double*>(q)[1];
    }
    const bool all_ok = __all_sync(0xffffffffu, ok);
    double r0 = 0.0, r1 = 0.0;
    for (int k = 0; k < rp.nranks; ++k) { r0 += __shfl_sync(0xffffffffu, a0, k); r1 += __shfl_sync(0xffffffffu, a1, k); }
    if (lane == 0) {
      if (!all_ok) {
        // a peer did not arrive: its slot holds stale data. Nobody may take a convergence decision from it --
        // NaN sums, stop the solve, flag the error for the host's next poll (check_p2p)
        atomicExch(err, 1);
        const double nan = __longlong_as_double(0x7ff8000000000000LL);
        sums[0] = nan; sums[1] = nan;
        sc->norm = nan; sc->converged = 0; sc->done = 1;
      } else {
        sums[0] = r0; sums[1] = r1;
        if (op != OP_NONE) scalar_step(op, sc, sums, reduction, out, aux);
      }
    }
  }
}

// receiver side of the halo sum with peer windows: acquire the neighbour's sequence flag, then add the
// three vertex columns it stored into my window (AddDataHandle)
__global__ void halo_wait_add_kernel(Dev g, double* v, const double* slot, const unsigned long long* flag,
                                     unsigned long long seq, int c0, int* err, long long timeout, Scalars* sc) {
  __shared__ int ok_sh;
  if (threadIdx.x == 0) ok_sh = wait_seq(flag, seq, timeout) ? 1 : 0;
  __syncthreads();
  if (!ok_sh) {   // the neighbour's columns never came: do not add stale data, stop the running solve
    if (threadIdx.x == 0) { atomicExch(err, 1); sc->converged = 0; sc->done = 1; }
    return;
  }
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 3 * g.nvy * 4;
  if (t >= n) return;
  const int k = t & 3, j = (t >> 2) % g.nvy, c = (t >> 2) / g.nvy;
  v[4 * (size_t)(c0 + c + g.nvx * j) + k] += reinterpret_cast<const volatile double*>(slot)[t];
}

__global__ void scalar_kernel(int op, Scalars* sc, const double* sums, double reduction, double* out, int aux) {
  scalar_step(op, sc, sums, reduction, out, aux);
}

// halo pack / unpack-add of 3 vertex columns per processor-boundary side
__global__ void halo_pack_kernel(Dev g, const double* v, double* buf, int c0, int side_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 3 * g.nvy * 4;
  if (t >= n) return;
  const int k = t & 3, j = (t >> 2) % g.nvy, c = (t >> 2) / g.nvy;
  buf[side_off + t] = v[4 * (size_t)(c0 + c + g.nvx * j) + k];
}
__global__ void halo_add_kernel(Dev g, double* v, const double* buf, int c0, int side_off) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = 3 * g.nvy * 4;
  if (t >= n) return;
  const int k = t & 3, j = (t >> 2) % g.nvy, c = (t >> 2) / g.nvy;
  v[4 * (size_t)(c0 + c + g.nvx * j) + k] += buf[side_off + t];
}

// ------------------------------------------------------------------------------------------
static inline int vec_grid(const ax1gpu_ctx* c) { return c->nred_blocks; }
static inline Own own_of(const ax1gpu_ctx* c) { return Own{c->own_lo, c->own_hi, c->g.nvx, c->nranks > 1}; }

#define AX1_NCCL(c, call)                                                             \
  do {                                                                                \
    ncclResult_t r_ = (call);                                                         \
    if (r_ != ncclSuccess) {                                                          \
      set_error(std::string(#call) + ": " + (c)->nccl.GetErrorString(r_));            \
      return AX1GPU_ENCCL;                                                            \
    }                                                                                 \
  } while (0)

// first stage of the reduction of many block partials (the dot products fused into the SpMV leave one partial per
// 64 block rows: 263 k at the named workload, which a single block took 135 us to add up): 512 blocks add a fixed
// contiguous chunk each -- the summation order stays fixed, so results stay bitwise reproducible
#define AX1_RED2 512
__global__ void __launch_bounds__(256) reduce_stage_kernel(const double* __restrict__ partials, int nblocks, double* out) {
  const int chunk = (nblocks + AX1_RED2 - 1) / AX1_RED2;
  const int lo = blockIdx.x * chunk, hi = min(nblocks, lo + chunk);
  double s[2] = {0.0, 0.0};
  for (int k = lo + threadIdx.x; k < hi; k += blockDim.x) { s[0] += partials[k]; s[1] += partials[nblocks + k]; }
  block_reduce_store<2>(s, out);
}

// partials -> sums (+ allreduce over ranks) -> scalar recurrence `op`
static int reduce_op_from(ax1gpu_ctx* c, const double* partials, int nblocks, int nsums, int op, double reduction,
                          double* out = nullptr, int aux = 0) {
  if (nblocks > 8 * AX1_RED2 && partials == c->d_spartials) {
    double* p2 = c->d_spartials + 2 * (size_t)spmv_blocks(c->g);
    reduce_stage_kernel<<<AX1_RED2, 256, 0, c->stream>>>(partials, nblocks, p2);
    AX1_LAUNCH_CHECK(c);
    partials = p2;
    nblocks = AX1_RED2;
  }
  if (c->nranks > 1 && c->p2p.on) {
    if (c->p2p.broken) return drv_check_p2p(c);
    RedP2P rp;
    for (int k = 0; k < AX1_MAX_RANKS; ++k) rp.peer[k] = c->p2p.peer[k];
    rp.rank = c->rank; rp.nranks = c->nranks; rp.seq = ++c->p2p.red_seq;
    rp.timeout = c->p2p.timeout_cycles;
    reduce_final_p2p_kernel<<<1, 1024, 0, c->stream>>>(partials, nblocks, c->d_sums, op, c->d_sc, reduction, out, rp,
                                                       c->d_err + 2, aux);
    AX1_LAUNCH_CHECK(c);
  } else if (c->nranks > 1) {
    reduce_final_kernel<<<1, 1024, 0, c->stream>>>(partials, nblocks, c->d_sums, OP_NONE, c->d_sc, reduction, out, aux);
    AX1_LAUNCH_CHECK(c);
    AX1_NCCL(c, c->nccl.AllReduce(c->d_sums, c->d_sums, nsums, ncclDouble, ncclSum, c->comm, c->stream));
    scalar_kernel<<<1, 1, 0, c->stream>>>(op, c->d_sc, c->d_sums, reduction, out, aux);
    AX1_LAUNCH_CHECK(c);
  } else {
    reduce_final_kernel<<<1, 1024, 0, c->stream>>>(partials, nblocks, c->d_sums, op, c->d_sc, reduction, out, aux);
    AX1_LAUNCH_CHECK(c);
  }
  return 0;
}
static int reduce_op(ax1gpu_ctx* c, int nblocks, int nsums, int op, double reduction, double* out = nullptr,
                     int aux = 0) {
  return reduce_op_from(c, c->d_partials, nblocks, nsums, op, reduction, out, aux);
}

static int halo_add(ax1gpu_ctx* c, double* v) {
  if (c->nranks <= 1) return 0;
  const Dev& g = c->g;
  const int n = 3 * g.nvy * 4, bs = 256, nb = (n + bs - 1) / bs;
  if (c->left_pb) { halo_pack_kernel<<<nb, bs, 0, c->stream>>>(g, v, c->d_halo_send, 0, 0); AX1_LAUNCH_CHECK(c); }
  if (c->right_pb) { halo_pack_kernel<<<nb, bs, 0, c->stream>>>(g, v, c->d_halo_send, g.nvx - 3, n); AX1_LAUNCH_CHECK(c); }
  AX1_NCCL(c, c->nccl.GroupStart());
  if (c->left_pb) {
    AX1_NCCL(c, c->nccl.Send(c->d_halo_send, n, ncclDouble, c->rank - 1, c->comm, c->stream));
    AX1_NCCL(c, c->nccl.Recv(c->d_halo_recv, n, ncclDouble, c->rank - 1, c->comm, c->stream));
  }
  if (c->right_pb) {
    AX1_NCCL(c, c->nccl.Send(c->d_halo_send + n, n, ncclDouble, c->rank + 1, c->comm, c->stream));
    AX1_NCCL(c, c->nccl.Recv(c->d_halo_recv + n, n, ncclDouble, c->rank + 1, c->comm, c->stream));
  }
  AX1_NCCL(c, c->nccl.GroupEnd());
  if (c->left_pb) { halo_add_kernel<<<nb, bs, 0, c->stream>>>(g, v, c->d_halo_recv, 0, 0); AX1_LAUNCH_CHECK(c); }
  if (c->right_pb) { halo_add_kernel<<<nb, bs, 0, c->stream>>>(g, v, c->d_halo_recv, g.nvx - 3, n); AX1_LAUNCH_CHECK(c); }
  return 0;
}

// ------------------------------------------------------------------------------------------
int drv_refresh_dev(ax1gpu_ctx* c) {
  Dev& g = c->g;
  const ax1gpu_params& p = c->prm;
  g.dt = c->dt;
  const double t_lop = c->time + c->dt;  // OneStepMethod sets the LOP time to t+dt
  g.stim_on = p.do_stimulation && (t_lop > p.tinj_start && t_lop < p.tinj_end);
  g.equil = p.do_equilibration && c->time < p.t_equilibrium;
  g.fstim_on = p.do_memb_flux_stimulation && (t_lop > p.tinj_start && t_lop < p.tinj_end);
  g.fstim_ion = p.memb_flux_ion;
  g.fstim_inj = p.memb_flux_inj;
  {
    const double t_rel = (t_lop - p.tinj_start) / p.tinj_end;
    g.fstim_amp = g.fstim_on ? 0.5 * (1 - std::cos(2 * con_pi * t_rel)) : 0.0;   // times memb_flux_inj / |face| on the device
  }
  g.fd_mask = (p.use_elec_jacobian_volume ? 0 : 1) | (p.use_memb_jacobian_volume ? 0 : 2) |
              (p.use_time_jacobian_volume ? 0 : 4);
  return 0;
}

int drv_set_uold(ax1gpu_ctx* c) {
  drv_refresh_dev(c);
  launch_residual(c->g, c->d_uold, nullptr, c->d_r0, 1, c->stream);
  AX1_LAUNCH_CHECK(c);
  return 0;
}

int drv_update_state(ax1gpu_ctx* c, const double* d_u) {  // membranefluxgridfunction_implicit.hh:429-501
  const ax1gpu_params& p = c->prm;
  if (!p.membrane_active) return 0;
  if (!c->channels_initialized && c->channels_partially_initialized) return 0;
  drv_refresh_dev(c);
  if (!c->channels_initialized) {
    if (!p.do_equilibration || (std::fabs(c->time - p.t_equilibrium) < 1e-6 || c->time > p.t_equilibrium)) {
      launch_channels(c->g, c->ch, d_u, 0, 0.0, p.automatic_channel_init, p.channel_resting_potential, c->d_err,
                      c->stream);
      AX1_LAUNCH_CHECK(c);
      c->channels_partially_initialized = true;
    }
  } else {
    const double real_dt = c->dt * p.time_scale;
    const double dt_ms = real_dt / 1e-3;  // DynamicChannelSet::TIME_SCALE
    launch_channels(c->g, c->ch, d_u, 1, dt_ms, 0, 0.0, c->d_err, c->stream);
    AX1_LAUNCH_CHECK(c);
  }
  return 0;
}

int drv_accept_state(ax1gpu_ctx* c) {  // :524-546
  if (!c->prm.membrane_active) return 0;
  if (c->channels_partially_initialized) {
    c->channels_initialized = true;
    c->channels_partially_initialized = false;
  }
  launch_channel_accept(c->g.nx, c->ch, c->stream);
  AX1_LAUNCH_CHECK(c);
  return 0;
}

int drv_norm(ax1gpu_ctx* c, const double* a, double* out_host) {
  dot_kernel<<<vec_grid(c), 256, 0, c->stream>>>(c->g.nv, a, a, nullptr, nullptr, 0, own_of(c), nullptr, c->d_partials);
  AX1_LAUNCH_CHECK(c);
  int rc = reduce_op(c, c->nred_blocks, 1, OP_PLAIN_NORM, 0.0, c->d_sums + 2);
  if (rc) return rc;
  AX1_CUDA(cudaMemcpyAsync(c->h_pin, c->d_sums + 2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (c->nranks > 1 && c->p2p.on)
    AX1_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  *out_host = c->h_pin[0];
  return drv_check_p2p(c);
}

int drv_residual(ax1gpu_ctx* c, const double* d_u, double* d_r, double* norm_host) {
  drv_refresh_dev(c);
  launch_residual(c->g, d_u, c->d_r0, d_r, 0, c->stream);
  AX1_LAUNCH_CHECK(c);
  if (norm_host) return drv_norm(c, d_r, norm_host);
  return 0;
}

int drv_jacobian(ax1gpu_ctx* c, const double* d_u) {
  drv_refresh_dev(c);
  launch_jacobian_volume(c->g, d_u, c->d_A, c->stream);
  AX1_LAUNCH_CHECK(c);
  if (c->g.fd_mask) {   // NumericalJacobianVolume for the operators configured without jacobian_volume
    launch_jacobian_volume_fd(c->g, d_u, c->d_A, c->stream);
    AX1_LAUNCH_CHECK(c);
  }
  if (c->prm.membrane_active) {
    launch_jacobian_boundary_fd(c->g, d_u, c->d_A, c->stream);
    AX1_LAUNCH_CHECK(c);
  }
  c->factored = false;
  return 0;
}

static int ensure_level_tables(ax1gpu_ctx* c) {
  if (c->d_lvl && c->lvl_slab_w == c->slab_w && c->lvl_fill == c->ilu_fill && c->lvl_jc == c->ilu_jc) return 0;
  std::vector<int> tab;
  if (c->ilu_fill == 1) ilu1_level_tables(c->g, c->slab_w, tab);
  else ilu_level_tables(c->g, c->slab_w, tab, c->ilu_jc);
  if (tab.size() > c->lvl_cap) {
    AX1_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_lvl) { AX1_CUDA(cudaFree(c->d_lvl)); c->d_lvl = nullptr; }
    c->lvl_cap = std::max<size_t>(tab.size(), 2 * (size_t)(192 + 3 * c->g.nvy + 1));
    AX1_CUDA(cudaMalloc((void**)&c->d_lvl, c->lvl_cap * sizeof(int)));
  }
  AX1_CUDA(cudaMemcpyAsync(c->d_lvl, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));  // tab is a local
  c->lvl_slab_w = c->slab_w;
  c->lvl_fill = c->ilu_fill;
  c->lvl_jc = c->ilu_jc;
  return 0;
}

// Where to cut the sub-slabs radially (solver.cu: ypart_select). The radial grid of the reference is refined
// geometrically towards the membrane and uniform in the far field; a cut inside the uniform far field drops
// couplings that are no stronger than the ones the sub-slab borders drop in x -- measured on the paper grid (oracle,
// 40 time steps): 150 Krylov iterations without a cut, 150 with a cut at rows 86..120, 160 at row 78, 336 at 74,
// 4728 at 60. Automatic choice: the first row of the far field (spacing >= half the largest), moved up to the
// middle of the grid if that balances the two parts; no cut if the longer part would still hold > 70 % of the rows.
static int auto_radial_cut(const ax1gpu_ctx* c) {
  const int nvy = c->g.nvy;
  if (nvy < 32) return 0;
  double dmax = 0.0;
  for (int j = 0; j + 1 < nvy; ++j) dmax = std::max(dmax, c->hy[j + 1] - c->hy[j]);
  int js = -1;
  for (int j = 0; j + 1 < nvy; ++j) {
    if (c->hy[j + 1] - c->hy[j] >= 0.5 * dmax) { if (js < 0) js = j; }
    else js = -1;            // the far field is the LAST run of coarse cells
  }
  if (js < 0) return 0;
  const int jc = std::max(js, (nvy + 1) / 2);
  if (jc < 8 || nvy - jc < 8) return 0;
  if (std::max(jc, nvy - jc) > (7 * nvy) / 10) return 0;
  return jc;
}

// (re)select the preconditioner: fill 0 = block ILU0, fill 1 = block ILU(1) (SeqILUn with n = 1, the
// reference default, acme2_cyl_setup.hh:764-765); the ILU(1) factor needs 13 instead of 9 blocks per row
#define AX1_SLAB_KEEP (-1)
static int select_preconditioner(ax1gpu_ctx* c, int slab_w, int fill, int amg = -1) {
  const bool use_amg = amg >= 0 ? amg != 0 : c->amg != 0;
  // slab_w: > 0 the caller's choice, 0 library default, AX1_SLAB_KEEP (internal callers: drv_solve, drv_time_kernel)
  // whatever the handle currently uses -- an automatically derived width then stays automatic
  bool picked_auto = false, one_warp_regime = false;
  if (slab_w == AX1_SLAB_KEEP) {
    if (c->slab_w > 0 && !c->slab_auto) slab_w = c->slab_w;
    else slab_w = 0;
  }
  if (fill < 0) fill = c->ilu_fill;
  if (fill > 1) { set_error("block ILU(n) is implemented on the device for n = 0 and n = 1"); return AX1GPU_EINVAL; }
  if (slab_w <= 0) {
    // library default: 128 columns per sub-slab when that still gives >= 2 CTAs per SM, otherwise narrower
    // slabs (>= 16 columns = one warp per CTA) -- small grids are bound by the sweep's per-level latency, and
    // more, narrower slabs shorten the wavefront count and keep more SMs busy
    const int gran = fill == 1 ? 24 : 16;   // columns one warp sweeps: 8 rows per level = w/2 (ILU0) or w/3 (ILU(1))
    int w = (c->g.nvx + 295) / 296;
    w = ((w + gran - 1) / gran) * gran;
    // up to four one-warp CTAs share an SM without slowing each other down, and the one-warp sweep (warp-level barrier
    // only) is the fastest per level: on the myelin grid (7,252 columns) 16 instead of 32 columns gave 0.336 vs 0.373 ms
    // per sweep and 16.8 vs 19.5 s for 20 converged time steps (ILU(1), 24 vs 48 columns: 23.3 vs 29.0 s)
    // (not under the multigrid cycle, whose smoother gains more from wider sub-slabs: 12.1 vs 10.5 s there)
    if (!use_amg && c->g.nvx <= gran * 4 * 148) { w = gran; one_warp_regime = true; }
    // a width the caller chose explicitly earlier stays; an automatic one is re-derived (fill / multigrid may differ)
    if (c->slab_w > 0 && !c->slab_auto) slab_w = c->slab_w;
    else { slab_w = std::min(128, std::max(gran, w)); picked_auto = true; }
  }
  // one quad per wavefront row with <= 256-thread CTAs: levels are ceil(w/2) (ILU0) / ceil(w/3) (ILU(1)) rows wide
  const int wmax = fill == 1 ? ilu1_max_slab_w(c->g) : 128;
  if (slab_w > wmax) slab_w = wmax;
  if (slab_w > c->g.nvx) slab_w = c->g.nvx;
  // radial split of the sub-slabs: block ILU0 only; automatic on latency-bound (narrow) sub-slabs without multigrid
  int jc = 0;
  if (fill == 0 && !use_amg) {
    if (c->ilu_split > 0) jc = c->ilu_split;
    // automatic only together with an automatic sub-slab width: a caller who fixes the width gets exactly the
    // sub-slab ILU0 it asked for unless it also asks for the cut (ax1gpu_set_ilu_split)
    // ... and only where the one-warp sub-slabs fit the GPU in one wave: there a sweep costs its dependency depth and
    // the cut halves it. With more sub-slabs than that the SMs are busy anyway, and on the axially refined grid
    // (dx = 100 nm, 11,648 columns, 48-column sub-slabs) the cut cost 30 instead of 25 Krylov iterations per
    // converged step and 0.515 instead of 0.474 ms per sweep.
    else if (c->ilu_split < 0 && picked_auto && one_warp_regime) jc = auto_radial_cut(c);
    if (jc > 0 && (jc < 2 || c->g.nvy - jc < 2)) { set_error("ilu split row must leave two rows on either side"); return AX1GPU_EINVAL; }
  }
  const size_t need = (size_t)c->g.nv * (fill == 1 ? 208 : 144);
  if (need > c->lu_cap) {
    AX1_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_LU) { AX1_CUDA(cudaFree(c->d_LU)); c->d_LU = nullptr; c->lu_cap = 0; }
    AX1_CUDA(cudaMalloc((void**)&c->d_LU, need * sizeof(double)));
    c->lu_cap = need;
  }
  if (slab_w != c->slab_w || fill != c->ilu_fill || jc != c->ilu_jc) c->factored = false;
  c->ilu_jc = jc;
  c->slab_w = slab_w;
  c->slab_auto = picked_auto;
  c->ilu_fill = fill;
  return ensure_level_tables(c);
}

// ---- aggregation multigrid (see amg.cu) ------------------------------------------------------
static void amg_free(ax1gpu_ctx* c) {
  for (auto& L : c->amg_levels) {
    cudaFree(L.A); cudaFree(L.LU); cudaFree(L.b); cudaFree(L.x); cudaFree(L.lvl);
  }
  c->amg_levels.clear();
}
void drv_amg_free(ax1gpu_ctx* c) {
  amg_free(c);
  if (c->d_amg_t) { cudaFree(c->d_amg_t); c->d_amg_t = nullptr; }
}

// allocate the hierarchy: nvx -> ceil(nvx / cx) -> ... until a level has at most 2 cx vertex columns
static int amg_allocate(ax1gpu_ctx* c) {
  if (!c->amg_levels.empty() && c->amg_slab_w == c->slab_w) return 0;
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  amg_free(c);
  if (!c->d_amg_t) AX1_CUDA(cudaMalloc((void**)&c->d_amg_t, 4 * (size_t)c->g.nv * sizeof(double)));
  int nvx = c->g.nvx;
  while (nvx > 2 * c->amg_cx) {
    ax1gpu_ctx::AmgLevel L;
    nvx = (nvx + c->amg_cx - 1) / c->amg_cx;
    L.g = c->g;
    L.g.nx = nvx - 1; L.g.nvx = nvx; L.g.nv = nvx * c->g.nvy;
    L.g.dmask = nullptr;
    L.slab_w = std::min(c->slab_w, nvx);
    const size_t nv = L.g.nv;
    AX1_CUDA(cudaMalloc((void**)&L.A, nv * AX1_ROW * sizeof(double)));
    AX1_CUDA(cudaMalloc((void**)&L.LU, nv * 144 * sizeof(double)));
    AX1_CUDA(cudaMalloc((void**)&L.b, 4 * nv * sizeof(double)));
    AX1_CUDA(cudaMalloc((void**)&L.x, 4 * nv * sizeof(double)));
    std::vector<int> tab;
    ilu_level_tables(L.g, L.slab_w, tab);
    AX1_CUDA(cudaMalloc((void**)&L.lvl, tab.size() * sizeof(int)));
    AX1_CUDA(cudaMemcpy(L.lvl, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->amg_levels.push_back(L);
  }
  c->amg_slab_w = c->slab_w;
  return 0;
}

// Galerkin coarse operators + their ILU0 factors (after the level-0 factorisation)
static int amg_setup(ax1gpu_ctx* c) {
  int rc = amg_allocate(c);
  if (rc) return rc;
  const double* Af = c->d_A;
  int nvx_f = c->g.nvx;
  const uint8_t* mask = c->g.dmask;
  for (auto& L : c->amg_levels) {
    launch_galerkin_x(nvx_f, L.g.nvx, c->g.nvy, c->amg_cx, Af, L.A, mask, c->stream);
    AX1_LAUNCH_CHECK(c);
    launch_ilu_factor(L.g, L.A, L.LU, L.slab_w, L.lvl, c->stream);
    AX1_LAUNCH_CHECK(c);
    Af = L.A; nvx_f = L.g.nvx; mask = nullptr;
  }
  return 0;
}

static int smoother0(ax1gpu_ctx* c, const double* d, double* v) {  // level-0 smoother: block ILU(fill)
  if (c->ilu_fill == 1) launch_ilu1_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream);
  else launch_ilu_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream, nullptr, c->ilu_jc);
  AX1_LAUNCH_CHECK(c);
  return 0;
}

// v = V(1,1)-cycle applied to d (zero initial guess on every level). Level 0 is (g, d_A, d_LU); level l >= 1
// is amg_levels[l-1]. Scratch of level-0 size serves every level: d_amg_t = defect, d_tmp = smoother output.
static int amg_vcycle(ax1gpu_ctx* c, const double* d, double* v) {
  const int nl = (int)c->amg_levels.size();
  const int cx = c->amg_cx, nvy = c->g.nvy;
  struct View { Dev g; const double* A; const double* b; double* x; const uint8_t* mask; };
  auto view = [&](int l) -> View {
    if (l == 0) return View{c->g, c->d_A, d, v, c->g.dmask};
    auto& L = c->amg_levels[l - 1];
    return View{L.g, L.A, L.b, L.x, nullptr};
  };
  auto smooth = [&](int l, const double* rhs, double* out) -> int {
    if (l == 0) return smoother0(c, rhs, out);
    auto& L = c->amg_levels[l - 1];
    launch_ilu_apply(L.g, L.LU, rhs, out, L.slab_w, L.lvl, c->stream);
    AX1_LAUNCH_CHECK(c);
    return 0;
  };
  int rc;
  if ((rc = smooth(0, d, v))) return rc;                       // pre-smoothing from zero
  for (int l = 0; l < nl; ++l) {                               // down
    const View F = view(l);
    auto& C = c->amg_levels[l];
    launch_spmv_residual(F.g, F.A, F.x, F.b, c->d_amg_t, 0, c->stream);
    AX1_LAUNCH_CHECK(c);
    launch_restrict_x(F.g.nvx, C.g.nvx, nvy, cx, c->d_amg_t, C.b, F.mask, c->stream);
    AX1_LAUNCH_CHECK(c);
    if ((rc = smooth(l + 1, C.b, C.x))) return rc;
  }
  for (int l = nl - 1; l >= 0; --l) {                          // up
    const View F = view(l);
    auto& C = c->amg_levels[l];
    launch_prolong_add_x(F.g.nvx, C.g.nvx, nvy, cx, C.x, F.x, F.mask, c->stream);
    AX1_LAUNCH_CHECK(c);
    launch_spmv_residual(F.g, F.A, F.x, F.b, c->d_amg_t, 0, c->stream);
    AX1_LAUNCH_CHECK(c);
    if ((rc = smooth(l, c->d_amg_t, c->d_tmp))) return rc;      // post-smoothing
    launch_vec_add(4 * F.g.nv, F.x, c->d_tmp, c->stream);
    AX1_LAUNCH_CHECK(c);
  }
  return 0;
}

#define AX1_COMPACT_MIN_NV (1 << 20)
static bool compact_enabled(const ax1gpu_ctx* c) {   // AX1_SPMV_COMPACT=0/1 overrides the automatic choice (A/B runs)
  static const int env = getenv("AX1_SPMV_COMPACT") ? atoi(getenv("AX1_SPMV_COMPACT")) : -1;
  const int mode = c->spmv_compact >= 0 ? c->spmv_compact : env;
  return mode < 0 ? c->g.nv >= AX1_COMPACT_MIN_NV : mode != 0;
}
static inline const double* solver_Ac(const ax1gpu_ctx* c) { return (c->factored && c->ac_valid) ? c->d_Ac : nullptr; }

int drv_factor(ax1gpu_ctx* c, int slab_w, int fill, int amg) {
  int rc = select_preconditioner(c, slab_w, fill, amg);
  if (rc) return rc;
  if (amg >= 0) c->amg = amg ? 1 : 0;
  if (c->ilu_fill == 1) launch_ilu1_factor(c->g, c->d_A, c->d_LU, c->slab_w, c->d_lvl, c->stream);
  else launch_ilu_factor(c->g, c->d_A, c->d_LU, c->slab_w, c->d_lvl, c->stream, c->ilu_jc);
  AX1_LAUNCH_CHECK(c);
  if (c->amg && (rc = amg_setup(c))) return rc;
  // The solver multiplies with A 40 times per Newton iteration: on HBM-bound grids it does so with a copy in
  // 64-byte blocks (576 instead of 720 B per block row; solver.cu: compact_rows_kernel), built here because every
  // change of A passes through a new factorisation. Small grids are latency-bound and skip it.
  c->ac_valid = false;
  if (compact_enabled(c)) {
    if (!c->d_Ac) AX1_CUDA(cudaMalloc((void**)&c->d_Ac, ((size_t)c->g.nv + 8) * AX1_CROW * sizeof(double)));
    AX1_CUDA(cudaMemsetAsync(c->d_err + 3, 0, sizeof(int), c->stream));
    launch_compact_rows(c->g, c->d_A, c->d_Ac, c->d_err + 3, c->stream);
    AX1_LAUNCH_CHECK(c);
    AX1_CUDA(cudaMemcpyAsync(c->h_err + 3, c->d_err + 3, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    AX1_CUDA(cudaStreamSynchronize(c->stream));
    c->ac_valid = c->h_err[3] == 0;     // potential rows not exactly q x valences (FD Jacobians): multiply with A itself
  }
  c->factored = true;
  return 0;
}

// HaloPush descriptor of the next halo exchange through the peer windows
static HaloPush next_halo_push(ax1gpu_ctx* c) {
  const size_t H = c->p2p.halo_len;
  HaloPush hp{};
  hp.seq = ++c->p2p.halo_seq;
  const size_t par = (size_t)(hp.seq & 1);
  hp.cnt = c->p2p.cnt;
  halo_contributors(c->g, c->slab_w, &hp.ncontrib[0], &hp.ncontrib[1], c->ilu_jc);
  if (c->left_pb) {   // my left columns -> slot "from the right neighbour" (side 1) of rank-1
    double* pw = c->p2p.peer[c->rank - 1];
    hp.dst[0] = pw + AX1_P2P_HALO_OFF + (2 + par) * H;
    hp.flag[0] = reinterpret_cast<unsigned long long*>(pw + AX1_P2P_FLAG_OFF + 1);
  }
  if (c->right_pb) {  // my right columns -> slot "from the left neighbour" (side 0) of rank+1
    double* pw = c->p2p.peer[c->rank + 1];
    hp.dst[1] = pw + AX1_P2P_HALO_OFF + (0 + par) * H;
    hp.flag[1] = reinterpret_cast<unsigned long long*>(pw + AX1_P2P_FLAG_OFF + 0);
  }
  return hp;
}
static int halo_wait_add(ax1gpu_ctx* c, double* v, unsigned long long seq) {
  const Dev& g = c->g;
  const size_t H = c->p2p.halo_len, par = (size_t)(seq & 1);
  const int n = 3 * g.nvy * 4, bs = 256, nb = (n + bs - 1) / bs;
  double* w = c->p2p.win;
  if (c->left_pb) {
    halo_wait_add_kernel<<<nb, bs, 0, c->stream>>>(g, v, w + AX1_P2P_HALO_OFF + (0 + par) * H,
        reinterpret_cast<const unsigned long long*>(w + AX1_P2P_FLAG_OFF + 0), seq, 0, c->d_err + 2,
        c->p2p.timeout_cycles, c->d_sc);
    AX1_LAUNCH_CHECK(c);
  }
  if (c->right_pb) {
    halo_wait_add_kernel<<<nb, bs, 0, c->stream>>>(g, v, w + AX1_P2P_HALO_OFF + (2 + par) * H,
        reinterpret_cast<const unsigned long long*>(w + AX1_P2P_FLAG_OFF + 1), seq, g.nvx - 3, c->d_err + 2,
        c->p2p.timeout_cycles, c->d_sc);
    AX1_LAUNCH_CHECK(c);
  }
  return 0;
}

int drv_prec_apply(ax1gpu_ctx* c, const double* d, double* v) {
  const bool p2p = c->nranks > 1 && c->p2p.on;
  if (p2p && c->p2p.broken) return drv_check_p2p(c);
  if (c->amg) {
    // per-rank multigrid cycle (the rank-local matrix incl. its overlap), then the additive halo sum
    int rc = amg_vcycle(c, d, v);
    if (rc) return rc;
    if (p2p) {
      const HaloPush hp = next_halo_push(c);
      launch_halo_push(c->g, v, hp, c->stream);
      AX1_LAUNCH_CHECK(c);
      return halo_wait_add(c, v, hp.seq);
    }
    return halo_add(c, v);
  }
  if (p2p) {
    // sweep with fused halo push into the neighbours' windows, then acquire + add what they pushed to me
    const HaloPush hp = next_halo_push(c);
    if (c->ilu_fill == 1) launch_ilu1_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream, &hp);
    else launch_ilu_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream, &hp, c->ilu_jc);
    AX1_LAUNCH_CHECK(c);
    return halo_wait_add(c, v, hp.seq);
  }
  if (c->ilu_fill == 1) launch_ilu1_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream);
  else launch_ilu_apply(c->g, c->d_LU, d, v, c->slab_w, c->d_lvl, c->stream, nullptr, c->ilu_jc);
  AX1_LAUNCH_CHECK(c);
  return halo_add(c, v);
}

int drv_op_apply(ax1gpu_ctx* c, const double* x, double* y) {
  launch_spmv(c->g, c->d_A, x, y, c->nranks > 1, c->stream, solver_Ac(c));
  AX1_LAUNCH_CHECK(c);
  return 0;
}

// A peer-window exchange that timed out (p2p.cuh) has set d_err[2], stopped the solve (Scalars::done) and left NaN
// sums. Report it where the host syncs anyway; afterwards the ranks' sequence counters disagree, so the handle's
// communication is marked broken and every later exchange fails immediately instead of reading stale windows.
static const char* P2P_TIMEOUT_MSG =
    "peer-memory exchange timed out (a neighbour rank did not arrive within AX1_P2P_TIMEOUT_S); "
    "the handle cannot communicate any more -- destroy it on all ranks";
int drv_check_p2p(ax1gpu_ctx* c) {
  if (c->nranks <= 1 || !c->p2p.on) return 0;
  if (c->p2p.broken) { set_error(P2P_TIMEOUT_MSG); return AX1GPU_ENCCL; }
  if (c->h_err[2]) {
    c->h_err[2] = 0;
    c->p2p.broken = true;
    cudaMemsetAsync(c->d_err + 2, 0, sizeof(int), c->stream);   // reported once; `broken` keeps the handle shut
    set_error(P2P_TIMEOUT_MSG);
    return AX1GPU_ENCCL;
  }
  return 0;
}

static int poll_scalars(ax1gpu_ctx* c) {
  AX1_CUDA(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
  if (c->nranks > 1 && c->p2p.on)
    AX1_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return drv_check_p2p(c);
}

// BiCGStab on A z = rhs from z = 0; rhs is overwritten with the final residual (ISTL aliases r = b)
int drv_solve(ax1gpu_ctx* c, double* r, double* x, double reduction, int maxit, ax1gpu_lin_result* res) {
  const int nv = c->g.nv;
  const size_t bytes = 4 * (size_t)nv * sizeof(double);
  const Own own = own_of(c);
  const int G = vec_grid(c);
  int rc;
  if (!c->factored) { rc = drv_factor(c, AX1_SLAB_KEEP); if (rc) return rc; }
  AX1_CUDA(cudaMemsetAsync(x, 0, bytes, c->stream));
  AX1_CUDA(cudaMemsetAsync(c->d_sc, 0, sizeof(Scalars), c->stream));
  // r = b - A*0 = b ; rt = r
  AX1_CUDA(cudaMemcpyAsync(c->d_rt, r, bytes, cudaMemcpyDeviceToDevice, c->stream));
  dot_kernel<<<G, 256, 0, c->stream>>>(nv, r, r, nullptr, nullptr, 0, own, nullptr, c->d_partials);
  AX1_LAUNCH_CHECK(c);
  if ((rc = reduce_op(c, G, 1, OP_NORM0, reduction))) return rc;
  int k = 0;
  // The convergence flag is read back after every half step (ISTL tests there too): a poll costs ~15 us of stream
  // drain, a preconditioner sweep never less than ~300 us (its depth is >= 2 Ny wavefronts), and the converged runs
  // need 2-3 iterations per Newton step -- polling every 4th iteration ran 8 sweeps where 4 were needed.
  // Fixed-work runs (reduction = 0) never poll.
  const bool poll = reduction > 0.0;
  const int SB = spmv_blocks(c->g);
  const int zc = c->nranks > 1;
  const double* Ac = solver_Ac(c);
  for (k = 0; 0.5 + k < maxit; ++k) {
    // rho_new, the breakdown tests and beta were produced by OP_NORM0 / OP_NORM_B of the step before
    pupdate_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_p, c->d_v, r, c->d_sc, k == 0);
    AX1_LAUNCH_CHECK(c);
    if ((rc = drv_prec_apply(c, c->d_p, c->d_y1))) return rc;
    launch_spmv_dot(c->g, c->d_A, c->d_y1, c->d_v, zc, 1, c->d_rt, own.lo, own.hi, own.masked, c->d_sc, c->d_spartials,
                    c->stream, Ac);                               // v = A y, h = <rt, v>
    AX1_LAUNCH_CHECK(c);
    if ((rc = reduce_op_from(c, c->d_spartials, SB, 1, OP_H, reduction))) return rc;
    r_update_kernel<<<G, 256, 0, c->stream>>>(nv, r, c->d_v, c->d_sc, own, c->d_partials);     // x += alpha y: deferred
    AX1_LAUNCH_CHECK(c);
    if ((rc = reduce_op(c, G, 1, OP_NORM_A, reduction))) return rc;
    if (poll) {
      if ((rc = poll_scalars(c))) return rc;
      if (c->h_sc->done) break;
    }
    if ((rc = drv_prec_apply(c, r, c->d_y2))) return rc;
    launch_spmv_dot(c->g, c->d_A, c->d_y2, c->d_t, zc, 2, r, own.lo, own.hi, own.masked, c->d_sc, c->d_spartials,
                    c->stream, Ac);                               // t = A y, <t, r>, <t, t>
    AX1_LAUNCH_CHECK(c);
    if ((rc = reduce_op_from(c, c->d_spartials, SB, 2, OP_OMEGA, reduction))) return rc;
    xr_update_kernel<<<G, 256, 0, c->stream>>>(nv, x, c->d_y1, c->d_y2, r, c->d_t, c->d_rt, c->d_sc, own, c->d_partials);
    AX1_LAUNCH_CHECK(c);
    if ((rc = reduce_op(c, G, 2, OP_NORM_B, reduction))) return rc;
    if (poll) {
      if ((rc = poll_scalars(c))) return rc;
      if (c->h_sc->done) break;
    }
  }
  if ((rc = poll_scalars(c))) return rc;
  const Scalars& s = *c->h_sc;
  if (s.x_pending) {   // converged at the half step: the kernels after it did nothing, y1 is still M^-1 p of that step
    x_fixup_kernel<<<G, 256, 0, c->stream>>>(nv, x, c->d_y1, c->d_sc);
    AX1_LAUNCH_CHECK(c);
  }
  // ISTL tests for a break-down at the top of the NEXT iteration; one flagged by the last permitted
  // iteration is therefore not an error, the solve just ends unconverged
  if (s.breakdown && s.it_half < maxit) { set_error("breakdown in BiCGSTAB"); res->converged = 0; return AX1GPU_ESTATE; }
  double it = s.it_half;
  if (it > maxit) it = maxit;
  res->converged = s.converged;
  res->it_half = it;
  res->iterations = (int)std::ceil(it);
  res->reduction = s.norm0 > 0 ? s.norm / s.norm0 : 0.0;
  return 0;
}


// ------------------------------------------------------------------------------------------
// Restarted GMRES on A z = rhs from z = 0: Dune::RestartedGMResSolver (dune-istl 2.3 solvers.hh) as driven by
// ISTLBackend_GMRES_AMG (dune/ax1/common/ax1_solverbackend.hh:196-213: restart 100, recalcDefect = false) [ext]:
// left preconditioning, modified Gram-Schmidt, Givens rotations, convergence on the preconditioned residual.
// The Hessenberg matrix, rotations and convergence flag live on the device; the host polls the flag.
int drv_solve_gmres(ax1gpu_ctx* c, double* b, double* x, double reduction, int restart, int maxit,
                    ax1gpu_lin_result* res) {
  const int nv = c->g.nv;
  const size_t len = 4 * (size_t)nv, bytes = len * sizeof(double);
  const Own own = own_of(c);
  const int G = vec_grid(c);
  const int zc = c->nranks > 1;
  int rc;
  int m = restart <= 0 ? 100 : restart;
  if (m > 1000) { set_error("gmres: restart must be <= 1000"); return AX1GPU_EINVAL; }
  if (m > maxit && maxit > 0) m = maxit;
  if (m > c->gm_cap) {
    AX1_CUDA(cudaStreamSynchronize(c->stream));
    if (c->d_gmV) { cudaFree(c->d_gmV); c->d_gmV = nullptr; }
    if (c->d_gm) { cudaFree(c->d_gm); c->d_gm = nullptr; }
    c->gm_cap = 0;
    AX1_CUDA(cudaMalloc((void**)&c->d_gmV, (size_t)(m + 1) * bytes));
    AX1_CUDA(cudaMalloc((void**)&c->d_gm, gm_len(m) * sizeof(double)));
    c->gm_cap = m;
  }
  if (!c->factored) { rc = drv_factor(c, AX1_SLAB_KEEP); if (rc) return rc; }
  double* V = c->d_gmV;
  double* gm = c->d_gm;
  double* w = c->d_p;
  auto Vk = [&](int k) { return V + (size_t)k * len; };
  AX1_CUDA(cudaMemsetAsync(x, 0, bytes, c->stream));
  AX1_CUDA(cudaMemsetAsync(c->d_sc, 0, sizeof(Scalars), c->stream));
  AX1_CUDA(cudaMemsetAsync(gm, 0, gm_len(m) * sizeof(double), c->stream));
  c->h_pin[0] = (double)m;
  AX1_CUDA(cudaMemcpyAsync(gm + GM_M, c->h_pin, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  // b - A*0 = b ; v0 = M^-1 b ; beta = |v0|
  if ((rc = drv_prec_apply(c, b, Vk(0)))) return rc;
  gmres_mgs_kernel<<<G, 256, 0, c->stream>>>(nv, Vk(0), nullptr, nullptr, gm, c->d_sc, own, c->d_partials);
  AX1_LAUNCH_CHECK(c);
  if ((rc = reduce_op(c, G, 1, OP_G_BETA, reduction, gm, 1))) return rc;
  if ((rc = poll_scalars(c))) return rc;
  int j = 1;
  bool stop = c->h_sc->done != 0;
  while (j <= maxit && !stop) {
    gmres_scale_kernel<<<G, 256, 0, c->stream>>>(nv, Vk(0), Vk(0), gm, c->d_sc);
    AX1_LAUNCH_CHECK(c);
    for (int i = 0; i < m && j <= maxit; ++i, ++j) {
      launch_spmv(c->g, c->d_A, Vk(i), Vk(i + 1), zc, c->stream, solver_Ac(c));      // v[i+1] as temporary
      AX1_LAUNCH_CHECK(c);
      if ((rc = drv_prec_apply(c, Vk(i + 1), w))) return rc;
      for (int k = 0; k <= i; ++k) {
        gmres_mgs_kernel<<<G, 256, 0, c->stream>>>(nv, w, k ? Vk(k - 1) : nullptr, Vk(k), gm, c->d_sc, own, c->d_partials);
        AX1_LAUNCH_CHECK(c);
        if ((rc = reduce_op(c, G, 1, OP_G_H, reduction, gm, (k << 10) | i))) return rc;
      }
      gmres_mgs_kernel<<<G, 256, 0, c->stream>>>(nv, w, Vk(i), nullptr, gm, c->d_sc, own, c->d_partials);
      AX1_LAUNCH_CHECK(c);
      if ((rc = reduce_op(c, G, 1, OP_G_NORMW, reduction, gm, i))) return rc;
      gmres_scale_kernel<<<G, 256, 0, c->stream>>>(nv, Vk(i + 1), w, gm, c->d_sc);
      AX1_LAUNCH_CHECK(c);
      if (reduction > 0.0) {  // every iteration: a poll is far cheaper than one more preconditioner application
        if ((rc = poll_scalars(c))) return rc;
        if (c->h_sc->done) { ++j; break; }
      }
    }
    // x += sum_c y_c v_c over the steps this cycle completed (count kept on the device)
    gmres_ysolve_kernel<<<1, 1, 0, c->stream>>>(gm);
    AX1_LAUNCH_CHECK(c);
    gmres_combine_kernel<<<G, 256, 0, c->stream>>>(nv, V, len, gm, w, x);
    AX1_LAUNCH_CHECK(c);
    // b -= A w (ISTL keeps the defect in b)
    launch_spmv_residual(c->g, c->d_A, w, b, c->d_t, zc, c->stream);
    AX1_LAUNCH_CHECK(c);
    AX1_CUDA(cudaMemcpyAsync(b, c->d_t, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = poll_scalars(c))) return rc;
    if (c->h_sc->done) break;
    if (j > maxit) break;
    if ((rc = drv_prec_apply(c, b, Vk(0)))) return rc;
    gmres_mgs_kernel<<<G, 256, 0, c->stream>>>(nv, Vk(0), nullptr, nullptr, gm, c->d_sc, own, c->d_partials);
    AX1_LAUNCH_CHECK(c);
    if ((rc = reduce_op(c, G, 1, OP_G_BETA, reduction, gm, 0))) return rc;
    if ((rc = poll_scalars(c))) return rc;
    stop = c->h_sc->done != 0;
  }
  if ((rc = poll_scalars(c))) return rc;
  const Scalars& s = *c->h_sc;
  if (s.breakdown) { set_error("breakdown in GMRes - |w| == 0.0"); res->converged = 0; return AX1GPU_ESTATE; }
  res->converged = s.converged;
  res->it_half = s.it_half;
  res->iterations = (int)s.it_half;
  res->reduction = s.norm0 > 0 ? s.norm / s.norm0 : 0.0;
  return 0;
}

// ------------------------------------------------------------------------------------------
static float ev_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int drv_newton(ax1gpu_ctx* c, const ax1gpu_newton_opts* o, ax1gpu_newton_result* res) {
  const int nv = c->g.nv;
  const size_t bytes = 4 * (size_t)nv * sizeof(double);
  const int G = vec_grid(c);
  std::memset(res, 0, sizeof(*res));
  res->reduction = 1.0; res->conv_rate = 1.0;
  int rc;
  double t_asm_ms = 0, t_lin_ms = 0;
  cudaEvent_t e0 = c->ev[0], e1 = c->ev[1], e2 = c->ev[2], e3 = c->ev[3];
  AX1_CUDA(cudaEventRecord(e3, c->stream));
  auto defect = [&](bool& finite) -> int {
    double nrm = 0.0;
    int r_ = drv_residual(c, c->d_u, c->d_r, &nrm);
    if (r_) return r_;
    res->residual_evals++;
    res->defect = nrm;
    finite = std::isfinite(nrm);
    return 0;
  };
  bool finite = true;
  if ((rc = defect(finite))) return rc;
  if (!finite) { res->status = AX1GPU_NEWTON_DEFECT_NAN; return 0; }
  res->first_defect = res->defect;
  double prev_defect = res->defect;
  double linear_reduction = o->min_lin_reduction;
  while (true) {
    if (!(o->force_iteration && res->iterations == 0)) {
      res->converged = res->defect < o->abs_limit || res->defect < res->first_defect * o->reduction;
      if (o->fixed_work) res->converged = res->iterations >= o->maxit;
      if (res->converged) break;
      if (res->iterations >= o->maxit) { res->status = AX1GPU_NEWTON_NOT_CONVERGED; break; }
    }
    // prepare_step: preAssembly (gating) + Jacobian + linear reduction
    AX1_CUDA(cudaEventRecord(e0, c->stream));
    if ((rc = drv_update_state(c, c->d_u))) return rc;
    if ((rc = drv_jacobian(c, c->d_u))) return rc;
    if (o->fixed_lin_reduction) linear_reduction = o->min_lin_reduction;
    else {
      const double stop_defect = std::max(res->first_defect * o->reduction, o->abs_limit);
      if (stop_defect / (10 * res->defect) > res->defect * res->defect / (prev_defect * prev_defect))
        linear_reduction = stop_defect / (10 * res->defect);
      else
        linear_reduction = std::min(o->min_lin_reduction, res->defect * res->defect / (prev_defect * prev_defect));
    }
    prev_defect = res->defect;
    if (c->prm.use_rownorm_preconditioner) { launch_rownorm(c->g, c->d_A, c->d_r, c->stream); AX1_LAUNCH_CHECK(c); }
    AX1_CUDA(cudaEventRecord(e1, c->stream));
    // linear solve (ILU0 re-factorised every Newton iteration, like SeqILUn in the backend's apply)
    ax1gpu_lin_result lr{};
    if ((rc = drv_factor(c, o->slab_w, o->ilu_fill, o->amg))) return rc;
    if (o->krylov == 1)
      rc = drv_solve_gmres(c, c->d_r, c->d_z, o->fixed_work ? 0.0 : linear_reduction, o->restart, o->lin_maxit, &lr);
    else
      rc = drv_solve(c, c->d_r, c->d_z, o->fixed_work ? 0.0 : linear_reduction, o->lin_maxit, &lr);
    AX1_CUDA(cudaEventRecord(e2, c->stream));
    AX1_CUDA(cudaEventSynchronize(e2));
    t_asm_ms += ev_ms(e0, e1);
    t_lin_ms += ev_ms(e1, e2);
    if (rc == AX1GPU_ESTATE) { res->status = AX1GPU_NEWTON_LINSOLVE_FAILED; break; }
    if (rc) return rc;
    res->lin_iterations += lr.iterations;
    if (!lr.converged && !o->fixed_work) { res->status = AX1GPU_NEWTON_LINSOLVE_FAILED; break; }
    // line search, strategy hackbuschReuskenAcceptBestNoThrow (ax1_newton.hh:372-502)
    AX1_CUDA(cudaEventRecord(e0, c->stream));
    AX1_CUDA(cudaMemcpyAsync(c->d_uprev, c->d_u, bytes, cudaMemcpyDeviceToDevice, c->stream));
    auto step = [&](double lam) -> int {
      newton_update_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_u, c->d_uprev, c->d_z, lam);
      AX1_LAUNCH_CHECK(c);
      int r_ = drv_update_state(c, c->d_u);
      if (r_) return r_;
      return defect(finite);
    };
    if (o->disable_line_search || o->fixed_work) {
      if ((rc = step(1.0))) return rc;
      if (!finite) { res->status = AX1GPU_NEWTON_DEFECT_NAN; break; }
    } else {
      double lambda = 1.0, best_lambda = 0.0, best_defect = res->defect;
      int i = 0;
      while (true) {
        if ((rc = step(lambda))) return rc;
        if (res->defect <= (1.0 - lambda / 4) * prev_defect) break;
        if (res->defect < best_defect) { best_defect = res->defect; best_lambda = lambda; }
        if (++i >= o->line_search_maxit) {
          if (best_lambda == 0.0) {
            best_lambda = 1.0;
            if ((rc = step(best_lambda))) return rc;
            if (!std::isfinite(res->defect)) break;  // NewtonDefectError outside the search's try block (:469)
          }
          if (best_lambda != lambda) {
            if ((rc = step(best_lambda))) return rc;
          }
          break;
        }
        lambda *= 0.5;
      }
      if (!std::isfinite(res->defect)) { res->status = AX1GPU_NEWTON_DEFECT_NAN; break; }
    }
    AX1_CUDA(cudaEventRecord(e1, c->stream));
    AX1_CUDA(cudaEventSynchronize(e1));
    t_asm_ms += ev_ms(e0, e1);
    res->reduction = res->defect / res->first_defect;
    res->iterations++;
    res->conv_rate = std::pow(res->reduction, 1.0 / res->iterations);
  }
  AX1_CUDA(cudaEventRecord(e2, c->stream));
  AX1_CUDA(cudaEventSynchronize(e2));
  res->t_assembler = t_asm_ms * 1e-3;
  res->t_lin_solv = t_lin_ms * 1e-3;
  res->t_elapsed = ev_ms(e3, e2) * 1e-3;
  // channel-kernel error flag (leak ratio outside [0,1], channelset.hh:298-299)
  AX1_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  if (c->h_err[0]) { set_error("Calculated leak channel ratio outside [0,1]"); return AX1GPU_ESTATE; }
  return drv_check_p2p(c);
}

// mean device time of one kernel on resident data (roofline evidence for bench.py)
int drv_time_kernel(ax1gpu_ctx* c, const char* name, int reps, double* ms) {
  const std::string n(name);
  const int nv = c->g.nv;
  const int G = vec_grid(c);
  drv_refresh_dev(c);
  auto run = [&]() -> int {
    if (n == "residual") { launch_residual(c->g, c->d_u, c->d_r0, c->d_r, 0, c->stream); }
    else if (n == "jacobian") { launch_jacobian_volume(c->g, c->d_u, c->d_A, c->stream); }
    else if (n == "spmv") { launch_spmv(c->g, c->d_A, c->d_y2, c->d_v, 0, c->stream); }
    // the two variants BiCGStab launches: v = A y with <rt, v>, t = A y with <t, r> and <t, t> fused in
    else if (n == "spmv_dot1" || n == "spmv_dot2") {
      const Own own = own_of(c);
      launch_spmv_dot(c->g, c->d_A, c->d_y2, c->d_v, c->nranks > 1, n == "spmv_dot1" ? 1 : 2, c->d_rt, own.lo, own.hi,
                      own.masked, c->d_sc, c->d_spartials, c->stream, solver_Ac(c));
    }
    // the two exchange steps of the overlapping Krylov solve, as the solver issues them (all ranks must call with
    // the same reps): "reduce" = block partials -> sums over all ranks -> scalar recurrence (peer windows: one launch;
    // NCCL: reduce + allreduce + scalar kernel); "halo" = additive halo sum of the three shared vertex columns
    else if (n == "reduce") {
      dot_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_rt, c->d_v, nullptr, nullptr, 0, own_of(c), nullptr, c->d_partials);
      AX1_LAUNCH_CHECK(c);
      int r_ = reduce_op(c, G, 1, OP_PLAIN_NORM, 0.0, c->d_sums + 2);
      if (r_) return r_;
      return 0;
    }
    else if (n == "halo") {
      if (c->nranks > 1 && c->p2p.on) {
        const HaloPush hp = next_halo_push(c);
        launch_halo_push(c->g, c->d_y2, hp, c->stream);
        AX1_LAUNCH_CHECK(c);
        return halo_wait_add(c, c->d_y2, hp.seq);
      }
      return halo_add(c, c->d_y2);
    }
    else if (n == "compact") {
      if (!c->d_Ac) { set_error("time_kernel compact: no compact copy on this grid"); return AX1GPU_EINVAL; }
      launch_compact_rows(c->g, c->d_A, c->d_Ac, c->d_err + 3, c->stream);
    }
    else if (n == "pupdate") { pupdate_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_tmp, c->d_v, c->d_t, c->d_sc, 0); }
    else if (n == "ilu_factor") {
      if (c->ilu_fill == 1) launch_ilu1_factor(c->g, c->d_A, c->d_LU, c->slab_w, c->d_lvl, c->stream);
      else launch_ilu_factor(c->g, c->d_A, c->d_LU, c->slab_w, c->d_lvl, c->stream, c->ilu_jc);
    }
    else if (n == "ilu_apply") {
      if (c->ilu_fill == 1) launch_ilu1_apply(c->g, c->d_LU, c->d_p, c->d_y2, c->slab_w, c->d_lvl, c->stream);
      else launch_ilu_apply(c->g, c->d_LU, c->d_p, c->d_y2, c->slab_w, c->d_lvl, c->stream, nullptr, c->ilu_jc);
    }
    else if (n == "dot") { dot_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_rt, c->d_v, nullptr, nullptr, 0, own_of(c), nullptr, c->d_partials); }
    else if (n == "axpy") { xr_update_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_tmp, c->d_y1, c->d_y2, c->d_t, c->d_v, c->d_rt, c->d_sc, own_of(c), c->d_partials); }
    else if (n == "rupdate") { r_update_kernel<<<G, 256, 0, c->stream>>>(nv, c->d_t, c->d_v, c->d_sc, own_of(c), c->d_partials); }
    else { set_error("unknown kernel name"); return AX1GPU_EINVAL; }
    AX1_LAUNCH_CHECK(c);
    return 0;
  };
  { int rc0 = select_preconditioner(c, AX1_SLAB_KEEP, c->ilu_fill); if (rc0) return rc0; }
  AX1_CUDA(cudaMemsetAsync(c->d_sc, 0, sizeof(Scalars), c->stream));
  int rc = run();
  if (rc) return rc;
  AX1_CUDA(cudaEventRecord(c->ev[0], c->stream));
  for (int i = 0; i < reps; ++i) if ((rc = run())) return rc;
  AX1_CUDA(cudaEventRecord(c->ev[1], c->stream));
  AX1_CUDA(cudaEventSynchronize(c->ev[1]));
  *ms = ev_ms(c->ev[0], c->ev[1]) / reps;
  return 0;
}

}  // namespace ax1
