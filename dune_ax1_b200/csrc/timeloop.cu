// timeloop.cu -- the caller of the hot path on the device side: one adaptive time step of
// Acme2CylSimulation::run (dune/ax1/acme2_cyl/common/acme2_cyl_simulation.hh:236-932) with the iterate,
// the previous step and the channel state resident in HBM:
//   dt control   :334-343  x1.1 if the last Newton solve took < lowerLimit iterations and not more than the
//                          one before, /1.2 if it took >= upperLimit. No first-step exception: diagInfo.iterations
//                          starts at 0 (acme2_cyl_output.hh:169), so the very first adaptive step is already 1.1 dt0
//   equilibration :253-263 dt is held while doEquilibration && time < tEquilibrium; from the first time point
//                          >= tEquilibrium on the `equilibrationOver` flag stays set (:241-247 only ever sets it), so
//                          every later step runs with dt = dtstart -- restated as the reference's code does it
//   clamps       :362-363  [minTimeStep, maxTimeStep]
//   AP limiter   :365-383  min / max membrane potential over all membrane elements (+ comm().min / max),
//                          dt <= maxTimeStepAP while stimulating or while max V_m > -50 mV
//   stim start   :262-266  dt = dtstart on the step that crosses tInj_start
//   protocol     :457, 589, 916-927  updateState(time, dt) -> solver.apply -> uold = unew -> acceptTimeStep
//   restart      :593-638  Newton failure -> dt /= 2, updateState and solver.apply again on the *modified* unew
//                          ("it might already contain better starting values", :563-566), up to
//                          maxNumberNewtonRestarts. Deviations, both explicit: a NaN defect makes the iterate unusable,
//                          so that one case restarts from uold; opts.restart_from_uold = 1 does so after every failure
// Only three scalars per step (min / max V_m, the Newton result) cross PCIe.
#include <cmath>
#include <cstring>
#include <limits>
#include "ctx.h"

namespace ax1 {

// min / max of the membrane potential [mV] over the membrane columns (one block, fixed order)
__global__ void __launch_bounds__(1024) vm_minmax_kernel(int nx, const double* __restrict__ vm, double* out) {
  __shared__ double smin[32], smax[32];
  double lo = 1.79769313486231570e308, hi = -1.79769313486231570e308;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) { const double v = vm[i]; lo = fmin(lo, v); hi = fmax(hi, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { smin[w] = lo; smax[w] = hi; }
  __syncthreads();
  if (w == 0) {
    lo = smin[lane]; hi = smax[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { out[0] = -lo; out[1] = hi; }   // (-min, max): one ncclMax all-reduce serves both
  }
}

#define AX1_TL_CHECK(c)                                                      \
  do {                                                                       \
    (c)->launches++;                                                         \
    cudaError_t e_ = cudaGetLastError();                                     \
    if (e_ != cudaSuccess) { set_error(std::string("kernel launch: ") + cudaGetErrorString(e_)); return AX1GPU_ECUDA; } \
  } while (0)

// V_m extrema of the device iterate over all ranks
int drv_vm_minmax(ax1gpu_ctx* c, double* vmin, double* vmax) {
  launch_memb_pot(c->g, c->d_u, c->d_vm, c->stream);
  AX1_TL_CHECK(c);
  vm_minmax_kernel<<<1, 1024, 0, c->stream>>>(c->g.nx, c->d_vm, c->d_sums + 2);
  AX1_TL_CHECK(c);
  if (c->nranks > 1) {
    if (!c->comm) { set_error("V_m extrema over ranks need the NCCL communicator (ax1gpu_comm_init)"); return AX1GPU_ESTATE; }
    ncclResult_t r = c->nccl.AllReduce(c->d_sums + 2, c->d_sums + 2, 2, ncclDouble, ncclMax, c->comm, c->stream);
    if (r != ncclSuccess) { set_error(std::string("ncclAllReduce: ") + c->nccl.GetErrorString(r)); return AX1GPU_ENCCL; }
  }
  AX1_CUDA(cudaMemcpyAsync(c->h_pin, c->d_sums + 2, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  *vmin = -c->h_pin[0];
  *vmax = c->h_pin[1];
  return 0;
}

int drv_time_step(ax1gpu_ctx* c, const ax1gpu_timeloop_opts* o, const ax1gpu_newton_opts* no, ax1gpu_timeloop_state* s,
                  ax1gpu_step_result* out) {
  const ax1gpu_params& p = c->prm;
  const size_t bytes = 4 * (size_t)c->g.nv * sizeof(double);
  int rc;
  std::memset(out, 0, sizeof(*out));
  // ---- choose dt (acme2_cyl_simulation.hh:255-439)
  bool accept = !o->adaptive || (p.do_equilibration && s->time < p.t_equilibrium);
  if (p.do_equilibration && (s->time > p.t_equilibrium || std::fabs(s->time - p.t_equilibrium) < 1e-6)) {
    s->dt = o->dtstart;      // equilibrationOver (latched in the reference): dt reset to the start value
    accept = true;
  }
  if (p.do_stimulation && (s->time + s->dt > p.tinj_start && s->time < p.tinj_start)) {
    accept = true;
    s->dt = o->dtstart;
  }
  if (!accept) {
    if (s->its < o->lower_limit && s->its <= s->last_its) s->dt *= 1.1;
    if (s->its >= o->upper_limit) s->dt /= 1.2;
    s->dt = std::min(s->dt, o->max_dt);
    s->dt = std::max(s->dt, o->min_dt);
    double vmin, vmax;
    if ((rc = drv_vm_minmax(c, &vmin, &vmax))) return rc;
    const bool stim_now = p.do_stimulation && s->time + s->dt > p.tinj_start && s->time + s->dt < p.tinj_end;
    if (stim_now || vmax > o->vm_ap_threshold) s->dt = std::min(s->dt, o->max_dt_ap);
  }
  // ---- the step, with halve-and-retry on Newton failure
  ax1gpu_newton_result nr{};
  int restarts = 0;
  bool have_uold = false;
  while (true) {
    c->time = s->time; c->dt = s->dt;
    drv_refresh_dev(c);
    // the reference retries from the iterate the failed solve left behind; a NaN iterate cannot be used
    if (have_uold && (o->restart_from_uold || nr.status == AX1GPU_NEWTON_DEFECT_NAN))
      AX1_CUDA(cudaMemcpyAsync(c->d_u, c->d_uold, bytes, cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = drv_update_state(c, c->d_u))) return rc;              // gfMembFlux.updateState(time, dt)
    if (!have_uold) {
      AX1_CUDA(cudaMemcpyAsync(c->d_uold, c->d_u, bytes, cudaMemcpyDeviceToDevice, c->stream));
      have_uold = true;
    }
    if ((rc = drv_set_uold(c))) return rc;                          // OneStepMethod: r0 = -M(uold)
    if ((rc = drv_newton(c, no, &nr))) return rc;
    out->lin_iterations += nr.lin_iterations;
    out->t_assembler += nr.t_assembler;
    out->t_lin_solv += nr.t_lin_solv;
    if (nr.status == AX1GPU_NEWTON_OK) break;
    if (++restarts > o->max_restarts) {
      AX1_CUDA(cudaMemcpyAsync(c->d_u, c->d_uold, bytes, cudaMemcpyDeviceToDevice, c->stream));
      out->restarts = restarts; out->newton_status = nr.status; out->time = s->time; out->dt = s->dt;
      return 0;                                                    // caller sees newton_status != 0
    }
    s->dt *= 0.5;
  }
  s->last_its = s->its;
  s->its = nr.iterations;
  s->time += s->dt;
  s->steps += 1;
  if ((rc = drv_accept_state(c))) return rc;                        // gfMembFlux.acceptTimeStep()
  double vmin, vmax;
  if ((rc = drv_vm_minmax(c, &vmin, &vmax))) return rc;
  out->time = s->time; out->dt = s->dt;
  out->newton_iterations = nr.iterations;
  out->restarts = restarts;
  out->newton_status = 0;
  out->vm_min = vmin; out->vm_max = vmax;
  out->defect = nr.defect; out->first_defect = nr.first_defect;
  return 0;
}

}  // namespace ax1
