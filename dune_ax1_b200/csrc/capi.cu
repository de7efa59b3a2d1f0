// capi.cu -- the extern "C" surface declared in include/ax1gpu.h.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "ctx.h"

namespace ax1 {
static thread_local std::string g_error;
void set_error(const std::string& s) { g_error = s; }
}  // namespace ax1

using namespace ax1;

#define AX1_CHECK_H(h) \
  if (!(h)) { set_error("null handle"); return AX1GPU_EINVAL; }

template <class T>
static int dev_alloc(T** p, size_t n) {
  AX1_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return 0;
}
template <class T>
static int dev_upload(T** p, const T* src, size_t n, cudaStream_t st) {
  int rc = dev_alloc(p, n);
  if (rc) return rc;
  AX1_CUDA(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

static int up(ax1gpu_ctx* c, double* dst, const double* src) {
  AX1_CUDA(cudaMemcpyAsync(dst, src, 4 * (size_t)c->g.nv * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return 0;
}
static int down(ax1gpu_ctx* c, double* dst, const double* src) {
  AX1_CUDA(cudaMemcpyAsync(dst, src, 4 * (size_t)c->g.nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" {

const char* ax1gpu_last_error(void) { return g_error.c_str(); }

void ax1gpu_params_default(ax1gpu_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->temperature = 279.45; p->eps_elec = 80.0; p->length_scale = 1e-9; p->time_scale = 1e-6;
  p->diff_water[0] = 1.33e-9; p->diff_water[1] = 1.96e-9; p->diff_water[2] = 2.03e-9;
  p->valence[0] = 1; p->valence[1] = 1; p->valence[2] = -1;
  p->pot_residual_scale = 1.0;
  p->do_volume_scaling = 1;
  p->dummy_value = 1.0;
  p->membrane_active = 1; p->automatic_channel_init = 1; p->channel_resting_potential = -65.0;
  p->leak_ratio_na = 0.13; p->leak_ratio_k = 0.87;
  p->use_elec_jacobian_volume = p->use_memb_jacobian_volume = p->use_time_jacobian_volume = 1;
  p->do_memb_flux_stimulation = 0; p->memb_flux_ion = 2; p->memb_flux_inj = 0.0;
  p->mori_membrane_neumann = 1;
}

void ax1gpu_newton_opts_default(ax1gpu_newton_opts* o) {
  o->reduction = 1e-5; o->abs_limit = 1e-5; o->min_lin_reduction = 1e-5;
  o->fixed_lin_reduction = 0; o->maxit = 50; o->line_search_maxit = 10; o->lin_maxit = 5000;
  o->force_iteration = 0; o->disable_line_search = 0; o->slab_w = 0; o->fixed_work = 0; o->ilu_fill = 0; o->amg = 0;
  o->krylov = 0; o->restart = 100;
}

int ax1gpu_sizeof_params(void) { return (int)sizeof(ax1gpu_params); }

int ax1gpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int ax1gpu_create(const ax1gpu_grid* grid, const unsigned char* cell_subdomain, const double* memb_perm,
                  const double* node_volume, const unsigned char* dirichlet_mask, const ax1gpu_params* params,
                  const double* g_leak, const double* g_nav, const double* g_kv, int device, void* stream,
                  ax1gpu_handle* out) {
  if (!grid || !cell_subdomain || !memb_perm || !node_volume || !dirichlet_mask || !params || !out || !g_leak ||
      !g_nav || !g_kv) { set_error("null argument"); return AX1GPU_EINVAL; }
  if (ax1gpu_device_count() <= 0) { set_error("no CUDA device: libax1gpu has no CPU fallback"); return AX1GPU_ENODEVICE; }
  const int nx = grid->nx, ny = grid->ny;
  if (nx < 1 || ny < 3) { set_error("grid too small"); return AX1GPU_EINVAL; }
  if (!grid->x || !grid->y) { set_error("null argument"); return AX1GPU_EINVAL; }
  for (int i = 0; i < nx; ++i)
    if (!(grid->x[i + 1] > grid->x[i])) { set_error("x coordinates must be strictly increasing"); return AX1GPU_EINVAL; }
  for (int j = 0; j < ny; ++j)
    if (!(grid->y[j + 1] > grid->y[j])) { set_error("y coordinates must be strictly increasing"); return AX1GPU_EINVAL; }
  if (!(grid->y[0] >= 0.0)) { set_error("radial coordinates must be non-negative"); return AX1GPU_EINVAL; }
  // the membrane must be exactly one full cell row (nMembraneElements = 1, numberMembranes = 1)
  int jm = -1;
  for (int j = 0; j < ny; ++j) {
    const unsigned char sd = cell_subdomain[(size_t)j * nx];
    for (int i = 1; i < nx; ++i)
      if (cell_subdomain[(size_t)j * nx + i] != sd) { set_error("subdomain tags must be constant along x"); return AX1GPU_EINVAL; }
    if (sd == SD_MEMBRANE) {
      if (jm >= 0) { set_error("exactly one membrane cell layer is supported"); return AX1GPU_EINVAL; }
      jm = j;
    }
  }
  if (jm < 1 || jm > ny - 2) { set_error("membrane row missing or on the boundary"); return AX1GPU_EINVAL; }
  if ((long long)(nx + 1) * (ny + 1) > 500000000LL) { set_error("grid too large for int indexing"); return AX1GPU_EINVAL; }

  AX1_CUDA(cudaSetDevice(device));
  ax1gpu_ctx* c = new ax1gpu_ctx;
  c->device = device;
  if (stream) c->stream = (cudaStream_t)stream;
  else { AX1_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  c->prm = *params;
  c->left_pb = grid->left_proc_boundary; c->right_pb = grid->right_proc_boundary;
  Dev& g = c->g;
  g.nx = nx; g.ny = ny; g.nvx = nx + 1; g.nvy = ny + 1; g.nv = g.nvx * g.nvy; g.jm = jm;
  c->own_lo = c->left_pb ? 2 : 0;
  c->own_hi = c->right_pb ? g.nvx - 2 : g.nvx - 1;
  c->hx.assign(grid->x, grid->x + nx + 1);
  c->hy.assign(grid->y, grid->y + ny + 1);
  const size_t nv = g.nv, n4 = 4 * nv;
  int rc = 0;
  cudaStream_t st = c->stream;
#define TRY(x) if ((rc = (x))) { ax1gpu_destroy(c); return rc; }
  TRY(dev_upload(&c->d_x, grid->x, nx + 1, st));
  TRY(dev_upload(&c->d_y, grid->y, ny + 1, st));
  TRY(dev_upload(&c->d_nodevol, node_volume, nv, st));
  TRY(dev_upload(&c->d_dmask, (const uint8_t*)dirichlet_mask, nv, st));
  TRY(dev_upload(&c->d_epsm, memb_perm, nx, st));
  for (double** p : {&c->d_u, &c->d_uold, &c->d_r0, &c->d_r, &c->d_z, &c->d_uprev, &c->d_rt, &c->d_p, &c->d_v,
                     &c->d_y2, &c->d_t, &c->d_tmp, &c->d_y1}) {
    TRY(dev_alloc(p, n4));
    if (cudaMemsetAsync(*p, 0, n4 * sizeof(double), st) != cudaSuccess) { ax1gpu_destroy(c); return AX1GPU_ECUDA; }
  }
  TRY(dev_alloc(&c->d_A, nv * AX1_ROW));
  TRY(dev_alloc(&c->d_LU, nv * 144));
  c->lu_cap = nv * 144;
  // channels: gbar[3], state = ratio[2] + gate[3] + gate_new[3], geff[4]
  {
    std::vector<double> gb(3 * (size_t)nx);
    std::memcpy(&gb[0], g_leak, nx * sizeof(double));
    std::memcpy(&gb[nx], g_nav, nx * sizeof(double));
    std::memcpy(&gb[2 * (size_t)nx], g_kv, nx * sizeof(double));
    TRY(dev_upload(&c->d_gbar, gb.data(), gb.size(), st));
    TRY(dev_alloc(&c->d_chstate, 8 * (size_t)nx));
    TRY(dev_alloc(&c->d_geff, 4 * (size_t)nx));
    TRY(dev_alloc(&c->d_vrest, 1));
    TRY(dev_alloc(&c->d_vm, nx));
    TRY(dev_alloc(&c->d_err, 4));
    if (cudaMemsetAsync(c->d_err, 0, 4 * sizeof(int), st) != cudaSuccess) { ax1gpu_destroy(c); return AX1GPU_ECUDA; }
    for (int k = 0; k < 3; ++k) c->ch.gbar[k] = c->d_gbar + (size_t)k * nx;
    for (int k = 0; k < 2; ++k) c->ch.ratio[k] = c->d_chstate + (size_t)k * nx;
    for (int k = 0; k < 3; ++k) { c->ch.gate[k] = c->d_chstate + (size_t)(2 + k) * nx; c->ch.gate_new[k] = c->d_chstate + (size_t)(5 + k) * nx; }
    c->ch.geff = c->d_geff; c->ch.v_rest = c->d_vrest;
  }
  c->nred_blocks = 148 * 8;
  TRY(dev_alloc(&c->d_partials, 2 * (size_t)c->nred_blocks));
  // + 2 x 512 doubles behind the SpMV partials: second-stage partials of the two-stage reduction (driver.cu)
  TRY(dev_alloc(&c->d_spartials, 2 * (size_t)((4 * nv + 255) / 256) + 1024));
  TRY(dev_alloc(&c->d_sums, 8));
  TRY(dev_alloc(&c->d_sc, 1));
  if (cudaMallocHost((void**)&c->h_sc, sizeof(Scalars)) != cudaSuccess ||
      cudaMallocHost((void**)&c->h_pin, 8 * sizeof(double)) != cudaSuccess ||
      cudaMallocHost((void**)&c->h_err, 4 * sizeof(int)) != cudaSuccess) { ax1gpu_destroy(c); set_error("cudaMallocHost failed"); return AX1GPU_ECUDA; }
  std::memset(c->h_err, 0, 4 * sizeof(int));
  for (int k = 0; k < 4; ++k)
    if (cudaEventCreate(&c->ev[k]) != cudaSuccess) { ax1gpu_destroy(c); set_error("cudaEventCreate failed"); return AX1GPU_ECUDA; }
  // constants (electrolyte.hh:61-67, default_config.hh:79-91)
  g.x = c->d_x; g.y = c->d_y; g.node_vol = c->d_nodevol; g.dmask = c->d_dmask; g.eps_memb = c->d_epsm; g.geff = c->d_geff;
  g.eps_elec = params->eps_elec;
  g.PC = con_e * con_e * con_mol * params->length_scale * params->length_scale / (con_eps0 * con_k * params->temperature);
  g.kT_e = (con_k * params->temperature) / con_e;
  for (int j = 0; j < 3; ++j) {
    g.D[j] = params->diff_water[j] * params->time_scale / (params->length_scale * params->length_scale);
    g.valence[j] = params->valence[j];
  }
  g.S_pot = params->pot_residual_scale; g.temperature = params->temperature;
  g.time_scale = params->time_scale; g.length_scale = params->length_scale;
  g.do_volume_scaling = params->do_volume_scaling; g.membrane_active = params->membrane_active;
  g.stim_x = params->stim_x; g.stim_y = params->stim_y; g.I_inj = params->I_inj; g.dummy_value = params->dummy_value;
  // pattern (FullVolumePattern, columns ascending)
  c->h_rowptr.assign(nv + 1, 0);
  c->h_col.clear();
  c->h_col.reserve(9 * nv);
  for (int j = 0; j < g.nvy; ++j)
    for (int i = 0; i < g.nvx; ++i) {
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          const int ii = i + di, jj = j + dj;
          if (ii < 0 || ii >= g.nvx || jj < 0 || jj >= g.nvy) continue;
          c->h_col.push_back(ii + g.nvx * jj);
        }
      c->h_rowptr[(size_t)i + (size_t)g.nvx * j + 1] = (int)c->h_col.size();
    }
  c->nblocks = (int)c->h_col.size();
  TRY(dev_upload(&c->d_rowptr, c->h_rowptr.data(), nv + 1, st));
  // static channel init (DynamicChannelSet::init + VoltageGatedChannel::init at v = v_rest)
  {
    std::vector<double> s(5 * (size_t)nx);
    auto vtrap = [](double x, double y) { return std::fabs(x / y) < 1e-8 ? y * (1 - x / y / 2) : x / (std::exp(x / y) - 1); };
    const double am = 0.1 * vtrap(25., 10.), bm = 4.0;
    const double ah = 0.07, bh = 1. / (std::exp(30. / 10) + 1);
    const double an = 0.01 * vtrap(10., 10.), bn = 0.125;
    for (int i = 0; i < nx; ++i) {
      s[i] = params->leak_ratio_na; s[nx + i] = params->leak_ratio_k;
      s[2 * (size_t)nx + i] = am / (am + bm); s[3 * (size_t)nx + i] = ah / (ah + bh); s[4 * (size_t)nx + i] = an / (an + bn);
    }
    if ((rc = ax1gpu_set_channel_state(c, s.data(), -65.0, 0))) { ax1gpu_destroy(c); return rc; }
  }
#undef TRY
  AX1_CUDA(cudaStreamSynchronize(st));
  *out = c;
  return AX1GPU_OK;
}

int ax1gpu_destroy(ax1gpu_handle c) {
  if (!c) return AX1GPU_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  drv_amg_free(c);
  if (c->d_gmV) cudaFree(c->d_gmV);
  if (c->d_gm) cudaFree(c->d_gm);
  if (c->p2p.win) {
    for (int k = 0; k < c->nranks && k < AX1_MAX_RANKS; ++k)
      if (k != c->rank && c->p2p.peer[k]) cudaIpcCloseMemHandle(c->p2p.peer[k]);
    cudaFree(c->p2p.win);
    if (c->p2p.cnt) cudaFree(c->p2p.cnt);
  }
  if (c->comm && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comm);
  for (void* p : {(void*)c->d_x, (void*)c->d_y, (void*)c->d_nodevol, (void*)c->d_epsm, (void*)c->d_dmask, (void*)c->d_rowptr,
                  (void*)c->d_u, (void*)c->d_uold, (void*)c->d_r0, (void*)c->d_r, (void*)c->d_z, (void*)c->d_uprev, (void*)c->d_A,
                  (void*)c->d_LU, (void*)c->d_rt, (void*)c->d_p, (void*)c->d_v, (void*)c->d_y2, (void*)c->d_t, (void*)c->d_tmp, (void*)c->d_y1,
                  (void*)c->d_bcrs, (void*)c->d_gbar, (void*)c->d_chstate, (void*)c->d_geff, (void*)c->d_vrest, (void*)c->d_vm, (void*)c->d_fluxterms,
                  (void*)c->d_mori_gamma, (void*)c->d_Ac, (void*)c->d_err, (void*)c->d_partials, (void*)c->d_sums, (void*)c->d_sc, (void*)c->d_halo_send,
                  (void*)c->d_halo_recv, (void*)c->d_lvl, (void*)c->d_spartials})
    if (p) cudaFree(p);
  if (c->h_sc) cudaFreeHost(c->h_sc);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_err) cudaFreeHost(c->h_err);
  for (int k = 0; k < 4; ++k) if (c->ev[k]) cudaEventDestroy(c->ev[k]);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return AX1GPU_OK;
}

int ax1gpu_pattern(ax1gpu_handle c, int* rowptr, int* colidx, int* nblocks) {
  AX1_CHECK_H(c);
  if (nblocks) *nblocks = c->nblocks;
  if (rowptr) std::memcpy(rowptr, c->h_rowptr.data(), c->h_rowptr.size() * sizeof(int));
  if (colidx) std::memcpy(colidx, c->h_col.data(), c->h_col.size() * sizeof(int));
  return AX1GPU_OK;
}

int ax1gpu_set_time(ax1gpu_handle c, double t, double dt) {
  AX1_CHECK_H(c);
  c->time = t; c->dt = dt;
  return drv_refresh_dev(c);
}

int ax1gpu_set_uold(ax1gpu_handle c, const double* uold) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = up(c, c->d_uold, uold);
  if (rc) return rc;
  return drv_set_uold(c);
}

int ax1gpu_set_uold_dev(ax1gpu_handle c) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  AX1_CUDA(cudaMemcpyAsync(c->d_uold, c->d_u, 4 * (size_t)c->g.nv * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return drv_set_uold(c);
}

int ax1gpu_upload_u(ax1gpu_handle c, const double* u) { AX1_CHECK_H(c); cudaSetDevice(c->device); return up(c, c->d_u, u); }
int ax1gpu_download_u(ax1gpu_handle c, double* u) { AX1_CHECK_H(c); cudaSetDevice(c->device); return down(c, u, c->d_u); }

int ax1gpu_update_state(ax1gpu_handle c, const double* u) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (u) { int rc = up(c, c->d_u, u); if (rc) return rc; }
  int rc = drv_update_state(c, c->d_u);
  if (rc) return rc;
  int herr = 0;
  AX1_CUDA(cudaMemcpyAsync(&herr, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  if (herr) { set_error("Calculated leak channel ratio outside [0,1]"); return AX1GPU_ESTATE; }
  return AX1GPU_OK;
}

int ax1gpu_accept_state(ax1gpu_handle c) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = drv_accept_state(c);
  if (rc) return rc;
  return drv_mori_accept_state(c);     // gfMoriFlux.acceptTimeStep(); no-op unless the Mori model is in use
}

int ax1gpu_mori_step(ax1gpu_handle c, const ax1gpu_newton_opts* opts, double* u, ax1gpu_mori_result* res) {
  AX1_CHECK_H(c);
  if (!res) { set_error("mori_step: res must not be NULL"); return AX1GPU_EINVAL; }
  cudaSetDevice(c->device);
  ax1gpu_newton_opts o;
  if (opts) o = *opts; else ax1gpu_newton_opts_default(&o);
  int rc;
  if (u && (rc = up(c, c->d_u, u))) return rc;
  if ((rc = drv_mori_step(c, &o, res))) return rc;
  if (u) return down(c, u, c->d_u);
  return AX1GPU_OK;
}

int ax1gpu_mori_assemble(ax1gpu_handle c, int which, const double* u, const double* ulast, double* r, double* blocks) {
  AX1_CHECK_H(c);
  if (!u || !ulast) { set_error("mori_assemble: u and ulast must not be NULL"); return AX1GPU_EINVAL; }
  cudaSetDevice(c->device);
  int rc;
  if ((rc = up(c, c->d_u, u))) return rc;
  if ((rc = up(c, c->d_uprev, ulast))) return rc;
  if ((rc = drv_mori_assemble(c, which ? 1 : 0, c->d_u, c->d_uprev))) return rc;
  if (r && (rc = down(c, r, c->d_r))) return rc;
  if (blocks) return ax1gpu_get_jacobian(c, blocks);
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_mori_update_state(ax1gpu_handle c, const double* u) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc;
  if (u && (rc = up(c, c->d_u, u))) return rc;
  return drv_mori_update_state(c, c->d_u);
}

int ax1gpu_mori_get_gamma(ax1gpu_handle c, double* gamma) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  const size_t n = 12 * (size_t)c->g.nx;
  if (!c->d_mori_gamma) { std::memset(gamma, 0, n * sizeof(double)); return AX1GPU_OK; }
  AX1_CUDA(cudaMemcpyAsync(gamma, c->d_mori_gamma, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_get_channel_state(ax1gpu_handle c, double* s, int newstate, double* v_rest) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  const size_t nx = c->g.nx;
  AX1_CUDA(cudaMemcpyAsync(s, c->d_chstate, 2 * nx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaMemcpyAsync(s + 2 * nx, c->d_chstate + (newstate ? 5 : 2) * nx, 3 * nx * sizeof(double),
                           cudaMemcpyDeviceToHost, c->stream));
  if (v_rest) AX1_CUDA(cudaMemcpyAsync(v_rest, c->d_vrest, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_set_channel_state(ax1gpu_handle c, const double* s, double v_rest, int initialized) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  const size_t nx = c->g.nx;
  AX1_CUDA(cudaMemcpyAsync(c->d_chstate, s, 5 * nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  AX1_CUDA(cudaMemcpyAsync(c->d_chstate + 5 * nx, s + 2 * nx, 3 * nx * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  AX1_CUDA(cudaMemcpyAsync(c->d_vrest, &v_rest, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  c->channels_initialized = initialized != 0;
  c->channels_partially_initialized = false;
  launch_channels(c->g, c->ch, c->d_u, 2, 0.0, 0, 0.0, c->d_err, c->stream);  // refresh geff
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  return AX1GPU_OK;
}

int ax1gpu_membrane_potential(ax1gpu_handle c, const double* u, double* vm) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (u) { int rc = up(c, c->d_u, u); if (rc) return rc; }
  launch_memb_pot(c->g, c->d_u, c->d_vm, c->stream);
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  AX1_CUDA(cudaMemcpyAsync(vm, c->d_vm, c->g.nx * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_membrane_flux_terms(ax1gpu_handle c, const double* u, double* terms) {
  AX1_CHECK_H(c);
  if (!terms) { set_error("ax1gpu_membrane_flux_terms: terms is null"); return AX1GPU_EINVAL; }
  cudaSetDevice(c->device);
  if (u) { int rc = up(c, c->d_u, u); if (rc) return rc; }
  const size_t n = 15 * (size_t)c->g.nx;
  if (!c->d_fluxterms) { int rc = dev_alloc(&c->d_fluxterms, n); if (rc) return rc; }
  launch_memb_flux_terms(c->g, c->d_u, c->d_fluxterms, c->stream);
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  AX1_CUDA(cudaMemcpyAsync(terms, c->d_fluxterms, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_residual(ax1gpu_handle c, const double* u, double* r, double* norm) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (u) { int rc = up(c, c->d_u, u); if (rc) return rc; }
  int rc = drv_residual(c, c->d_u, c->d_r, norm);
  if (rc) return rc;
  if (r) return down(c, r, c->d_r);
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

static int ensure_bcrs(ax1gpu_ctx* c) {
  if (!c->d_bcrs) return dev_alloc(&c->d_bcrs, (size_t)c->nblocks * 16);
  return 0;
}

int ax1gpu_get_jacobian(ax1gpu_handle c, double* blocks) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = ensure_bcrs(c);
  if (rc) return rc;
  launch_compact_bcrs(c->g, c->d_A, c->d_bcrs, c->d_rowptr, c->stream);
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  AX1_CUDA(cudaMemcpyAsync(blocks, c->d_bcrs, (size_t)c->nblocks * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_set_jacobian(ax1gpu_handle c, const double* blocks) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = ensure_bcrs(c);
  if (rc) return rc;
  AX1_CUDA(cudaMemcpyAsync(c->d_bcrs, blocks, (size_t)c->nblocks * 16 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  AX1_CUDA(cudaMemsetAsync(c->d_err + 1, 0, sizeof(int), c->stream));
  launch_expand_bcrs(c->g, c->d_bcrs, c->d_A, c->d_rowptr, c->d_err + 1, c->stream);
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  int viol = 0;
  AX1_CUDA(cudaMemcpyAsync(&viol, c->d_err + 1, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  c->factored = false;
  if (viol) {
    set_error("ax1gpu_set_jacobian: non-zero entry outside the PNP block shape (con_i couples to con_i and pot only)");
    return AX1GPU_EINVAL;
  }
  return AX1GPU_OK;
}

int ax1gpu_jacobian(ax1gpu_handle c, const double* u, double* blocks) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (u) { int rc = up(c, c->d_u, u); if (rc) return rc; }
  int rc = drv_jacobian(c, c->d_u);
  if (rc) return rc;
  if (blocks) return ax1gpu_get_jacobian(c, blocks);
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_rownorm(ax1gpu_handle c, double* r) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (r) { int rc = up(c, c->d_r, r); if (rc) return rc; }
  launch_rownorm(c->g, c->d_A, c->d_r, c->stream);
  c->launches++;
  AX1_CUDA(cudaGetLastError());
  c->factored = false;
  if (r) return down(c, r, c->d_r);
  return AX1GPU_OK;
}

int ax1gpu_spmv(ax1gpu_handle c, const double* x, double* y) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = up(c, c->d_y2, x);
  if (rc) return rc;
  if ((rc = drv_op_apply(c, c->d_y2, c->d_v))) return rc;
  return down(c, y, c->d_v);
}

int ax1gpu_ilu_factor(ax1gpu_handle c, int slab_w) { return ax1gpu_ilu_factor_n(c, slab_w, -1); }

int ax1gpu_ilu_factor_n(ax1gpu_handle c, int slab_w, int fill) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = drv_factor(c, slab_w, fill);
  if (rc) return rc;
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  return AX1GPU_OK;
}

int ax1gpu_set_amg(ax1gpu_handle c, int on, int cx) {
  AX1_CHECK_H(c);
  if (cx > 0 && cx != c->amg_cx) {
    if (cx < 2 || cx > 64) { set_error("amg: columns per aggregate must be in [2, 64]"); return AX1GPU_EINVAL; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drv_amg_free(c);
    c->amg_cx = cx;
  }
  c->amg = on ? 1 : 0;
  c->factored = false;
  return AX1GPU_OK;
}

int ax1gpu_set_ilu_split(ax1gpu_handle c, int jcut) {
  AX1_CHECK_H(c);
  if (jcut > 0 && (jcut < 2 || jcut > c->g.nvy - 2)) { set_error("ilu split: row out of range"); return AX1GPU_EINVAL; }
  c->ilu_split = jcut < 0 ? -1 : jcut;
  c->factored = false;
  return AX1GPU_OK;
}

int ax1gpu_set_spmv_compact(ax1gpu_handle c, int mode) {
  AX1_CHECK_H(c);
  c->spmv_compact = mode < 0 ? -1 : (mode ? 1 : 0);
  c->factored = false;
  return AX1GPU_OK;
}

int ax1gpu_get_spmv_compact(ax1gpu_handle c, int* active) {
  AX1_CHECK_H(c);
  if (active) *active = c->factored && c->ac_valid;
  return AX1GPU_OK;
}

int ax1gpu_get_ilu_split(ax1gpu_handle c, int* active) {
  AX1_CHECK_H(c);
  if (active) *active = c->ilu_jc;
  return AX1GPU_OK;
}

int ax1gpu_ilu_apply(ax1gpu_handle c, const double* d, double* v) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (!c->factored) { set_error("ax1gpu_ilu_apply before ax1gpu_ilu_factor"); return AX1GPU_ESTATE; }
  int rc = up(c, c->d_p, d);
  if (rc) return rc;
  if ((rc = drv_prec_apply(c, c->d_p, c->d_y2))) return rc;
  if (c->nranks > 1 && c->p2p.on)   // the halo exchange may have timed out: never hand stale columns back
    AX1_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, 3 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  if ((rc = down(c, v, c->d_y2))) return rc;
  return drv_check_p2p(c);
}

int ax1gpu_solve(ax1gpu_handle c, double* r, double reduction, int maxit, int slab_w, double* z, ax1gpu_lin_result* res) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc;
  if (r) { if ((rc = up(c, c->d_r, r))) return rc; }
  if (!c->factored || (slab_w > 0 && slab_w != c->slab_w)) { if ((rc = drv_factor(c, slab_w))) return rc; }
  ax1gpu_lin_result lr{};
  if ((rc = drv_solve(c, c->d_r, c->d_z, reduction, maxit, &lr))) return rc;
  if (res) *res = lr;
  if (z) { if ((rc = down(c, z, c->d_z))) return rc; }
  if (r) { if ((rc = down(c, r, c->d_r))) return rc; }
  return AX1GPU_OK;
}

int ax1gpu_solve_gmres(ax1gpu_handle c, double* r, double reduction, int restart, int maxit, int slab_w, double* z,
                       ax1gpu_lin_result* res) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc;
  if (r) { if ((rc = up(c, c->d_r, r))) return rc; }
  if (!c->factored || (slab_w > 0 && slab_w != c->slab_w)) { if ((rc = drv_factor(c, slab_w))) return rc; }
  ax1gpu_lin_result lr{};
  if ((rc = drv_solve_gmres(c, c->d_r, c->d_z, reduction, restart, maxit, &lr))) return rc;
  if (res) *res = lr;
  if (z) { if ((rc = down(c, z, c->d_z))) return rc; }
  if (r) { if ((rc = down(c, r, c->d_r))) return rc; }
  return AX1GPU_OK;
}

int ax1gpu_newton_dev(ax1gpu_handle c, const ax1gpu_newton_opts* opts, ax1gpu_newton_result* res) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  ax1gpu_newton_opts o;
  if (opts) o = *opts; else ax1gpu_newton_opts_default(&o);
  ax1gpu_newton_result r{};
  int rc = drv_newton(c, &o, &r);
  if (res) *res = r;
  return rc;
}

int ax1gpu_newton(ax1gpu_handle c, const ax1gpu_newton_opts* opts, double* u, ax1gpu_newton_result* res) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  int rc = up(c, c->d_u, u);
  if (rc) return rc;
  if ((rc = ax1gpu_newton_dev(c, opts, res))) return rc;
  return down(c, u, c->d_u);
}

void ax1gpu_timeloop_opts_default(ax1gpu_timeloop_opts* o) {
  // [timeStepping] defaults of acme2_cyl_parametertree.hh (seconds) over the default time_scale 1e-6
  o->dtstart = 1.0; o->max_dt = 50.0; o->min_dt = 0.05; o->max_dt_ap = 10.0; o->vm_ap_threshold = -50.0;
  o->lower_limit = 3; o->upper_limit = 5; o->max_restarts = 3; o->adaptive = 1; o->restart_from_uold = 0;
}

int ax1gpu_time_step(ax1gpu_handle c, const ax1gpu_timeloop_opts* opts, const ax1gpu_newton_opts* newton,
                     ax1gpu_timeloop_state* state, ax1gpu_step_result* res) {
  AX1_CHECK_H(c);
  if (!state || !res) { set_error("time_step: state and res must not be NULL"); return AX1GPU_EINVAL; }
  cudaSetDevice(c->device);
  ax1gpu_timeloop_opts o;
  if (opts) o = *opts; else ax1gpu_timeloop_opts_default(&o);
  ax1gpu_newton_opts no;
  if (newton) no = *newton; else ax1gpu_newton_opts_default(&no);
  return drv_time_step(c, &o, &no, state, res);
}

int ax1gpu_membrane_potential_range(ax1gpu_handle c, double* vm_min, double* vm_max) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  double lo = 0, hi = 0;
  int rc = drv_vm_minmax(c, &lo, &hi);
  if (rc) return rc;
  if (vm_min) *vm_min = lo;
  if (vm_max) *vm_max = hi;
  return AX1GPU_OK;
}

int ax1gpu_membrane_potential_dev(ax1gpu_handle c, double* vm) { return ax1gpu_membrane_potential(c, nullptr, vm); }

int ax1gpu_time_kernel(ax1gpu_handle c, const char* name, int reps, double* ms) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  return drv_time_kernel(c, name, reps, ms);
}

long long ax1gpu_launch_count(ax1gpu_handle c) { return c ? c->launches : 0; }
void* ax1gpu_stream(ax1gpu_handle c) { return c ? (void*)c->stream : nullptr; }

int ax1gpu_nccl_unique_id(const char* libnccl_path, char* id128) {
  NcclApi api;
  if (!api.load(libnccl_path)) { set_error(std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "?")); return AX1GPU_ENCCL; }
  ncclUniqueId id;
  if (api.GetUniqueId(&id) != ncclSuccess) { set_error("ncclGetUniqueId failed"); return AX1GPU_ENCCL; }
  std::memcpy(id128, id.internal, 128);
  return AX1GPU_OK;
}

}  // extern "C"

// Peer windows: every rank exports one small cudaMalloc'ed window with CUDA IPC; the handles are gathered
// with the NCCL communicator that already exists, and every rank maps all peers' windows. If any rank
// cannot map a peer (no NVLink/P2P, AX1_P2P=0), all ranks fall back to the NCCL send/recv + allreduce path.
static int p2p_setup(ax1gpu_ctx* c) {
  const char* env = getenv("AX1_P2P");
  int want = !(env && env[0] == '0') && c->nranks <= AX1_MAX_RANKS;
  auto& P = c->p2p;
  P.halo_len = 3 * (size_t)c->g.nvy * 4;
  const size_t wlen = AX1_P2P_HALO_OFF + 4 * P.halo_len;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  int ok = want;
  if (ok && cudaMalloc((void**)&P.win, wlen * sizeof(double)) != cudaSuccess) { cudaGetLastError(); ok = 0; P.win = nullptr; }
  if (ok && cudaMalloc((void**)&P.cnt, 2 * sizeof(unsigned int)) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  if (ok) {
    cudaMemsetAsync(P.win, 0, wlen * sizeof(double), c->stream);
    cudaMemsetAsync(P.cnt, 0, 2 * sizeof(unsigned int), c->stream);
    if (cudaIpcGetMemHandle(&mine, P.win) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // gather the handles (64 B each) through NCCL
  char *d_in = nullptr, *d_all = nullptr;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  AX1_CUDA(cudaMalloc((void**)&d_in, hb));
  AX1_CUDA(cudaMalloc((void**)&d_all, hb * c->nranks));
  AX1_CUDA(cudaMemcpyAsync(d_in, &mine, hb, cudaMemcpyHostToDevice, c->stream));
  if (c->nccl.AllGather(d_in, d_all, hb, ncclInt8, c->comm, c->stream) != ncclSuccess) { set_error("ncclAllGather failed"); return AX1GPU_ENCCL; }
  std::vector<cudaIpcMemHandle_t> all(c->nranks);
  AX1_CUDA(cudaMemcpyAsync(all.data(), d_all, hb * c->nranks, cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  if (ok) {
    for (int k = 0; k < c->nranks; ++k) {
      if (k == c->rank) { P.peer[k] = P.win; continue; }
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, all[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      P.peer[k] = (double*)ptr;
    }
  }
  // agree: peer windows are used only if every rank mapped every peer (min over ranks); the all-reduce also
  // guarantees that every rank's window has been zeroed before anybody writes into it
  int* d_ok = (int*)d_in;
  AX1_CUDA(cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (c->nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->comm, c->stream) != ncclSuccess) { set_error("ncclAllReduce failed"); return AX1GPU_ENCCL; }
  AX1_CUDA(cudaMemcpyAsync(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_in); cudaFree(d_all);
  P.on = ok != 0;
  {
    double sec = AX1_P2P_TIMEOUT_DEFAULT_S;
    if (const char* t = getenv("AX1_P2P_TIMEOUT_S")) { const double v = atof(t); if (v > 0.0) sec = v; }
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device) != cudaSuccess || khz <= 0) { cudaGetLastError(); khz = 1965000; }
    P.timeout_cycles = (long long)(sec * 1e3 * (double)khz);
  }
  return AX1GPU_OK;
}

extern "C" {

int ax1gpu_comm_init(ax1gpu_handle c, const char* libnccl_path, const char* id128, int rank, int nranks) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (nranks <= 1) { c->rank = 0; c->nranks = 1; return AX1GPU_OK; }
  if (!c->nccl.load(libnccl_path)) { set_error("cannot load NCCL"); return AX1GPU_ENCCL; }
  if ((rank > 0) != c->left_pb || (rank < nranks - 1) != c->right_pb) {
    set_error("processor-boundary flags of the grid do not match rank/nranks"); return AX1GPU_EINVAL;
  }
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclResult_t r = c->nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { set_error(std::string("ncclCommInitRank: ") + c->nccl.GetErrorString(r)); return AX1GPU_ENCCL; }
  c->rank = rank; c->nranks = nranks;
  const size_t n = 2 * 3 * (size_t)c->g.nvy * 4;
  int rc;
  if ((rc = dev_alloc(&c->d_halo_send, n))) return rc;
  if ((rc = dev_alloc(&c->d_halo_recv, n))) return rc;
  return p2p_setup(c);
}

// ---- peer windows without NCCL: the caller exchanges the CUDA IPC handles itself (MPI_Allgather in a DUNE build,
// torch.distributed / gloo in the test harness). Two calls because the exchange sits between them.
int ax1gpu_p2p_export(ax1gpu_handle c, int rank, int nranks, char* handle64) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  if (nranks < 2 || nranks > AX1_MAX_RANKS || rank < 0 || rank >= nranks) { set_error("p2p_export: bad rank / nranks"); return AX1GPU_EINVAL; }
  if ((rank > 0) != c->left_pb || (rank < nranks - 1) != c->right_pb) {
    set_error("processor-boundary flags of the grid do not match rank/nranks"); return AX1GPU_EINVAL;
  }
  if (c->p2p.win) { set_error("p2p_export: communication already initialised"); return AX1GPU_ESTATE; }
  auto& P = c->p2p;
  P.halo_len = 3 * (size_t)c->g.nvy * 4;
  const size_t wlen = AX1_P2P_HALO_OFF + 4 * P.halo_len;
  AX1_CUDA(cudaMalloc((void**)&P.win, wlen * sizeof(double)));
  AX1_CUDA(cudaMalloc((void**)&P.cnt, 2 * sizeof(unsigned int)));
  AX1_CUDA(cudaMemsetAsync(P.win, 0, wlen * sizeof(double), c->stream));
  AX1_CUDA(cudaMemsetAsync(P.cnt, 0, 2 * sizeof(unsigned int), c->stream));
  AX1_CUDA(cudaStreamSynchronize(c->stream));      // zeroed before any peer can learn the handle
  cudaIpcMemHandle_t mine;
  AX1_CUDA(cudaIpcGetMemHandle(&mine, P.win));
  std::memcpy(handle64, &mine, sizeof(mine));
  c->rank = rank; c->nranks = nranks;
  return AX1GPU_OK;
}

int ax1gpu_p2p_attach(ax1gpu_handle c, const char* all_handles) {
  AX1_CHECK_H(c);
  cudaSetDevice(c->device);
  auto& P = c->p2p;
  if (!P.win || c->nranks < 2) { set_error("p2p_attach before p2p_export"); return AX1GPU_ESTATE; }
  for (int k = 0; k < c->nranks; ++k) {
    if (k == c->rank) { P.peer[k] = P.win; continue; }
    cudaIpcMemHandle_t hk;
    std::memcpy(&hk, all_handles + (size_t)k * sizeof(hk), sizeof(hk));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, hk, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); return AX1GPU_ECUDA; }
    P.peer[k] = (double*)ptr;
  }
  double sec = AX1_P2P_TIMEOUT_DEFAULT_S;
  if (const char* t = getenv("AX1_P2P_TIMEOUT_S")) { const double v = atof(t); if (v > 0.0) sec = v; }
  int khz = 0;
  if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device) != cudaSuccess || khz <= 0) { cudaGetLastError(); khz = 1965000; }
  P.timeout_cycles = (long long)(sec * 1e3 * (double)khz);
  P.on = true;
  return AX1GPU_OK;
}

int ax1gpu_comm_mode(ax1gpu_handle c) { return c ? (c->nranks > 1 ? (c->p2p.on ? 2 : 1) : 0) : 0; }

}  // extern "C"
