// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY. C API of the CPU oracle (see ax1_oracle.h).
#include "oracle.hpp"
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>

using namespace ax1o;

struct ax1o_problem { Problem p; };

static thread_local std::string g_err;
#define AX1O_TRY try {
#define AX1O_CATCH(ret) } catch (std::exception& e) { g_err = e.what(); return ret; }

extern "C" {

const char* ax1o_last_error(void) { return g_err.c_str(); }

void ax1o_config_default(ax1o_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->temperature = 279.45; c->eps_elec = 80.0; c->length_scale = 1e-9; c->time_scale = 1e-6;
  c->diff_water[0] = 1.33e-9; c->diff_water[1] = 1.96e-9; c->diff_water[2] = 2.03e-9;
  c->valence[0] = 1; c->valence[1] = 1; c->valence[2] = -1;
  c->pot_residual_scale = 1.0;
  c->ymemb = 500.0; c->dmemb = 5.0;
  c->do_volume_scaling = 1; c->volume_scaling_threshold = -1.0;
  c->top_dirichlet[0] = c->top_dirichlet[1] = 1;
  c->dummy_value = 1.0;
  c->membrane_active = 1; c->automatic_channel_init = 1; c->channel_resting_potential = -65.0;
  c->analytic_volume_jacobian = 7;
  c->leak_ratio_na = 0.13; c->leak_ratio_k = 0.87;
  c->do_memb_flux_stimulation = 0; c->memb_flux_ion = 2 /* Cl */; c->memb_flux_inj = 0.0;
  c->mori_membrane_neumann = 1;
}

void ax1o_newton_opts_default(ax1o_newton_opts* o) {
  o->reduction = 1e-5; o->abs_limit = 1e-5; o->min_lin_reduction = 1e-5;
  o->fixed_lin_reduction = 0; o->maxit = 50; o->line_search_maxit = 10; o->lin_maxit = 5000;
  o->force_iteration = 0; o->disable_line_search = 0;
  o->ilu_fill = 1; o->slab_w = 0; o->fixed_work = 0;
  o->krylov = 0; o->restart = 100; o->amg = 0; o->amg_cx = 4;
}

int ax1o_sizeof_config(void) { return (int)sizeof(ax1o_config); }

void ax1o_set_ilu_split(ax1o_problem* h, int jcut) { h->p.ilu_jcut = jcut > 0 ? jcut : 0; }

int ax1o_generate_y(double ymin, double ymax, double ymemb, double dmemb, double dy_cell,
                    double dy_cell_min, int n_dy_cell_min, double dy, double dy_min, int n_dy_min,
                    double* out, int cap) {
  AX1O_TRY
  std::vector<double> y = generate_y(ymin, ymax, ymemb, dmemb, dy_cell, dy_cell_min, n_dy_cell_min, dy,
                                     dy_min, n_dy_min);
  if (out) for (int i = 0; i < (int)y.size() && i < cap; ++i) out[i] = y[i];
  return (int)y.size();
  AX1O_CATCH(-1)
}

double ax1o_line_search(int n, double* u, const double* z, double cur_defect, double prev_defect, int maxit,
                        ax1o_defect_cb cb, void* ctx, double* trace, int cap, int* ntrace) {
  std::vector<double> prev_u(n), tr;
  double d = ax1o::line_search((size_t)n, u, z, prev_u.data(), cur_defect, prev_defect, maxit,
                               [&]() { return cb(ctx, u, n); }, &tr);
  *ntrace = (int)tr.size();
  for (int i = 0; i < (int)tr.size() && i < cap; ++i) trace[i] = tr[i];
  return d;
}

int ax1o_gridvector(int op, double start, int n, double a, double b, double c, double* out, int cap) {
  AX1O_TRY
  std::vector<double> g = gridvector_op(op, start, n, a, b, c);
  for (int i = 0; i < (int)g.size() && i < cap; ++i) out[i] = g[i];
  return (int)g.size();
  AX1O_CATCH(-1)
}

ax1o_problem* ax1o_create(const ax1o_config* cfg, int nx, int ny, const double* x, const double* y,
                          int left_pb, int right_pb, const double* eps_memb, const double* g_leak,
                          const double* g_nav, const double* g_kv) {
  AX1O_TRY
  ax1o_problem* h = new ax1o_problem;
  Problem& p = h->p;
  p.cfg = Config(*cfg);
  p.nx = nx; p.ny = ny;
  p.x.assign(x, x + nx + 1); p.y.assign(y, y + ny + 1);
  p.left_proc_boundary = left_pb; p.right_proc_boundary = right_pb;
  if (eps_memb) p.eps_memb.assign(eps_memb, eps_memb + nx);
  if (g_leak) p.gbar[0].assign(g_leak, g_leak + nx);
  if (g_nav) p.gbar[1].assign(g_nav, g_nav + nx);
  if (g_kv) p.gbar[2].assign(g_kv, g_kv + nx);
  try { setup_problem(p); } catch (...) { delete h; throw; }
  p.r0.assign(4 * (size_t)p.nv, 0.0);
  return h;
  AX1O_CATCH(nullptr)
}

void ax1o_destroy(ax1o_problem* p) { delete p; }
int ax1o_num_vertices(const ax1o_problem* h) { return h->p.nv; }
int ax1o_membrane_row(const ax1o_problem* h) { return h->p.jm; }

void ax1o_get_maps(const ax1o_problem* h, unsigned char* cell_sub, double* node_volume, unsigned char* dirichlet) {
  const Problem& p = h->p;
  if (cell_sub) std::memcpy(cell_sub, p.cell_sub.data(), p.cell_sub.size());
  if (node_volume) std::memcpy(node_volume, p.node_volume.data(), p.nv * sizeof(double));
  if (dirichlet) std::memcpy(dirichlet, p.dirichlet.data(), p.nv);
}

int ax1o_pattern(const ax1o_problem* h, int* rowptr, int* col) {
  BCRS A;
  build_pattern(h->p, A);
  if (rowptr) std::memcpy(rowptr, A.rowptr.data(), A.rowptr.size() * sizeof(int));
  if (col) std::memcpy(col, A.col.data(), A.col.size() * sizeof(int));
  return (int)A.col.size();
}

void ax1o_constants(const ax1o_problem* h, double* out) {
  out[0] = h->p.PC; out[1] = h->p.kT_e; out[2] = h->p.D[0]; out[3] = h->p.D[1]; out[4] = h->p.D[2];
}

void ax1o_set_time(ax1o_problem* h, double t, double dt) { h->p.time = t; h->p.dt = dt; }
void ax1o_set_uold(ax1o_problem* h, const double* uold) { set_uold(h->p, uold); }
int ax1o_update_state(ax1o_problem* h, const double* u) {
  AX1O_TRY update_state(h->p, u); return 0; AX1O_CATCH(1)
}
void ax1o_accept_state(ax1o_problem* h) {
  accept_state(h->p);
  if (!h->p.mori_gamma_int_old[0].empty()) mori_accept_state(h->p);   // gfMoriFlux.acceptTimeStep()
}

int ax1o_mori_assemble(ax1o_problem* h, int which, const double* u, const double* ulast, double* r, double* vals) {
  AX1O_TRY
  BCRS A;
  if (vals) build_pattern(h->p, A);
  if (which == 0) mori_potential_assemble(h->p, u, ulast, r, vals ? &A : nullptr);
  else mori_concentration_assemble(h->p, u, ulast, r, vals ? &A : nullptr);
  if (vals) std::memcpy(vals, A.val.data(), A.val.size() * sizeof(double));
  return 0;
  AX1O_CATCH(1)
}

void ax1o_mori_get_gamma(const ax1o_problem* h, double* out) {
  const Problem& p = h->p;
  const int nx = p.nx;
  if (p.mori_gamma_int_old[0].empty()) { for (int k = 0; k < 12 * nx; ++k) out[k] = 0.0; return; }
  for (int j = 0; j < 3; ++j)
    for (int i = 0; i < nx; ++i) {
      out[((0 * 2 + 0) * 3 + j) * nx + i] = p.mori_gamma_int_old[j][i];
      out[((0 * 2 + 1) * 3 + j) * nx + i] = p.mori_gamma_int_new[j][i];
      out[((1 * 2 + 0) * 3 + j) * nx + i] = p.mori_gamma_ext_old[j][i];
      out[((1 * 2 + 1) * 3 + j) * nx + i] = p.mori_gamma_ext_new[j][i];
    }
}

void ax1o_mori_update_state(ax1o_problem* h, const double* u) { mori_update_state(h->p, u); }

int ax1o_mori_step(ax1o_problem* h, double* u, const ax1o_newton_opts* o, ax1o_mori_result* res) {
  AX1O_TRY
  NewtonOpts no;
  static_cast<ax1o_newton_opts&>(no) = *o;
  MoriResult mr;
  mori_step(h->p, u, no, mr);
  *res = mr;
  return 0;
  AX1O_CATCH(-1)
}

void ax1o_get_channel_state(const ax1o_problem* h, double* s, int newstate) {
  const Problem& p = h->p;
  const int n = p.nx;
  for (int i = 0; i < n; ++i) {
    s[i] = p.ratio[0][i]; s[n + i] = p.ratio[1][i];
    for (int g = 0; g < 3; ++g) s[(2 + g) * n + i] = newstate ? p.gate_new[g][i] : p.gate[g][i];
  }
}

void ax1o_set_channel_state(ax1o_problem* h, const double* s, double v_rest, int initialized) {
  Problem& p = h->p;
  const int n = p.nx;
  for (int i = 0; i < n; ++i) {
    p.ratio[0][i] = s[i]; p.ratio[1][i] = s[n + i];
    for (int g = 0; g < 3; ++g) p.gate[g][i] = p.gate_new[g][i] = s[(2 + g) * n + i];
  }
  p.v_rest = v_rest;
  p.channels_initialized = initialized;
  p.channels_partially_initialized = false;
}

double ax1o_get_v_rest(const ax1o_problem* h) { return h->p.v_rest; }
void ax1o_membrane_potential(const ax1o_problem* h, const double* u, double* vm) { membrane_potential_mV(h->p, u, vm); }

int ax1o_cell_kernel(const ax1o_problem* h, int i, int j, int which, const double* xl, const double* u, double weight,
                     double* out16) {
  AX1O_TRY cell_kernel(h->p, i, j, which, xl, u, weight, out16); return 0; AX1O_CATCH(1)
}

int ax1o_cell_jacobian(const ax1o_problem* h, int i, int j, int which, const double* xl, double weight, double* out256) {
  AX1O_TRY cell_jacobian(h->p, i, j, which, xl, weight, out256); return 0; AX1O_CATCH(1)
}

int ax1o_membrane_flux_terms(const ax1o_problem* h, const double* u, double* out) {
  AX1O_TRY
  for (int i = 0; i < h->p.nx; ++i) {
    double t[15];
    membrane_flux_terms(h->p, i, u, t);
    for (int k = 0; k < 15; ++k) out[(size_t)k * h->p.nx + i] = t[k];
  }
  return 0;
  AX1O_CATCH(1)
}

int ax1o_residual(const ax1o_problem* h, const double* u, double* r, double* abs_terms) {
  AX1O_TRY residual(h->p, u, r, abs_terms); return 0; AX1O_CATCH(1)
}

int ax1o_jacobian(const ax1o_problem* h, const double* u, double* vals) {
  AX1O_TRY
  BCRS A;
  build_pattern(h->p, A);
  jacobian(h->p, u, A);
  std::memcpy(vals, A.val.data(), A.val.size() * sizeof(double));
  return 0;
  AX1O_CATCH(1)
}

static void load(const Problem& p, const double* vals, BCRS& A) {
  build_pattern(p, A);
  std::memcpy(A.val.data(), vals, A.val.size() * sizeof(double));
}

void ax1o_rownorm(const ax1o_problem* h, double* vals, double* r) {
  BCRS A; load(h->p, vals, A);
  rownorm_scale(A, r);
  std::memcpy(vals, A.val.data(), A.val.size() * sizeof(double));
}

void ax1o_spmv(const ax1o_problem* h, const double* vals, const double* x, double* y) {
  BCRS A; load(h->p, vals, A);
  spmv(A, x, y);
}

int ax1o_ilu_apply(const ax1o_problem* h, const double* vals, int ilu_fill, int slab_w, const double* d, double* v) {
  AX1O_TRY
  BCRS A; load(h->p, vals, A);
  ILU M; ilu_factor(h->p, A, ilu_fill, slab_w, M);
  ilu_apply(M, d, v);
  return 0;
  AX1O_CATCH(1)
}

int ax1o_ilu_num_blocks(const ax1o_problem* h, int ilu_fill, int slab_w) {
  AX1O_TRY
  BCRS A; build_pattern(h->p, A);
  for (int i = 0; i < A.n; ++i) { double* b = A.block(A.find(i, i)); for (int q = 0; q < 4; ++q) b[5 * q] = 1.0; }
  ILU M; ilu_factor(h->p, A, ilu_fill, slab_w, M);
  return (int)M.LU.col.size();
  AX1O_CATCH(-1)
}

int ax1o_solve(const ax1o_problem* h, const double* vals, double* r, double* z, double reduction, int maxit,
               int ilu_fill, int slab_w, ax1o_lin_result* res) {
  AX1O_TRY
  BCRS A; load(h->p, vals, A);
  ILU M; ilu_factor(h->p, A, ilu_fill, slab_w, M);
  std::memset(z, 0, 4 * (size_t)h->p.nv * sizeof(double));
  LinResult lr;
  bicgstab(h->p, A, M, z, r, reduction, maxit, nullptr, lr);
  *res = lr;
  return 0;
  AX1O_CATCH(1)
}

int ax1o_solve_ex(const ax1o_problem* h, const double* vals, double* r, double* z, double reduction, int maxit,
                  int ilu_fill, int slab_w, int krylov, int restart, int amg, int amg_cx, ax1o_lin_result* res) {
  AX1O_TRY
  BCRS A; load(h->p, vals, A);
  ILU M; AMGx H; Prec prec;
  if (amg) {
    amg_build(h->p, A, ilu_fill, slab_w, amg_cx, H);
    prec = [&H](const double* d, double* v) { amg_vcycle(H, d, v); };
  } else {
    ilu_factor(h->p, A, ilu_fill, slab_w, M);
    prec = [&M](const double* d, double* v) { ilu_apply(M, d, v); };
  }
  std::memset(z, 0, 4 * (size_t)h->p.nv * sizeof(double));
  LinResult lr;
  if (krylov == 1) gmres(h->p, A, prec, z, r, reduction, restart, maxit, nullptr, lr);
  else bicgstab(h->p, A, prec, z, r, reduction, maxit, nullptr, lr);
  *res = lr;
  return 0;
  AX1O_CATCH(1)
}

int ax1o_amg_apply(const ax1o_problem* h, const double* vals, int ilu_fill, int slab_w, int amg_cx, const double* d,
                   double* v) {
  AX1O_TRY
  BCRS A; load(h->p, vals, A);
  AMGx H; amg_build(h->p, A, ilu_fill, slab_w, amg_cx, H);
  amg_vcycle(H, d, v);
  return 0;
  AX1O_CATCH(1)
}

int ax1o_newton(ax1o_problem* h, double* u, const ax1o_newton_opts* o, ax1o_newton_result* res) {
  AX1O_TRY
  NewtonResult r;
  newton(h->p, u, NewtonOpts(*o), nullptr, r);
  *res = r;
  return r.status;
  AX1O_CATCH(-1)
}

int ax1o_newton_ext(ax1o_problem* h, double* u, const ax1o_newton_opts* o, ax1o_newton_result* res, int rank,
                    ax1o_halo_add_fn halo_add, ax1o_allreduce_fn allreduce, void* ctx) {
  AX1O_TRY
  Comm c; c.ctx = ctx; c.rank = rank; c.halo_add = halo_add; c.allreduce_sum = allreduce;
  NewtonResult r;
  newton(h->p, u, NewtonOpts(*o), &c, r);
  *res = r;
  return r.status;
  AX1O_CATCH(-1)
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// thread-per-rank world: stand-in for the reference's MPI ranks (additive overlap sum = AddDataHandle
// on All_All_Interface, dot = allreduce; ax1_solverbackend.hh:346-381, 413-437)
namespace {

struct World {
  int n;
  std::vector<Problem*> probs;
  std::vector<double> partial;
  std::vector<std::vector<double>> sendL, sendR;  // packed 3 vertex columns per side
  std::mutex m;
  std::condition_variable cv;
  int count = 0, generation = 0;
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    int gen = generation;
    if (++count == n) { count = 0; ++generation; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != generation; });
  }
};

void pack_cols(const Problem& p, const double* v, int c0, std::vector<double>& buf) {
  buf.resize((size_t)3 * p.nvy * 4);
  size_t o = 0;
  for (int c = c0; c < c0 + 3; ++c)
    for (int j = 0; j < p.nvy; ++j)
      for (int k = 0; k < 4; ++k) buf[o++] = v[4 * (size_t)p.vid(c, j) + k];
}
void add_cols(const Problem& p, double* v, int c0, const std::vector<double>& buf) {
  size_t o = 0;
  for (int c = c0; c < c0 + 3; ++c)
    for (int j = 0; j < p.nvy; ++j)
      for (int k = 0; k < 4; ++k) v[4 * (size_t)p.vid(c, j) + k] += buf[o++];
}

void world_halo_add(void* ctx, int rank, double* v) {
  World& w = *(World*)ctx;
  const Problem& p = *w.probs[rank];
  if (p.left_proc_boundary) pack_cols(p, v, 0, w.sendL[rank]);
  if (p.right_proc_boundary) pack_cols(p, v, p.nvx - 3, w.sendR[rank]);
  w.barrier();
  if (p.left_proc_boundary) add_cols(p, v, 0, w.sendR[rank - 1]);
  if (p.right_proc_boundary) add_cols(p, v, p.nvx - 3, w.sendL[rank + 1]);
  w.barrier();
}

double world_allreduce(void* ctx, int rank, double local) {
  World& w = *(World*)ctx;
  w.partial[rank] = local;
  w.barrier();
  double s = 0.0;
  for (int r = 0; r < w.n; ++r) s += w.partial[r];
  w.barrier();
  return s;
}

}  // namespace

extern "C" {

int ax1o_newton_mt(ax1o_problem** ranks, int nranks, double** u, const ax1o_newton_opts* o, ax1o_newton_result* res) {
  World w;
  w.n = nranks; w.partial.assign(nranks, 0.0); w.sendL.resize(nranks); w.sendR.resize(nranks);
  for (int r = 0; r < nranks; ++r) w.probs.push_back(&ranks[r]->p);
  std::vector<std::thread> th;
  std::vector<int> status(nranks, 0);
  for (int r = 0; r < nranks; ++r)
    th.emplace_back([&, r] {
      Comm c; c.ctx = &w; c.rank = r; c.halo_add = world_halo_add; c.allreduce_sum = world_allreduce;
      NewtonResult nr;
      try { newton(ranks[r]->p, u[r], NewtonOpts(*o), nranks > 1 ? &c : nullptr, nr); }
      catch (std::exception&) { nr.status = -1; }
      res[r] = nr; status[r] = nr.status;
    });
  for (auto& t : th) t.join();
  int s = 0;
  for (int r = 0; r < nranks; ++r) if (status[r]) s = status[r];
  return s;
}

int ax1o_solve_mt(ax1o_problem** ranks, int nranks, double** vals, double** r, double** z, double reduction,
                  int maxit, int ilu_fill, int slab_w, ax1o_lin_result* res) {
  World w;
  w.n = nranks; w.partial.assign(nranks, 0.0); w.sendL.resize(nranks); w.sendR.resize(nranks);
  for (int k = 0; k < nranks; ++k) w.probs.push_back(&ranks[k]->p);
  std::vector<std::thread> th;
  std::vector<int> status(nranks, 0);
  for (int k = 0; k < nranks; ++k)
    th.emplace_back([&, k] {
      try {
        const Problem& p = ranks[k]->p;
        Comm c; c.ctx = &w; c.rank = k; c.halo_add = world_halo_add; c.allreduce_sum = world_allreduce;
        BCRS A; load(p, vals[k], A);
        ILU M; ilu_factor(p, A, ilu_fill, slab_w, M);
        std::memset(z[k], 0, 4 * (size_t)p.nv * sizeof(double));
        LinResult lr;
        bicgstab(p, A, M, z[k], r[k], reduction, maxit, nranks > 1 ? &c : nullptr, lr);
        res[k] = lr;
      } catch (std::exception&) { status[k] = 1; }
    });
  for (auto& t : th) t.join();
  for (int k = 0; k < nranks; ++k) if (status[k]) return 1;
  return 0;
}

}  // extern "C"
