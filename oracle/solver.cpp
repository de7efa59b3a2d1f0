// oracle/solver.cpp -- TEST INFRASTRUCTURE ONLY.
// Restatement of the linear/non-linear solver stack the reference selects
// (dune/ax1/acme2_cyl/common/acme2_cyl_setup.hh:727-811, 822-869):
//   BiCGSTAB            Dune::BiCGSTABSolver (dune-istl 2.3, solvers.hh)                        [ext]
//   block ILU(n)        Dune::SeqILUn -> bilu_decomposition / bilu0_decomposition / bilu_backsolve [ext]
//   overlap wrappers    PDELab ovlpistlsolverbackend.hh, restated in dune/ax1/common/ax1_solverbackend.hh:306-443
//   Newton              PDELab newton.hh [ext] + dune/ax1/common/ax1_newton.hh:194-205,216-293,308-503
// [ext] = third-party code absent from /root/reference; algorithm restated from the pinned
// versions (dune-istl 35b879c, dune-pdelab 9f4d7a9), see SURVEY.md App. A.9/A.10.
#include "oracle.hpp"
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <stdexcept>

namespace ax1o {

void spmv(const BCRS& A, const double* x, double* y) {
  for (int i = 0; i < A.n; ++i) {
    double acc[4] = {0, 0, 0, 0};
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k) {
      const double* b = A.block(k);
      const double* xx = &x[4 * A.col[k]];
      for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) acc[a] += b[4 * a + c] * xx[c];
    }
    for (int a = 0; a < 4; ++a) y[4 * i + a] = acc[a];
  }
}

namespace {

void mat4_mul(const double* A, const double* B, double* C) {  // C = A*B
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
      C[4 * i + j] = s;
    }
}

void mat4_invert(double* M) {  // Gauss-Jordan with partial pivoting (FieldMatrix::invert analogue)
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = M[4 * i + j]; a[i][4 + j] = (i == j); }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0) throw std::runtime_error("singular diagonal block in ILU");
    if (piv != c) for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
    double inv = 1.0 / a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] *= inv;
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      double f = a[r][c];
      if (f != 0.0) for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
    }
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) M[4 * i + j] = a[i][4 + j];
}

}  // namespace

// bilu_decomposition(A, n, ILU) of dune-istl ilu.hh: symbolic phase with "generation" numbers,
// then bilu0_decomposition on the enlarged pattern. slab_w > 0 first drops all blocks that couple
// vertex columns of different sub-slabs (our many-slabs-per-rank reading of additive Schwarz).
void ilu_factor(const Problem& p, const BCRS& A, int n_fill, int slab_w, ILU& f) {
  ilu_factor_nvx(p.nvx, A, n_fill, slab_w, f, p.ilu_jcut);
}
void ilu_factor_nvx(int nvx, const BCRS& A, int n_fill, int slab_w, ILU& f, int jcut) {
  const int N = A.n;
  auto slab_of = [&](int v) {
    return 2 * (slab_w > 0 ? (v % nvx) / slab_w : 0) + ((jcut > 0 && v / nvx >= jcut) ? 1 : 0);
  };
  BCRS& LU = f.LU;
  LU.n = N;
  LU.rowptr.assign(N + 1, 0);
  LU.col.clear();
  std::vector<int> gen;  // generation per entry
  std::vector<std::map<int, int>> rows(n_fill > 0 ? N : 0);
  for (int i = 0; i < N; ++i) {
    std::map<int, int> rowpattern;
    for (int k = A.rowptr[i]; k < A.rowptr[i + 1]; ++k)
      if (slab_of(A.col[k]) == slab_of(i)) rowpattern[A.col[k]] = 0;
    if (n_fill > 0) {
      for (auto ik = rowpattern.begin(); ik != rowpattern.end() && ik->first < i; ++ik) {
        if (ik->second < n_fill) {
          const std::map<int, int>& rk = rows[ik->first];
          for (auto kj = rk.upper_bound(ik->first); kj != rk.end(); ++kj) {
            int generation = kj->second;
            if (generation < n_fill && rowpattern.find(kj->first) == rowpattern.end())
              rowpattern[kj->first] = generation + 1;
          }
        }
      }
      rows[i] = rowpattern;
    }
    for (auto& e : rowpattern) { LU.col.push_back(e.first); gen.push_back(e.second); }
    LU.rowptr[i + 1] = (int)LU.col.size();
  }
  LU.val.assign(LU.col.size() * 16, 0.0);
  f.diag.assign(N, -1);
  for (int i = 0; i < N; ++i)
    for (int k = LU.rowptr[i]; k < LU.rowptr[i + 1]; ++k) {
      if (LU.col[k] == i) f.diag[i] = k;
      int ka = A.find(i, LU.col[k]);
      if (ka >= 0) std::memcpy(LU.block(k), A.block(ka), 16 * sizeof(double));
    }
  // bilu0_decomposition
  for (int i = 0; i < N; ++i) {
    const int endi = LU.rowptr[i + 1];
    int ij = LU.rowptr[i];
    for (; LU.col[ij] < i; ++ij) {
      const int j = LU.col[ij];
      const int jj = f.diag[j];
      double tmp[16];
      mat4_mul(LU.block(ij), LU.block(jj), tmp);  // (*ij).rightmultiply(*jj), jj holds the inverse
      std::memcpy(LU.block(ij), tmp, sizeof tmp);
      int jk = jj + 1, ik = ij + 1;
      const int endj = LU.rowptr[j + 1];
      while (ik != endi && jk != endj) {
        if (LU.col[ik] == LU.col[jk]) {
          double B[16];
          mat4_mul(LU.block(ij), LU.block(jk), B);  // B.leftmultiply(*ij)
          double* t = LU.block(ik);
          for (int q = 0; q < 16; ++q) t[q] -= B[q];
          ++ik; ++jk;
        } else if (LU.col[ik] < LU.col[jk]) ++ik;
        else ++jk;
      }
    }
    if (LU.col[ij] != i) throw std::runtime_error("diagonal entry missing");
    mat4_invert(LU.block(ij));
  }
}

void ilu_apply(const ILU& f, const double* d, double* v) {  // bilu_backsolve
  const BCRS& LU = f.LU;
  const int N = LU.n;
  for (int i = 0; i < N; ++i) {
    double rhs[4] = {d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]};
    for (int k = LU.rowptr[i]; LU.col[k] < i; ++k) {
      const double* b = LU.block(k);
      const double* x = &v[4 * LU.col[k]];
      for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) rhs[a] -= b[4 * a + c] * x[c];
    }
    for (int a = 0; a < 4; ++a) v[4 * i + a] = rhs[a];
  }
  for (int i = N - 1; i >= 0; --i) {
    double rhs[4] = {v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]};
    for (int k = f.diag[i] + 1; k < LU.rowptr[i + 1]; ++k) {
      const double* b = LU.block(k);
      const double* x = &v[4 * LU.col[k]];
      for (int a = 0; a < 4; ++a)
        for (int c = 0; c < 4; ++c) rhs[a] -= b[4 * a + c] * x[c];
    }
    const double* dinv = LU.block(f.diag[i]);
    for (int a = 0; a < 4; ++a) {
      double s = 0.0;
      for (int c = 0; c < 4; ++c) s += dinv[4 * a + c] * rhs[c];
      v[4 * i + a] = s;
    }
  }
}

double masked_dot(const Problem& p, const double* a, const double* b, const Comm* comm) {
  double s = 0.0;
  if (!comm) {
    const size_t n = 4 * (size_t)p.nv;
    for (size_t k = 0; k < n; ++k) s += a[k] * b[k];
    return s;
  }
  // OVLPScalarProduct: owner-masked local dot + allreduce (ax1_solverbackend.hh:413-437)
  for (int j = 0; j < p.nvy; ++j)
    for (int i = p.own_lo; i <= p.own_hi; ++i) {
      const size_t o = 4 * (size_t)p.vid(i, j);
      for (int c = 0; c < 4; ++c) s += a[o + c] * b[o + c];
    }
  return comm->allreduce_sum(comm->ctx, comm->rank, s);
}

void zero_constrained(const Problem& p, double* v) {
  for (int k = 0; k < p.nv; ++k)
    if (p.dirichlet[k])
      for (int c = 0; c < 4; ++c)
        if ((p.dirichlet[k] >> c) & 1) v[4 * k + c] = 0.0;
}

void op_apply(const Problem& p, const BCRS& A, const double* x, double* y, const Comm* comm) {
  spmv(A, x, y);
  if (comm) zero_constrained(p, y);  // OverlappingOperator::apply
}

// OverlappingWrappedPreconditioner::apply around any rank-local preconditioner
void prec_apply(const Problem& p, const Prec& M, const double* d, double* v, std::vector<double>& dd,
                const Comm* comm) {
  if (!comm) { M(d, v); return; }
  dd.assign(d, d + 4 * (size_t)p.nv);
  zero_constrained(p, dd.data());
  M(dd.data(), v);
  comm->halo_add(comm->ctx, comm->rank, v);
}

void bicgstab(const Problem& p, const BCRS& A, const ILU& M, double* x, double* b, double reduction,
              int maxit, const Comm* comm, LinResult& res) {
  bicgstab(p, A, Prec([&M](const double* d, double* v) { ilu_apply(M, d, v); }), x, b, reduction, maxit, comm, res);
}

void bicgstab(const Problem& p, const BCRS& A, const Prec& M, double* x, double* b, double reduction,
              int maxit, const Comm* comm, LinResult& res) {
  const size_t n = 4 * (size_t)p.nv;
  const double EPSILON = 1e-80;
  std::vector<double> r(b, b + n), rt(n), pv(n, 0.0), v(n, 0.0), y(n), t(n), tmp(n), dd;
  auto dot = [&](const double* a, const double* c) { return masked_dot(p, a, c, comm); };
  auto norm = [&](const double* a) { return std::sqrt(dot(a, a)); };
  // r = b - A x
  op_apply(p, A, x, tmp.data(), comm);
  for (size_t k = 0; k < n; ++k) r[k] -= tmp[k];
  rt = r;
  double rho = 1, alpha = 1, omega = 1, rho_new, beta, h;
  double nrm, norm_0;
  nrm = norm_0 = norm(r.data());
  res = LinResult();
  double it = 0;
  if (nrm < reduction * norm_0 || nrm < 1e-30) {
    res.converged = 1; res.iterations = 0; res.reduction = 0;
    std::memcpy(b, r.data(), n * sizeof(double));
    return;
  }
  for (it = 0.5; it < maxit; it += .5) {
    rho_new = dot(rt.data(), r.data());
    if (std::fabs(rho) <= EPSILON) throw std::runtime_error("breakdown in BiCGSTAB - rho");
    if (std::fabs(omega) <= EPSILON) throw std::runtime_error("breakdown in BiCGSTAB - omega");
    if (it < 1) pv = r;
    else {
      beta = (rho_new / rho) * (alpha / omega);
      for (size_t k = 0; k < n; ++k) pv[k] += -omega * v[k];
      for (size_t k = 0; k < n; ++k) pv[k] *= beta;
      for (size_t k = 0; k < n; ++k) pv[k] += r[k];
    }
    std::fill(y.begin(), y.end(), 0.0);
    prec_apply(p, M, pv.data(), y.data(), dd, comm);
    op_apply(p, A, y.data(), v.data(), comm);
    h = dot(rt.data(), v.data());
    if (std::fabs(h) < EPSILON) throw std::runtime_error("h=0 in BiCGSTAB");
    alpha = rho_new / h;
    for (size_t k = 0; k < n; ++k) x[k] += alpha * y[k];
    for (size_t k = 0; k < n; ++k) r[k] += -alpha * v[k];
    nrm = norm(r.data());
    if (nrm < reduction * norm_0) { res.converged = 1; break; }
    it += .5;
    std::fill(y.begin(), y.end(), 0.0);
    prec_apply(p, M, r.data(), y.data(), dd, comm);
    op_apply(p, A, y.data(), t.data(), comm);
    omega = dot(t.data(), r.data()) / dot(t.data(), t.data());
    for (size_t k = 0; k < n; ++k) x[k] += omega * y[k];
    for (size_t k = 0; k < n; ++k) r[k] += -omega * t[k];
    rho = rho_new;
    nrm = norm(r.data());
    if (nrm < reduction * norm_0 || nrm < 1e-30) { res.converged = 1; break; }
  }
  it = std::min((double)maxit, it);
  res.it_half = it;
  res.iterations = (int)std::ceil(it);
  res.reduction = nrm / norm_0;
  std::memcpy(b, r.data(), n * sizeof(double));
}

// ---------------------------------------------------------------------------------------------
// Hackbusch-Reusken line search, strategy hackbuschReuskenAcceptBestNoThrow (ax1_newton.hh:308-503).
// `eval` recomputes the defect norm of the current u (NaN allowed: a NaN trial is simply retried
// with a smaller lambda, :384-392); returns the defect of the accepted u. `trace` (optional)
// receives every trial step length in evaluation order.
double line_search(size_t n, double* u, const double* z, double* prev_u, double cur_defect, double prev_defect,
                   int maxit, const std::function<double()>& eval, std::vector<double>* trace) {
  auto axpy_u = [&](double lam) { for (size_t k = 0; k < n; ++k) u[k] += -lam * z[k]; };
  auto trial = [&](double lam) { axpy_u(lam); if (trace) trace->push_back(lam); return eval(); };
  double lambda = 1.0, best_lambda = 0.0, best_defect = cur_defect, d;
  std::memcpy(prev_u, u, n * sizeof(double));
  int i = 0;
  while (true) {
    d = trial(lambda);
    if (d <= (1.0 - lambda / 4) * prev_defect) break;
    if (d < best_defect) { best_defect = d; best_lambda = lambda; }
    if (++i >= maxit) {
      if (best_lambda == 0.0) {  // no trial improved: accept the full step anyway (:462-474)
        best_lambda = 1.0;
        std::memcpy(u, prev_u, n * sizeof(double));
        d = trial(best_lambda);
        if (!std::isfinite(d)) return d;  // outside the search's try block: NewtonDefectError ends the solve
      }
      if (best_lambda != lambda) {
        std::memcpy(u, prev_u, n * sizeof(double));
        d = trial(best_lambda);
      }
      break;
    }
    lambda *= 0.5;
    std::memcpy(u, prev_u, n * sizeof(double));
  }
  return d;
}

// ---------------------------------------------------------------------------------------------
// Newton: PDELab NewtonSolver::apply [ext] with Ax1Newton::{defect,prepare_step,line_search}
void newton(Problem& p, double* u, const NewtonOpts& o, const Comm* comm, NewtonResult& res) {
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const size_t n = 4 * (size_t)p.nv;
  std::vector<double> r(n), z(n), prev_u(n);
  BCRS A;
  build_pattern(p, A);
  ILU M;
  AMGx H;
  res = NewtonResult();
  auto t_start = clk::now();
  auto defect = [&]() -> bool {  // false on NaN
    residual(p, u, r.data());
    res.residual_evals++;
    res.defect = std::sqrt(masked_dot(p, r.data(), r.data(), comm));
    return std::isfinite(res.defect);
  };
  res.reduction = 1.0; res.conv_rate = 1.0;
  if (!defect()) { res.status = NEWTON_DEFECT_NAN; return; }
  res.first_defect = res.defect;
  double prev_defect = res.defect;
  double linear_reduction = o.min_lin_reduction;
  while (true) {
    // terminate()
    if (!(o.force_iteration && res.iterations == 0)) {
      res.converged = res.defect < o.abs_limit || res.defect < res.first_defect * o.reduction;
      if (o.fixed_work) res.converged = res.iterations >= o.maxit;
      if (res.converged) break;
      if (res.iterations >= o.maxit) { res.status = NEWTON_NOT_CONVERGED; break; }
    }
    // prepare_step: preAssembly + Jacobian + linear reduction (ax1_newton.hh:216-293)
    auto t0 = clk::now();
    update_state(p, u);
    jacobian(p, u, A);
    if (o.fixed_lin_reduction) linear_reduction = o.min_lin_reduction;
    else {
      double stop_defect = std::max(res.first_defect * o.reduction, o.abs_limit);
      if (stop_defect / (10 * res.defect) > res.defect * res.defect / (prev_defect * prev_defect))
        linear_reduction = stop_defect / (10 * res.defect);
      else
        linear_reduction = std::min(o.min_lin_reduction, res.defect * res.defect / (prev_defect * prev_defect));
    }
    prev_defect = res.defect;
    if (p.cfg.use_rownorm_preconditioner) rownorm_scale(A, r.data());
    auto t1 = clk::now();
    res.t_assembler += secs(t0, t1);
    // linear solve
    std::fill(z.begin(), z.end(), 0.0);
    LinResult lr;
    try {
      Prec prec;
      if (o.amg) {
        amg_build(p, A, o.ilu_fill, o.slab_w, o.amg_cx, H);
        prec = [&H](const double* d, double* v) { amg_vcycle(H, d, v); };
      } else {
        ilu_factor(p, A, o.ilu_fill, o.slab_w, M);
        prec = [&M](const double* d, double* v) { ilu_apply(M, d, v); };
      }
      // fixed work (throughput runs): exactly lin_maxit Krylov iterations per solve, as the device driver does
      // (reduction 0 is never reached), not "up to lin_maxit"
      const double lin_red = o.fixed_work ? 0.0 : linear_reduction;
      if (o.krylov == 1) gmres(p, A, prec, z.data(), r.data(), lin_red, o.restart, o.lin_maxit, comm, lr);
      else bicgstab(p, A, prec, z.data(), r.data(), lin_red, o.lin_maxit, comm, lr);
    } catch (std::exception&) {
      res.t_lin_solv += secs(t1, clk::now());
      res.status = NEWTON_LINSOLVE_FAILED;
      break;
    }
    res.t_lin_solv += secs(t1, clk::now());
    res.lin_iterations += lr.iterations;
    if (!lr.converged && !o.fixed_work) { res.status = NEWTON_LINSOLVE_FAILED; break; }
    // line search (ax1_newton.hh:308-503), strategy hackbuschReuskenAcceptBestNoThrow
    auto t2 = clk::now();
    auto axpy_u = [&](double lam) { for (size_t k = 0; k < n; ++k) u[k] += -lam * z[k]; };
    if (o.disable_line_search || o.fixed_work) {
      axpy_u(1.0);
      update_state(p, u);
      if (!defect()) { res.status = NEWTON_DEFECT_NAN; break; }
    } else {
      res.defect = line_search(n, u, z.data(), prev_u.data(), res.defect, prev_defect, o.line_search_maxit,
                               [&]() { update_state(p, u); defect(); return res.defect; }, nullptr);
      if (!std::isfinite(res.defect)) { res.status = NEWTON_DEFECT_NAN; break; }
    }
    res.t_assembler += secs(t2, clk::now());
    res.reduction = res.defect / res.first_defect;
    res.iterations++;
    res.conv_rate = std::pow(res.reduction, 1.0 / res.iterations);
  }
  res.t_elapsed = secs(t_start, clk::now());
}

}  // namespace ax1o
