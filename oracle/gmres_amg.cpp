// oracle/gmres_amg.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's second linear solver backend, used by the myelin build
// (dune/ax1/acme2_cyl/common/acme2_cyl_setup.hh:749-760 -> ISTLBackend_GMRES_AMG_ILU0,
// dune/ax1/common/ax1_solverbackend.hh:73-302):
//   gmres        Dune::RestartedGMResSolver (dune-istl 2.3 solvers.hh), restart 100, recalcDefect
//                false (ax1_solverbackend.hh:201-204)                                           [ext]
//   amg_*        stand-in for Dune::Amg::AMG (non-smoothed aggregation, SeqILU0 smoother,
//                1 pre + 1 post smoothing step, ax1_solverbackend.hh:183-193). Dune's greedy graph
//                aggregation [ext, dune-istl paamg/aggregates.hh] is absent from /root/reference and
//                not reproducible bit for bit; what is restated here is the aggregation the DEVICE
//                uses (csrc/amg.cu: aggregates = cx consecutive vertex columns at full radial
//                resolution), so this part is the checker of the device cycle, not of Dune's.
#include "oracle.hpp"
#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace ax1o {

namespace {

// Dune::RestartedGMResSolver::generatePlaneRotation / applyPlaneRotation [ext]
void generate_plane_rotation(double dx, double dy, double& cs, double& sn) {
  if (dy == 0.0) { cs = 1.0; sn = 0.0; }
  else if (std::fabs(dy) > std::fabs(dx)) {
    const double temp = dx / dy;
    sn = 1.0 / std::sqrt(1.0 + temp * temp);
    cs = temp * sn;
  } else {
    const double temp = dy / dx;
    cs = 1.0 / std::sqrt(1.0 + temp * temp);
    sn = temp * cs;
  }
}
void apply_plane_rotation(double& dx, double& dy, double cs, double sn) {
  const double temp = cs * dx + sn * dy;
  dy = -sn * dx + cs * dy;
  dx = temp;
}

}  // namespace

// Left-preconditioned restarted GMRES with modified Gram-Schmidt and Givens rotations; the
// convergence test is on the PRECONDITIONED residual norm (norm_0 = |M^-1 (b - A x0)|).
void gmres(const Problem& p, const BCRS& A, const Prec& M, double* x, double* b, double reduction,
           int restart, int maxit, const Comm* comm, LinResult& res) {
  const size_t n = 4 * (size_t)p.nv;
  const int m = restart;
  auto dot = [&](const double* a, const double* c) { return masked_dot(p, a, c, comm); };
  auto norm = [&](const double* a) { return std::sqrt(dot(a, a)); };
  std::vector<double> s(m + 1), cs(m), sn(m), w(n), tmp(n), dd;
  std::vector<std::vector<double>> H(m + 1, std::vector<double>(m, 0.0));
  std::vector<std::vector<double>> v(m + 1, std::vector<double>(n, 0.0));
  res = LinResult();
  // b = b - A x ; v[0] = M^-1 b ; beta = |v[0]|
  op_apply(p, A, x, tmp.data(), comm);
  for (size_t k = 0; k < n; ++k) b[k] -= tmp[k];
  std::fill(v[0].begin(), v[0].end(), 0.0);
  prec_apply(p, M, b, v[0].data(), dd, comm);
  double beta = norm(v[0].data());
  double norm_0 = beta;
  if (norm_0 == 0.0) norm_0 = 1.0;
  double nrm = beta;
  int i = 0, j = 1;
  bool converged = false;
  if (nrm <= reduction * norm_0) {
    res.converged = 1; res.iterations = 0; res.reduction = nrm / norm_0;
    return;
  }
  while (j <= maxit && !converged) {
    for (size_t k = 0; k < n; ++k) v[0][k] *= (1.0 / beta);
    for (i = 1; i <= m; ++i) s[i] = 0.0;
    s[0] = beta;
    for (i = 0; i < m && j <= maxit && !converged; ++i, ++j) {
      std::fill(w.begin(), w.end(), 0.0);
      op_apply(p, A, v[i].data(), v[i + 1].data(), comm);       // v[i+1] as temporary
      prec_apply(p, M, v[i + 1].data(), w.data(), dd, comm);
      for (int k = 0; k < i + 1; ++k) {
        H[k][i] = dot(w.data(), v[k].data());
        const double h = H[k][i];
        const double* vk = v[k].data();
        for (size_t q = 0; q < n; ++q) w[q] += -h * vk[q];      // w.axpy(-H[k][i], v[k])
      }
      H[i + 1][i] = norm(w.data());
      if (H[i + 1][i] == 0.0) throw std::runtime_error("breakdown in GMRes - |w| == 0.0");
      const double inv = 1.0 / H[i + 1][i];
      for (size_t q = 0; q < n; ++q) v[i + 1][q] = w[q] * inv;
      for (int k = 0; k < i; ++k) apply_plane_rotation(H[k][i], H[k + 1][i], cs[k], sn[k]);
      generate_plane_rotation(H[i][i], H[i + 1][i], cs[i], sn[i]);
      apply_plane_rotation(H[i][i], H[i + 1][i], cs[i], sn[i]);
      apply_plane_rotation(s[i], s[i + 1], cs[i], sn[i]);
      nrm = std::fabs(s[i + 1]);
      if (nrm < reduction * norm_0) converged = true;
    }
    // update(w, i-1, H, s, v): back substitution, w = sum_j y_j v[j]; x += w
    std::vector<double> y(s.begin(), s.end());
    for (int a = i - 1; a >= 0; --a) {
      y[a] /= H[a][a];
      for (int c = a - 1; c >= 0; --c) y[c] -= H[c][a] * y[a];
    }
    std::fill(w.begin(), w.end(), 0.0);
    for (int c = 0; c < i; ++c) {
      const double yc = y[c];
      const double* vc = v[c].data();
      for (size_t q = 0; q < n; ++q) w[q] += yc * vc[q];
    }
    for (size_t q = 0; q < n; ++q) x[q] += w[q];
    // b -= A w ; v[0] = M^-1 b ; beta = |v[0]| for the next cycle
    op_apply(p, A, w.data(), tmp.data(), comm);
    for (size_t q = 0; q < n; ++q) b[q] -= tmp[q];
    if (!converged) {
      std::fill(v[0].begin(), v[0].end(), 0.0);
      prec_apply(p, M, b, v[0].data(), dd, comm);
      beta = norm(v[0].data());
      nrm = beta;
      if (nrm < reduction * norm_0) converged = true;   // the true preconditioned residual after the restart
    }
  }
  res.converged = converged;
  res.iterations = j - 1;       // number of Krylov steps performed
  res.it_half = j - 1;
  res.reduction = nrm / norm_0;
}

// ---------------------------------------------------------------------------------------------
// x-aggregation multigrid
namespace {

// 9-point block pattern on a tensor grid of nvx x nvy vertices, x fastest, columns ascending
void stencil_pattern(int nvx, int nvy, BCRS& A) {
  A.n = nvx * nvy;
  A.rowptr.assign(A.n + 1, 0);
  A.col.clear();
  for (int j = 0; j < nvy; ++j)
    for (int i = 0; i < nvx; ++i) {
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          const int ii = i + di, jj = j + dj;
          if (ii < 0 || ii >= nvx || jj < 0 || jj >= nvy) continue;
          A.col.push_back(ii + nvx * jj);
        }
      A.rowptr[i + nvx * j + 1] = (int)A.col.size();
    }
  A.val.assign(A.col.size() * 16, 0.0);
}

// A_c = P^T A_f P with piecewise-constant P over aggregates of cx vertex columns; constrained fine
// DOFs (mask, finest level only) take no part; a coarse scalar row without free fine DOF -> unit row.
// Loop order as in csrc/amg.cu: galerkin_x_kernel (per coarse entry: fine row i ascending, di ascending).
void galerkin_x(int nvx_f, int nvx_c, int nvy, int cx, const BCRS& Af, BCRS& Ac, const uint8_t* mask) {
  stencil_pattern(nvx_c, nvy, Ac);
  for (int j = 0; j < nvy; ++j)
    for (int I = 0; I < nvx_c; ++I) {
      const int vc = I + nvx_c * j;
      const int i_end = std::min(cx * I + cx, nvx_f);
      for (int a = 0; a < 4; ++a) {
        int free_rows = 0;
        for (int i = cx * I; i < i_end; ++i)
          if (!(mask && ((mask[i + nvx_f * j] >> a) & 1))) ++free_rows;
        for (int dj = -1; dj <= 1; ++dj)
          for (int DI = -1; DI <= 1; ++DI) {
            const int II = I + DI, jj = j + dj;
            if (II < 0 || II >= nvx_c || jj < 0 || jj >= nvy) continue;
            const int kc = Ac.find(vc, II + nvx_c * jj);
            for (int b = 0; b < 4; ++b) {
              double sum = 0.0;
              for (int i = cx * I; i < i_end; ++i) {
                const int vf = i + nvx_f * j;
                if (mask && ((mask[vf] >> a) & 1)) continue;
                for (int di = -1; di <= 1; ++di) {
                  const int ii = i + di;
                  if (ii < 0 || ii >= nvx_f) continue;
                  if (ii / cx - I != DI) continue;
                  if (mask && ((mask[ii + nvx_f * jj] >> b) & 1)) continue;
                  const int kf = Af.find(vf, ii + nvx_f * jj);
                  if (kf >= 0) sum += Af.block(kf)[4 * a + b];
                }
              }
              if (free_rows == 0 && DI == 0 && dj == 0 && a == b) sum = 1.0;
              Ac.block(kc)[4 * a + b] = sum;
            }
          }
      }
    }
}

}  // namespace

void amg_build(const Problem& p, const BCRS& A, int fill_level, int slab_w, int cx, AMGx& H) {
  if (cx <= 0) cx = 4;
  if (slab_w <= 0 || slab_w > p.nvx) slab_w = p.nvx;
  H.cx = cx; H.nvy = p.nvy;
  H.mask = p.dirichlet;
  H.lv.clear();
  H.lv.emplace_back();
  H.lv[0].nvx = p.nvx;
  H.lv[0].A = A;
  ilu_factor_nvx(p.nvx, A, fill_level, slab_w, H.lv[0].M);
  int nvx = p.nvx;
  while (nvx > 2 * cx) {
    const int nvx_f = nvx;
    nvx = (nvx + cx - 1) / cx;
    H.lv.emplace_back();
    AMGx::Level& C = H.lv.back();
    const AMGx::Level& F = H.lv[H.lv.size() - 2];
    C.nvx = nvx;
    galerkin_x(nvx_f, nvx, p.nvy, cx, F.A, C.A, H.lv.size() == 2 ? H.mask.data() : nullptr);
    ilu_factor_nvx(nvx, C.A, 0, std::min(slab_w, nvx), C.M);     // coarse smoothers: block ILU0
    C.b.assign(4 * (size_t)nvx * p.nvy, 0.0);
    C.x.assign(4 * (size_t)nvx * p.nvy, 0.0);
  }
  H.t.assign(4 * (size_t)p.nv, 0.0);
  H.s.assign(4 * (size_t)p.nv, 0.0);
}

// v = one V(1,1) cycle applied to d, zero initial guess on every level (csrc/driver.cu: amg_vcycle)
void amg_vcycle(AMGx& H, const double* d, double* v) {
  const int nl = (int)H.lv.size() - 1;   // number of coarse levels
  const int cx = H.cx, nvy = H.nvy;
  auto rhs = [&](int l) -> const double* { return l == 0 ? d : H.lv[l].b.data(); };
  auto sol = [&](int l) -> double* { return l == 0 ? v : H.lv[l].x.data(); };
  auto defect = [&](int l) {   // t = b_l - A_l x_l
    const AMGx::Level& L = H.lv[l];
    spmv(L.A, sol(l), H.t.data());
    const double* b = rhs(l);
    const size_t n = 4 * (size_t)L.nvx * nvy;
    for (size_t q = 0; q < n; ++q) H.t[q] = b[q] - H.t[q];
  };
  ilu_apply(H.lv[0].M, d, v);
  for (int l = 0; l < nl; ++l) {
    defect(l);
    const int nvx_f = H.lv[l].nvx, nvx_c = H.lv[l + 1].nvx;
    const uint8_t* mask = l == 0 ? H.mask.data() : nullptr;
    double* bc = H.lv[l + 1].b.data();
    for (int j = 0; j < nvy; ++j)
      for (int I = 0; I < nvx_c; ++I)
        for (int a = 0; a < 4; ++a) {
          double s = 0.0;
          for (int i = cx * I; i < std::min(cx * I + cx, nvx_f); ++i) {
            const int vf = i + nvx_f * j;
            if (mask && ((mask[vf] >> a) & 1)) continue;
            s += H.t[4 * (size_t)vf + a];
          }
          bc[4 * (size_t)(I + nvx_c * j) + a] = s;
        }
    ilu_apply(H.lv[l + 1].M, bc, H.lv[l + 1].x.data());
  }
  for (int l = nl - 1; l >= 0; --l) {
    const int nvx_f = H.lv[l].nvx, nvx_c = H.lv[l + 1].nvx;
    const uint8_t* mask = l == 0 ? H.mask.data() : nullptr;
    double* xf = sol(l);
    const double* xc = H.lv[l + 1].x.data();
    for (int j = 0; j < nvy; ++j)
      for (int i = 0; i < nvx_f; ++i)
        for (int a = 0; a < 4; ++a) {
          const int vf = i + nvx_f * j;
          if (mask && ((mask[vf] >> a) & 1)) continue;
          xf[4 * (size_t)vf + a] += xc[4 * (size_t)(i / cx + nvx_c * j) + a];
        }
    defect(l);
    ilu_apply(H.lv[l].M, H.t.data(), H.s.data());
    const size_t n = 4 * (size_t)nvx_f * nvy;
    for (size_t q = 0; q < n; ++q) xf[q] += H.s[q];
  }
}

}  // namespace ax1o
