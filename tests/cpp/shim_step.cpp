// tests/cpp/shim_step.cpp -- C++ host driver over the PDELab-concept shim (include/ax1gpu_pdelab_shim.hh):
// one implicit-Euler time step of acme2_cyl's time loop (acme2_cyl_simulation.hh:457, 589, 916-927)
//   gfMembFlux.updateState(time, dt) -> OneStepMethod::apply -> Ax1Newton::apply -> acceptTimeStep
// on a problem read from a flat binary file written by tests/test_cpp_shim.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ax1gpu_pdelab_shim.hh"

typedef std::vector<double> U;
typedef std::vector<double> M;

template <class T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (fread(v.data(), sizeof(T), n, f) != n) { std::fprintf(stderr, "short read\n"); std::exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: shim_step <problem.bin> <out.bin>\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<int> hdr = rd<int>(f, 4);  // nx, ny, slab_w, use_backend_path
  const int nx = hdr[0], ny = hdr[1];
  const size_t nv = (size_t)(nx + 1) * (ny + 1);
  std::vector<double> x = rd<double>(f, nx + 1), y = rd<double>(f, ny + 1);
  std::vector<unsigned char> cs = rd<unsigned char>(f, (size_t)nx * ny);
  std::vector<double> eps = rd<double>(f, nx), nvol = rd<double>(f, nv);
  std::vector<unsigned char> dm = rd<unsigned char>(f, nv);
  std::vector<double> gl = rd<double>(f, nx), gn = rd<double>(f, nx), gk = rd<double>(f, nx);
  U u = rd<double>(f, 4 * nv);
  std::fclose(f);

  ax1gpu_params prm;
  ax1gpu_params_default(&prm);
  ax1gpu_grid grid = {nx, ny, x.data(), y.data(), 0, 0};
  try {
    ax1gpu::Device dev(grid, cs.data(), eps.data(), nvol.data(), dm.data(), prm, gl.data(), gn.data(), gk.data());
    ax1gpu::MembraneFluxShim<U> gfMembFlux(dev);
    ax1gpu::GridOperatorShim<U, M> igo(dev);
    ax1gpu::SolverBackendShim<U, M> ls(dev, 5000, hdr[2]);
    ax1gpu::NewtonShim<U> pdesolver(dev);
    pdesolver.setReduction(1e-9); pdesolver.setAbsoluteLimit(1e-14); pdesolver.setMinLinearReduction(1e-10);
    pdesolver.setSlabWidth(hdr[2]);
    const double time = 0.0, dt = 1.0;
    if (hdr[3] == 3) {
      // the Mori build's step body through MoriStepShim (acme2_cyl_mori_simulation.hh:562-600)
      U uold = u;
      gfMembFlux.updateState(time, dt, uold);
      ax1gpu::MoriStepShim<U> mori(dev);
      mori.setReduction(1e-10); mori.setMinDefect(1e-12); mori.setLinearSolverMaxIterations(500);
      mori.options().slab_w = hdr[2];
      mori.apply(time, dt, uold, u);
      gfMembFlux.acceptTimeStep();
      const ax1gpu_mori_result& mr = mori.result();
      std::printf("mori: pot its %d defect %.3e -> %.3e, con its %d defect %.3e -> %.3e\n", mr.pot_iterations,
                  mr.pot_first_defect, mr.pot_defect, mr.con_iterations, mr.con_first_defect, mr.con_defect);
      FILE* o = std::fopen(argv[2], "wb");
      std::fwrite(u.data(), sizeof(double), u.size(), o);
      std::fclose(o);
      return 0;
    }
    if (hdr[3] == 2) {
      // the same step through the library's own time loop (state stays on the device)
      ax1gpu::TimeLoopShim<U> loop(dev, time, dt);
      ax1gpu_newton_opts& no = loop.newtonOptions();
      no.reduction = 1e-9; no.abs_limit = 1e-14; no.min_lin_reduction = 1e-10; no.slab_w = hdr[2];
      loop.load(u);
      const ax1gpu_step_result& sr = loop.step();
      loop.store(u);
      std::printf("time %.3f dt %.3f newton its %d lin its %d vm [%.4f, %.4f]\n", sr.time, sr.dt, sr.newton_iterations,
                  sr.lin_iterations, sr.vm_min, sr.vm_max);
      FILE* o = std::fopen(argv[2], "wb");
      std::fwrite(u.data(), sizeof(double), u.size(), o);
      std::fclose(o);
      return 0;
    }
    U uold = u;
    gfMembFlux.updateState(time, dt, uold);
    igo.preStep(time, dt, uold);
    double lin_red = 0;
    unsigned lin_it = 0;
    if (hdr[3]) {
      // exercise the IGO / LS seam the way PDELab's Newton would: r = residual(u); A = jacobian(u); z = A^-1 r
      U r(4 * nv), z(4 * nv, 0.0);
      M A;
      igo.residual(u, r);
      igo.jacobian(u, A);
      const double d0 = ls.norm(r);
      ls.apply(A, z, r, 1e-8);
      lin_red = ls.norm(r) / d0; lin_it = ls.result().iterations;
      if (!ls.result().converged) { std::fprintf(stderr, "backend did not converge\n"); return 3; }
    }
    pdesolver.apply(u);
    gfMembFlux.acceptTimeStep();
    const ax1gpu::NewtonResult& res = pdesolver.result();
    std::printf("newton its %u lin its %u first_defect %.6e defect %.6e backend_lin_it %u backend_red %.3e\n", res.iterations,
                res.linear_solver_iterations, res.first_defect, res.defect, lin_it, lin_red);
    FILE* o = std::fopen(argv[2], "wb");
    std::fwrite(u.data(), sizeof(double), u.size(), o);
    std::fclose(o);
  } catch (const ax1gpu::NewtonError& e) {
    std::fprintf(stderr, "NewtonError: %s\n", e.what());
    return 4;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 5;
  }
  return 0;
}
