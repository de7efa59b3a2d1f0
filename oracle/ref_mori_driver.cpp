// oracle/ref_mori_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// The reference's OWN electroneutral (Mori) classes on single cells, compiled FROM /root/reference (never copied)
// against the stand-ins of oracle/ref_stubs and of ref_operator_driver.cpp (one axis-aligned cell, Q1 local function
// space tree, a Physics that answers look-ups from numbers the test supplies):
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_potential_operator.hh       alpha_volume, alpha_boundary
//   dune/ax1/acme2_cyl/operator/acme2_cyl_mori_concentration_operator.hh   alpha_volume, alpha_boundary
//   dune/ax1/acme2_cyl/configurations/mori/mori_poisson_parameters.hh      A, b, bGrad, c, f, bctype, j
//   dune/ax1/acme2_cyl/configurations/mori/mori_nernst_planck_parameters.hh  A, b, c, f, bctype, j
//   dune/ax1/common/charge_layer_mori_implicit.hh                          evaluate, init, timeStep, acceptTimeStep
//   dune/ax1/common/membranefluxgridfunction_implicit.hh                   channel flux inside the Mori flux sum
// Built into oracle/_ref/libax1refmori.so; pins oracle/mori.cpp (tests/test_reference_pins.py).
#define AX1_REF_MORI 1
#include <cstdio>
#include <cstdlib>
#include "ref_operator_driver.cpp"

namespace {

typedef ChargeLayerMoriImplicit<DGFCon, DGFPot, Physics, SolutionContainer> MoriFlux;
typedef MoriPoissonParameters<GV, double, Physics> MParamPot;
typedef MoriNernstPlanckParameters<GV, double, Physics, InitialGF, MembFlux, MoriFlux> MParamNP;
typedef PowerParameters<MParamNP, 3> MParamCon;
typedef Acme2CylMoriPotentialOperator<MParamCon, MParamPot, FEM, FEM, SolutionContainer, false> LopMoriPot;
typedef Acme2CylMoriConcentrationOperator<MParamCon, MParamPot, FEM, FEM, SolutionContainer, false> LopMoriCon;

}  // namespace

extern "C" {

// ChargeLayerMoriImplicit::init / timeStep / acceptTimeStep on one membrane element: gamma[side][ion] after `init`
// with the first concentration pair, then (nsteps > 0) after each timeStep(dt_s) with the following pairs.
//   con = (1 + nsteps) x {cytosol Na K Cl, ES Na K Cl};  out = (1 + nsteps) x {int Na K Cl, ext Na K Cl}
int ax1ref_mori_gamma(const double* con, int nsteps, double dt_s, const int* valence, double* out) {
  try {
    const double cell[4] = {0.0, 1.0, 0.0, 1.0};
    const double phys[7] = {80.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1e-6};
    const double nodevol4[4] = {1.0, 1.0, 1.0, 1.0};
    const double stim[7] = {0, 0, 0, 0, 0, 0, 0};
    Setup S(cell, phys, valence, nodevol4, 0, stim, CYTOSOL);
    S.ph.prm.use_mori = true;
    S.ph.channels.addMembraneElement(0);
    MoriFlux mf(S.dgfCon, S.dgfPot, S.ph, S.sol, 0.0);
    for (int k = 0; k <= nsteps; ++k) {
      Dune::FieldVector<double, 3> cc, ce;
      for (int j = 0; j < 3; ++j) { cc[j] = con[6 * k + j]; ce[j] = con[6 * k + 3 + j]; }
      if (k == 0) mf.init(0, cc, ce);
      else { mf.timeStep(0, dt_s, cc, ce); }
      for (int j = 0; j < 3; ++j) { out[6 * k + j] = mf.gamma_int_new[0][j]; out[6 * k + 3 + j] = mf.gamma_ext_new[0][j]; }
      mf.acceptTimeStep();
    }
    return 0;
  } catch (std::exception&) {
    return 1;
  }
}

// One local kernel of the split operators on one electrolyte cell.
//   which: 0 potential operator, 1 concentration operator;  boundary: 0 alpha_volume, 1 alpha_boundary on the cell's
//   membrane-interface face (from_cytosol: its top edge)
//   xl     unknowns handed to the operator (16 coefficients, comp * 4 + vertex)
//   xnew   what SolutionContainer::getSolutionPotNew / getSolutionConNew hold in this phase (16 coefficients)
//   xold   getSolutionConOld (16 coefficients)
//   memb   {y_opposite, pot_opposite (PotNew), Na K Cl opposite (ConNew), ext_surface, pot_old_this, pot_old_opposite,
//           membrane permittivity}
//   gamma  {old Na K Cl, new Na K Cl} of this side's charge layer;  misc = {temperature, dt, time, dMemb, membrane
//   Neumann (0/1), charge layer active (0/1)};  other arguments as ax1ref_elec_alpha_boundary
int ax1ref_mori_operator(int which, int boundary, const double* cell, int from_cytosol, const double* xl,
                         const double* xnew, const double* xold, const double* phys, const int* valence,
                         const double* nodevol4, int do_volume_scaling, const double* stim, const double* memb,
                         const double* geff, const double* gamma, const double* misc, double* out16) {
  try {
    Setup S(cell, phys, valence, nodevol4, do_volume_scaling, stim, from_cytosol ? CYTOSOL : ES);
    S.ph.electro = Electrolyte<double>(80.0, misc[0], con_mol, 1e-9);
    S.ph.mV = true;
    S.ph.prm.use_mori = misc[5] != 0.0;
    S.ph.prm.memb_neumann = misc[4] != 0.0;
    S.ph.prm.dmemb = misc[3];
    typedef Physics::ChannelSet CS;
    CS& cs = S.ph.channels;
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(Na)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new LeakChannel<double, ChVec>(K)));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new NavChannel<double, ChVec>()));
    cs.addChannel(Dune::shared_ptr<Channel<double, ChVec> >(new KvChannel<double, ChVec>()));
    const int elem = 0;
    cs.addMembraneElement(elem);
    cs.resize();
    for (int k = 0; k < cs.size(); ++k) cs.getChannelPtr(k)->init();
    ChVec ones(1); ones = 1.0;
    for (int k = 0; k < 4; ++k) {
      cs.setConductance(k, elem, geff[k]);
      if (k < 2) cs.setRatio(k, elem, 1.0);
      else for (int g = 0; g < cs.getChannel(k).numGatingParticles(); ++g) cs.getChannelPtr(k)->setGatingParticle(g, ones);
    }
    S.sol.potNew.xl = xnew; S.sol.conNew.xl = xnew; S.sol.conOld.xl = xold; S.sol.potOld.xl = xold;
    S.flux.reset(new MembFlux(S.dgfCon, S.dgfPot, S.ph, S.sol, 0.0));
    S.flux->time = misc[2];
    MoriFlux mf(S.dgfCon, S.dgfPot, S.ph, S.sol, 0.0);
    mf.time = misc[2]; mf.dt = misc[1];
    for (int j = 0; j < 3; ++j) {
      mf.gamma_int_old[0][j] = gamma[j]; mf.gamma_int_new[0][j] = gamma[3 + j];
      mf.gamma_ext_old[0][j] = gamma[j]; mf.gamma_ext_new[0][j] = gamma[3 + j];
    }
    MParamPot paramPot(S.gv, S.ph, phys[4]);
    MParamCon paramCon;
    for (int j = 0; j < 3; ++j)
      paramCon.push_back(Dune::shared_ptr<MParamNP>(new MParamNP(j, S.gv, S.ph, S.igf, *S.flux, mf, 0.0)));
    // the time loop announces the new time level before the solve (acme2_cyl_mori_simulation.hh); j() refuses to
    // evaluate the membrane flux at an "old" time otherwise
    for (int j = 0; j < 3; ++j) paramCon[j]->prepareNextTimeStep(stim[1]);
    Intersection is;
    is.geo.x0 = cell[0]; is.geo.x1 = cell[1]; is.geo.y = from_cytosol ? cell[3] : cell[2];
    is.gii.eta = from_cytosol ? 1.0 : 0.0;
    is.in = &S.e;
    is.ny = from_cytosol ? 1.0 : -1.0;
    is.y_opposite = memb[0]; is.pot_opposite = memb[1];
    for (int j = 0; j < 3; ++j) is.con_opposite[j] = memb[2 + j];
    is.ext_surface = memb[5];
    is.pot_old_this = memb[6]; is.pot_old_opposite = memb[7]; is.memb_perm = memb[8];
    is.index = elem;
    IntersectionGeometry ig{&is};
    ElementGeometry eg{&S.e};
    XView x(xl);
    for (int k = 0; k < 16; ++k) out16[k] = 0.0;
    RView r{out16};
    if (which == 0) {
      LopMoriPot lop(paramCon, paramPot, 4, S.sol);
      lop.setTime(stim[1]);
      if (boundary) lop.alpha_boundary(ig, S.lfs.pot, x, S.lfs.pot, r);
      else lop.alpha_volume(eg, S.lfs.pot, x, S.lfs.pot, r);
    } else {
      LopMoriCon lop(paramCon, paramPot, 4, S.sol);
      lop.setTime(stim[1]);
      if (boundary) lop.alpha_boundary(ig, S.lfs.con, x, S.lfs.con, r);
      else lop.alpha_volume(eg, S.lfs.con, x, S.lfs.con, r);
    }
    return 0;
  } catch (std::exception& e) {
    if (getenv("AX1_REF_DEBUG")) std::fprintf(stderr, "ax1ref_mori_operator: %s\n", e.what());
    return 1;
  }
}

}  // extern "C"
