"""Where the truth lies when two double-precision implementations disagree in the potential.

The oracle compiled a second time in extended precision (x87 long double, eps 1.1e-19: oracle/ld_prelude.hpp,
oracle/pyoracle_ld.py) runs the same restatement with 2048 times less rounding error. On the small problem of the
GPU parity tests it shows:
  * concentrations are insensitive: any double run agrees with the extended-precision solution to ~1e-12;
  * the potential carries one weakly determined mode (Neumann on three sides, Dirichlet 10 mm away) that hardly
    registers in the defect norm. A Newton solve stopped by the usual `reduction 1e-8` criterion is still 7e-6 away
    from the solution in the potential -- in EITHER precision -- so two correct double implementations stopped there
    agree with each other only as far as their iterates happen to coincide (the 5e-7 the round-1 tests allowed);
  * one more Newton iteration (defect at the rounding floor of the double residual, ~1) brings the double
    potential to 1e-8 of the truth: that is the floor of double precision for this unknown, and the tolerance the
    GPU tests assert when they drive Newton to the floor (tests/test_gpu_parity.py::test_fields_at_rounding_floor).
CPU only."""
import numpy as np

NX, DX = 8, 100.0e3
KW = dict(abs_limit=1e-14, min_lin_reduction=1e-10, ilu_fill=0, slab_w=4, maxit=8)


def _problems(oracle):
    from oracle import cpu_setup as CS
    from oracle import pyoracle_ld as OL
    P, u0 = CS.make_problem(NX, DX, perturbation=1e-3)
    Q = OL.Problem(P.cfg, P.x, P.y, eps_memb=np.full(NX, 2.0), g_leak=np.full(NX, 0.5), g_nav=np.full(NX, 120.0),
                   g_kv=np.full(NX, 36.0))
    for X in (P, Q):
        X.set_time(0.0, 1.0); X.set_uold(u0); X.update_state(u0)
    return P, Q, u0


def _err(a, b):
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b.reshape(-1, 4)), axis=0)
    e = np.max(np.abs(np.asarray(a, dtype=np.float64) - b).reshape(-1, 4) / scale, axis=0)
    return float(e[:3].max()), float(e[3])


def test_extended_precision_residual_brackets_the_double_rounding_floor(oracle):
    P, Q, u0 = _problems(oracle)
    rd, scale = P.residual(u0, with_abs=True)
    rl = Q.residual(u0)
    diff = np.abs(rd - rl.astype(np.float64))
    # same restatement: the two evaluations differ by double rounding of the summed terms only
    assert np.all(diff <= 4e-16 * scale * 16)
    # ... which is the floor a double Newton solve can reach: ~0.1 against a defect norm of 7e10
    assert 1e-3 < diff.max() < 2.0 and np.linalg.norm(rd) > 1e10


def test_potential_accuracy_is_set_by_the_newton_stop_not_by_the_arithmetic(oracle):
    P, Q, u0 = _problems(oracle)
    ul, rl = Q.newton(u0, oracle.default_newton_opts(reduction=1e-15, **KW))      # runs into its own floor
    ul2, rl2 = Q.newton(u0, oracle.default_newton_opts(reduction=1e-13, **KW))
    assert float(rl2.defect) < 2e-3                      # extended floor ~ 5e-4 = double floor / 2048
    econ, epot = _err(ul2, ul)
    assert econ < 1e-13 and epot < 5e-9                  # the "truth" is known to 1e-9 in the potential
    # double, stopped by `reduction 1e-8` after 2 iterations: concentrations exact, potential 7e-6 off
    ud, rd = P.newton(u0, oracle.default_newton_opts(reduction=1e-8, **KW))
    econ, epot = _err(ud, ul)
    assert rd.iterations == 2 and econ < 1e-11 and 1e-6 < epot < 1e-4
    # the extended-precision run stopped at the same iteration is just as far off: not a rounding effect
    ul3, rl3 = Q.newton(u0, oracle.default_newton_opts(reduction=1e-8, **KW))
    assert rl3.iterations == 2 and 1e-6 < _err(ul3, ul)[1] < 1e-4 and _err(ud, ul3)[1] < 1e-7
    # double driven to its rounding floor (one more iteration): potential within 2e-8 of the truth
    ud, rd = P.newton(u0, oracle.default_newton_opts(reduction=1e-10, **KW))
    econ, epot = _err(ud, ul)
    assert rd.iterations == 3 and rd.defect < 3.0 and econ < 1e-11 and epot < 2e-8
    # iterating on at the floor neither helps nor hurts
    ud, rd = P.newton(u0, oracle.default_newton_opts(reduction=1e-12, **KW))
    assert rd.status == 1 and _err(ud, ul)[1] < 3e-8
