"""Problem set-ups for the CPU oracle built WITHOUT the product library.

TEST INFRASTRUCTURE ONLY (bench.py's cpu_baseline / --impl reference legs and tests/): the reference arm of the
benchmark must not map libax1gpu.so, so the same synthetic / paper problems that dune_ax1_b200/setups.py hands to
the device are restated here on top of oracle/pyoracle.py and numpy alone. tests/test_cpu_setup.py holds the two
builders against each other (grid, partition, initial state, Physics maps).

  x vector     GridVector::equidistant_n_xend (ax1_gridvector.hh:23-31): repeated addition, exact end point
  partition    YaspGrid block partition of the cell columns + one overlap cell column per interior side
               (ax1_gridgenerator.hh:601-631)
  state        the shipped 1-column equilibrium (src/simulation_states/equilibrium/cyl/yasp_square10mm_p0.dat,
               Ax1SimulationState format, ax1_simulationstate.hh:30-57) replicated in x (doExtrapolateXData) and
               perturbed deterministically as SURVEY 8(d) prescribes
"""
import os

import numpy as np

from . import pyoracle as O

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dune_ax1_b200", "data")
STATE_SQUARE10MM = os.path.join(_DATA, "yasp_square10mm_p0.dat")


def parse_state(path):
    """Ax1SimulationState ASCII file -> dict (header scalars, named vectors as float64 arrays)."""
    tok = open(path).read().split()
    i, out = 0, {}
    while i < len(tok):
        key = tok[i]
        if key in ("#elements", "#vectors", "timestep", "mpi_rank", "mpi_size"):
            out[key.strip("#")] = int(float(tok[i + 1])); i += 2
        elif key in ("time", "dt"):
            out[key] = float(tok[i + 1]); i += 2
        else:
            n = int(tok[i + 1])
            out[key] = np.array([float(t) for t in tok[i + 2:i + 2 + n]])
            i += 2 + n
    return out


def equilibrium_column(path=STATE_SQUARE10MM):
    g = parse_state(path)
    y = g["y"]
    return y, g["unew"].reshape(len(y), 2, 4)[:, 0, :].copy(), g["channelStates"]


def equidistant_x(xmin, xmax, n):
    x = [float(xmin)]
    h = (xmax - xmin) / n
    for _ in range(n - 1):
        x.append(x[-1] + h)
    x.append(float(xmax))
    return np.array(x)


def slab(nx_global, nranks, rank):
    """(c0, c1) interior cell columns, (l0, l1) local columns incl. one overlap cell column per interior side."""
    base, rem = divmod(nx_global, nranks)
    a = rank * base + min(rank, rem)
    b = a + base + (1 if rank < rem else 0)
    return a, b, a - (1 if rank > 0 else 0), b + (1 if rank < nranks - 1 else 0)


def initial_state(ucol, nvx, amp=0.0, i_offset=0):
    """replicated equilibrium column times (1 + amp sin(2 pi i/97) cos(2 pi j/31)), vertex-blocked, x fastest"""
    nvy = ucol.shape[0]
    u = np.empty((nvy, nvx, 4))
    u[:, :, :] = ucol[:, None, :]
    if amp:
        i = np.arange(nvx) + i_offset
        j = np.arange(nvy)
        f = 1.0 + amp * np.sin(2 * np.pi * i / 97.0)[None, :] * np.cos(2 * np.pi * j / 31.0)[:, None]
        u *= f[:, :, None]
    return u.reshape(-1)


def make_problem(nx, dx, rank=0, nranks=1, perturbation=0.0, stimulation=False, hh=True, **cfg_over):
    """One rank's oracle problem + its initial state for the synthetic (dx = 100 nm) or paper (dx = 100 um) grid with
    the default boundary set-up (Dirichlet top, Neumann elsewhere) -- the arrays dune_ax1_b200.setups.make_setup
    and tests/helpers.oracle_problem produce, without touching the product library."""
    y, ucol, _ = equilibrium_column()
    xmax = nx * dx
    xg = equidistant_x(0.0, xmax, nx)
    c0, c1, l0, l1 = slab(nx, nranks, rank) if nranks > 1 else (0, nx, 0, nx)
    x = xg[l0:l1 + 1].copy()
    nxl = len(x) - 1
    cfg = O.default_config(xmin_global=0.0, xmax_global=xmax, ymin=float(y[0]), ymax=float(y[-1]),
                           do_stimulation=int(stimulation), tinj_start=20.0, tinj_end=2.0e3, stim_x=150.0e3, stim_y=0.0,
                           I_inj=1.0e13, do_equilibration=0, t_equilibrium=0.0, dummy_value=1.0,
                           analytic_volume_jacobian=7, top_dirichlet=[1, 1], **cfg_over)
    P = O.Problem(cfg, x, y, left_pb=rank > 0, right_pb=rank < nranks - 1, eps_memb=np.full(nxl, 2.0),
                  g_leak=np.full(nxl, 0.5), g_nav=np.full(nxl, 120.0 if hh else 0.0),
                  g_kv=np.full(nxl, 36.0 if hh else 0.0))
    u0 = initial_state(ucol, nxl + 1, perturbation, i_offset=l0)
    return P, u0


class MultiRankProblem:
    """`nranks` overlapping x-slab problems behind the single-problem interface the time loop drives (set_time,
    update_state, set_uold, newton, accept_state, membrane_potential): the reference's MPI run with one thread per
    rank (ax1o_newton_mt). Global vectors are vertex-blocked, x fastest, on the global grid."""

    def __init__(self, nx, dx, nranks, **kw):
        self.nx, self.nranks = nx, nranks
        self.parts = [slab(nx, nranks, r) if nranks > 1 else (0, nx, 0, nx) for r in range(nranks)]
        self.probs, u0s = [], []
        for r in range(nranks):
            P, u0 = make_problem(nx, dx, rank=r, nranks=nranks, **kw)
            self.probs.append(P); u0s.append(u0)
        self.nvy = self.probs[0].ny + 1
        self.cfg = self.probs[0].cfg
        self.u0 = self.gather(u0s)

    def scatter(self, u):
        g = np.asarray(u).reshape(self.nvy, self.nx + 1, 4)
        return [np.ascontiguousarray(g[:, l0:l1 + 1, :]).reshape(-1) for (_, _, l0, l1) in self.parts]

    def gather(self, us):
        g = np.empty((self.nvy, self.nx + 1, 4))
        for (c0, c1, l0, l1), u in zip(self.parts, us):
            loc = u.reshape(self.nvy, l1 - l0 + 1, 4)
            g[:, c0:c1 + 1, :] = loc[:, c0 - l0:c1 - l0 + 1, :]
        return g.reshape(-1)

    def set_time(self, t, dt):
        for P in self.probs:
            P.set_time(t, dt)

    def set_uold(self, u):
        for P, ul in zip(self.probs, self.scatter(u)):
            P.set_uold(ul)

    def update_state(self, u):
        for P, ul in zip(self.probs, self.scatter(u)):
            P.update_state(ul)

    def accept_state(self):
        for P in self.probs:
            P.accept_state()

    def newton(self, u, opts):
        us, res, st = O.newton_mt(self.probs, self.scatter(u), opts)
        return self.gather(us), res[0]

    def membrane_potential(self, u):
        vm = np.empty(self.nx)
        for (c0, c1, l0, l1), P, ul in zip(self.parts, self.probs, self.scatter(u)):
            vm[c0:c1] = P.membrane_potential(ul)[c0 - l0:c1 - l0]
        return vm

    def set_ilu_split(self, jcut):
        for P in self.probs:
            P.set_ilu_split(jcut)
