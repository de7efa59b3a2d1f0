"""Problem set-ups of the BASELINE configs as plain arrays (grid, Physics maps, channel densities,
initial state), shared by bench.py, the tests and the examples.

  paper      src/config/square10mm.config: x 0..1e7 nm, dx 100e3 (Nx=100), 181-point radial vector
  synthetic  SURVEY 8(d): same radial vector, x equidistant with dx = 100 nm, Nx cell columns,
             HH channels everywhere, deterministic perturbation of the replicated equilibrium column
The equilibrium column is the state file the reference ships
(src/simulation_states/equilibrium/cyl/yasp_square10mm_p0.dat; input data under dune_ax1_b200/data/, read with the
library's own state loader ax1host_state_load, as the reference's loadState does).
"""
import os

import numpy as np

from . import api

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
STATE_SQUARE10MM = os.path.join(_DATA, "yasp_square10mm_p0.dat")
STATE_SQUARE1MM = os.path.join(_DATA, "yasp_square1mm_p0.dat")


def equidistant_x(xmin, xmax, n):
    """GridVector::equidistant_n_xend (ax1_gridvector.hh:23-31): repeated addition, exact end point."""
    x = [float(xmin)]
    h = (xmax - xmin) / n
    for _ in range(n - 1):
        x.append(x[-1] + h)
    x.append(float(xmax))
    return np.array(x)


def load_equilibrium_column(path=STATE_SQUARE10MM):
    """y vector, one vertex column [j][comp] of the stored solution, serialized channel states of a shipped
    1-column equilibrium state (Ax1SimulationState file)."""
    g = api.load_state(path)
    y = g["y"]
    u = g["unew"].reshape(len(y), 2, 4)[:, 0, :].copy()
    return y, u, g["channelStates"]


def replicate_column(ucol, nvx):
    """doExtrapolateXData: copy the 1-column equilibrium to every x (vertex-blocked, x fastest)."""
    nvy = ucol.shape[0]
    u = np.empty((nvy, nvx, 4))
    u[:, :, :] = ucol[:, None, :]
    return u.reshape(-1)


def perturb(u, nvx, nvy, amp=1e-3, i_offset=0):
    """Deterministic synthetic perturbation u (1 + amp sin(2 pi i/97) cos(2 pi j/31)), SURVEY 8(d)."""
    i = np.arange(nvx) + i_offset
    j = np.arange(nvy)
    f = 1.0 + amp * np.sin(2 * np.pi * i / 97.0)[None, :] * np.cos(2 * np.pi * j / 31.0)[:, None]
    return (u.reshape(nvy, nvx, 4) * f[:, :, None]).reshape(-1)


def make_setup(nx, dx=None, xmax=None, stimulation=False, hh=True, equilibration=False, bc=None,
               rank=0, nranks=1, perturbation=0.0, y=None, ucol=None, xg=None, groups=None, **param_over):
    """Arrays of one rank's local problem. Returns a dict usable for api.Ax1Gpu and for the oracle."""
    if y is None or ucol is None:
        y, ucol, _ = load_equilibrium_column()
    if xg is not None:                      # explicit (possibly non-uniform) global x vector
        xg = np.asarray(xg, dtype=np.float64)
        nx, xmax = len(xg) - 1, float(xg[-1])
    else:
        if xmax is None:
            xmax = nx * (dx if dx is not None else 100.0e3)
        xg = equidistant_x(0.0, xmax, nx)
    c0, c1, l0, l1 = api.slab(nx, nranks, rank) if nranks > 1 else (0, nx, 0, nx)
    x = xg[l0:l1 + 1].copy()
    left_pb, right_pb = rank > 0, rank < nranks - 1
    nxl = len(x) - 1
    bc = bc or api.BoundaryConfig()
    cs, nvol, dm = api.physics_maps(x, y, xmin_global=0.0, xmax_global=xmax, bc=bc, left_pb=left_pb,
                                    right_pb=right_pb)
    u0 = replicate_column(ucol, nxl + 1)
    if perturbation:
        u0 = perturb(u0, nxl + 1, len(y), perturbation, i_offset=l0)
    params = dict(do_stimulation=int(stimulation), tinj_start=20.0, tinj_end=2.0e3, stim_x=150.0e3, stim_y=0.0,
                  I_inj=1.0e13, do_equilibration=int(equilibration), t_equilibrium=1.0e9 if equilibration else 0.0,
                  dummy_value=10.0 if equilibration else 1.0)
    params.update(param_over)
    eps_memb = np.full(nxl, 2.0)
    g_leak, g_nav, g_kv = np.full(nxl, 0.5), np.full(nxl, 120.0 if hh else 0.0), np.full(nxl, 36.0 if hh else 0.0)
    if groups:
        # membrane groups [membrane.<name>] of the INI file: (x_start, x_end, permittivity, d_memb, leak, Nav, Kv);
        # a membrane cell belongs to the group that contains its centre; effective permittivity =
        # permittivity * d_memb_default / d_memb (acme2_cyl_physics.hh:2592-2600)
        xc = 0.5 * (x[:-1] + x[1:])
        for (xs, xe, perm, dm_g, gl, gn, gk) in groups:
            m = (xc >= xs) & (xc < xe)
            eps_memb[m] = perm * (5.0 / dm_g)
            g_leak[m], g_nav[m], g_kv[m] = gl, gn, gk
    return dict(x=x, y=y, nx=nxl, ny=len(y) - 1, xmax=xmax, cell_subdomain=cs, node_volume=nvol, dirichlet=dm,
                eps_memb=eps_memb, g_leak=g_leak, g_nav=g_nav, g_kv=g_kv, u0=u0, params=params, left_pb=left_pb,
                right_pb=right_pb, bc=bc, slab=(c0, c1, l0, l1))


STATE_MYELIN = os.path.join(_DATA, "yasp_square10mm_myelin_p0.dat")


def myelin_groups(xmax=48.0e6, node_stride=1000.0e3, node_start=500.0e3, node_width=1.0e3, dx=10.0e3):
    """[membrane.*] sections of src/config/acme2_cyl_par_fiona_myelin_48mmx10mm_48nodes_...config
    (docker-compose-myelin.yml): nodes of Ranvier every `node_stride`, myelin (default group) in between."""
    groups = {
        "node_of_ranvier": dict(start=node_start, width=node_width, stride=node_stride, permittivity=2.0, dx=100.0,
                                smoothTransition=False, leak=0.5, Nav=1200.0, Kv=360.0),
        "myelin": dict(permittivity=6.0, d_memb=500.0, smoothTransition=True, geometricRefinement=True,
                       dx_transition=100.0, n_transition=10, leak=0.0, Nav=0.0, Kv=0.0),
    }
    return api.MembraneGroups(groups, "myelin", xmax, dx, smooth_permittivities=True)


def make_myelin_setup(xmax=48.0e6, rank=0, nranks=1, stimulation=True, check_boundaries=True, **param_over):
    """BASELINE config 5: the myelinated axon (48 mm x 10 mm, 48 nodes of Ranvier by default). Grid vector and
    membrane arrays from the host model of Ax1GridGenerator / Physics; initial state = the two shipped
    equilibrium columns (node / myelin, loadMultipleStates = yes, interpolateMultipleStates = no): a vertex
    column takes the state of the group of the cell column to its right [approximation of the element-wise
    PDELab interpolation of Ax1CoarseToFineGridTransferGridFunction, acme2_cyl_simloader.hh:257-290]."""
    mg = myelin_groups(xmax)
    xg = mg.x_vector()
    nx = len(xg) - 1
    y, ucol_node, _ = load_equilibrium_column()
    y2, ucol_myel, _ = load_equilibrium_column(STATE_MYELIN)
    assert np.array_equal(y, y2)
    c0, c1, l0, l1 = api.slab(nx, nranks, rank) if nranks > 1 else (0, nx, 0, nx)
    if check_boundaries and nranks > 1 and rank < nranks - 1:
        mg.check_processor_boundary(xg[c1])
    x = xg[l0:l1 + 1].copy()
    left_pb, right_pb = rank > 0, rank < nranks - 1
    nxl = len(x) - 1
    bc = api.BoundaryConfig()
    cs, nvol, dm = api.physics_maps(x, y, xmin_global=0.0, xmax_global=xmax, bc=bc, left_pb=left_pb, right_pb=right_pb)
    eps_memb, g_leak, g_nav, g_kv, grp = mg.membrane_arrays(x)
    node = mg.names.index("node_of_ranvier")
    vgrp = np.concatenate([grp, grp[-1:]])                      # vertex column -> group of the cell to its right
    u0 = np.where((vgrp == node)[None, :, None], ucol_node[:, None, :], ucol_myel[:, None, :]).reshape(-1)
    # square10mm-type stimulation at the first node (config: position_x = 500e3, I_inj = 1.1e11, t in (20, 3000) us)
    params = dict(do_stimulation=int(stimulation), tinj_start=20.0, tinj_end=3.0e3, stim_x=500.0e3, stim_y=0.0,
                  I_inj=1.1e11, do_equilibration=0, t_equilibrium=0.0, dummy_value=1.0)
    params.update(param_over)
    return dict(x=x, y=y, nx=nxl, ny=len(y) - 1, xmax=xmax, cell_subdomain=cs, node_volume=nvol, dirichlet=dm,
                eps_memb=eps_memb, g_leak=g_leak, g_nav=g_nav, g_kv=g_kv, u0=u0, params=params, left_pb=left_pb,
                right_pb=right_pb, bc=bc, slab=(c0, c1, l0, l1), xg=xg, groups=mg, cell_group=grp)


def make_device(s, device=0, stream=None):
    p = api.default_params(**s["params"])
    return api.Ax1Gpu(s["x"], s["y"], s["cell_subdomain"], s["eps_memb"], s["node_volume"], s["dirichlet"], p,
                      s["g_leak"], s["g_nav"], s["g_kv"], left_pb=s["left_pb"], right_pb=s["right_pb"],
                      device=device, stream=stream)
