"""ctypes binding of libax1gpu.so plus the host-side mirror of the reference's interface for the
PNP Newton hot path (same names, argument meaning and error behaviour as the C++ concepts that
Acme2CylSetup::setup instantiates, dune/ax1/acme2_cyl/common/acme2_cyl_setup.hh:226-1018):

    GridOperator.residual / .jacobian          <- IGO            (acme2_cyl_setup.hh:700-708)
    ISTLBackend_BCGS_ILU0.apply / .norm        <- LS             (acme2_cyl_setup.hh:727-811)
    Ax1Newton.apply / .result                  <- PDESOLVER      (acme2_cyl_setup.hh:822-854)
    MembraneFlux.updateState / .acceptTimeStep <- gfMembFlux     (acme2_cyl_simulation.hh:457,921-927)

There is no CPU fallback: every compute entry point raises if libax1gpu.so or a CUDA device is
missing. PyTorch is only used by callers for streams / torch.distributed rendezvous.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libax1gpu.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("x", c_double_p), ("y", c_double_p),
                ("left_proc_boundary", C.c_int), ("right_proc_boundary", C.c_int)]


class Params(C.Structure):
    _fields_ = [
        ("temperature", C.c_double), ("eps_elec", C.c_double), ("length_scale", C.c_double),
        ("time_scale", C.c_double), ("diff_water", C.c_double * 3), ("valence", C.c_int * 3),
        ("pot_residual_scale", C.c_double), ("do_volume_scaling", C.c_int),
        ("do_stimulation", C.c_int), ("tinj_start", C.c_double), ("tinj_end", C.c_double),
        ("stim_x", C.c_double), ("stim_y", C.c_double), ("I_inj", C.c_double),
        ("do_equilibration", C.c_int), ("t_equilibrium", C.c_double), ("dummy_value", C.c_double),
        ("membrane_active", C.c_int), ("automatic_channel_init", C.c_int),
        ("channel_resting_potential", C.c_double), ("use_rownorm_preconditioner", C.c_int),
        ("leak_ratio_na", C.c_double), ("leak_ratio_k", C.c_double),
        ("use_elec_jacobian_volume", C.c_int), ("use_memb_jacobian_volume", C.c_int),
        ("use_time_jacobian_volume", C.c_int),
        ("do_memb_flux_stimulation", C.c_int), ("memb_flux_ion", C.c_int), ("memb_flux_inj", C.c_double),
        ("mori_membrane_neumann", C.c_int),
    ]


class NewtonOpts(C.Structure):
    _fields_ = [
        ("reduction", C.c_double), ("abs_limit", C.c_double), ("min_lin_reduction", C.c_double),
        ("fixed_lin_reduction", C.c_int), ("maxit", C.c_int), ("line_search_maxit", C.c_int),
        ("lin_maxit", C.c_int), ("force_iteration", C.c_int), ("disable_line_search", C.c_int),
        ("slab_w", C.c_int), ("fixed_work", C.c_int), ("ilu_fill", C.c_int), ("amg", C.c_int),
        ("krylov", C.c_int), ("restart", C.c_int),
    ]


class NewtonResult(C.Structure):
    _fields_ = [
        ("converged", C.c_int), ("iterations", C.c_int), ("lin_iterations", C.c_int),
        ("status", C.c_int), ("residual_evals", C.c_int),
        ("first_defect", C.c_double), ("defect", C.c_double), ("reduction", C.c_double),
        ("conv_rate", C.c_double), ("t_assembler", C.c_double), ("t_lin_solv", C.c_double),
        ("t_elapsed", C.c_double),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TimeLoopOpts(C.Structure):
    _fields_ = [("dtstart", C.c_double), ("max_dt", C.c_double), ("min_dt", C.c_double), ("max_dt_ap", C.c_double),
                ("vm_ap_threshold", C.c_double), ("lower_limit", C.c_int), ("upper_limit", C.c_int),
                ("max_restarts", C.c_int), ("adaptive", C.c_int), ("restart_from_uold", C.c_int)]


class TimeLoopState(C.Structure):
    _fields_ = [("time", C.c_double), ("dt", C.c_double), ("its", C.c_int), ("last_its", C.c_int), ("steps", C.c_int)]


class StepResult(C.Structure):
    _fields_ = [("time", C.c_double), ("dt", C.c_double), ("newton_iterations", C.c_int),
                ("lin_iterations", C.c_int), ("restarts", C.c_int), ("newton_status", C.c_int),
                ("vm_min", C.c_double), ("vm_max", C.c_double), ("defect", C.c_double),
                ("first_defect", C.c_double), ("t_assembler", C.c_double), ("t_lin_solv", C.c_double)]

    @property
    def iterations(self):
        return self.newton_iterations


class MoriResult(C.Structure):
    _fields_ = [("status", C.c_int), ("pot_iterations", C.c_int), ("con_iterations", C.c_int),
                ("pot_first_defect", C.c_double), ("pot_defect", C.c_double), ("con_first_defect", C.c_double),
                ("con_defect", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LinResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("reduction", C.c_double),
                ("it_half", C.c_double)]


class MembraneGroupC(C.Structure):
    _fields_ = [("start", C.c_double), ("width", C.c_double), ("stride", C.c_double), ("dx", C.c_double),
                ("smooth_transition", C.c_int), ("geometric_refinement", C.c_int), ("dx_transition", C.c_double),
                ("n_transition", C.c_int), ("permittivity", C.c_double), ("d_memb", C.c_double),
                ("g_leak", C.c_double), ("g_nav", C.c_double), ("g_kv", C.c_double)]


class HostBC(C.Structure):
    _fields_ = [("top", C.c_int * 2), ("bottom", C.c_int * 2), ("left_cytosol", C.c_int * 2),
                ("right_cytosol", C.c_int * 2), ("left_debye", C.c_int * 2), ("right_debye", C.c_int * 2),
                ("left_extra", C.c_int * 2), ("right_extra", C.c_int * 2), ("debye_layer_width", C.c_double)]


# Newton status codes -> names of the Dune::PDELab::NewtonError subclasses the time loop catches
NEWTON_ERRORS = {1: "NewtonNotConverged", 2: "NewtonLinearSolverError", 3: "NewtonDefectError",
                 4: "NewtonLineSearchError"}


class Ax1GpuError(RuntimeError):
    pass


class NewtonError(Ax1GpuError):
    """Dune::PDELab::NewtonError family (acme2_cyl_simulation.hh:593-692 halves dt and retries)."""

    def __init__(self, status, result):
        super().__init__(NEWTON_ERRORS.get(status, "NewtonError(%d)" % status))
        self.status = status
        self.result = result


_lib = None


def lib():
    """Load libax1gpu.so (built in-tree by dune_ax1_b200.build); fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise Ax1GpuError("libax1gpu.so not built: run python -c 'import __graft_entry__ as g; g.build()'")
    L = C.CDLL(_LIBPATH)
    L.ax1gpu_last_error.restype = C.c_char_p
    H = C.c_void_p
    L.ax1gpu_create.argtypes = [C.POINTER(Grid), c_ubyte_p, c_double_p, c_double_p, c_ubyte_p, C.POINTER(Params),
                                c_double_p, c_double_p, c_double_p, C.c_int, C.c_void_p, C.POINTER(H)]
    L.ax1gpu_destroy.argtypes = [H]
    L.ax1gpu_pattern.argtypes = [H, c_int_p, c_int_p, c_int_p]
    L.ax1gpu_set_time.argtypes = [H, C.c_double, C.c_double]
    L.ax1gpu_set_uold.argtypes = [H, c_double_p]
    L.ax1gpu_update_state.argtypes = [H, c_double_p]
    L.ax1gpu_accept_state.argtypes = [H]
    L.ax1gpu_get_channel_state.argtypes = [H, c_double_p, C.c_int, c_double_p]
    L.ax1gpu_set_channel_state.argtypes = [H, c_double_p, C.c_double, C.c_int]
    L.ax1gpu_membrane_potential.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_residual.argtypes = [H, c_double_p, c_double_p, c_double_p]
    L.ax1gpu_jacobian.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_set_jacobian.argtypes = [H, c_double_p]
    L.ax1gpu_get_jacobian.argtypes = [H, c_double_p]
    L.ax1gpu_solve.argtypes = [H, c_double_p, C.c_double, C.c_int, C.c_int, c_double_p, C.POINTER(LinResult)]
    L.ax1gpu_solve_gmres.argtypes = [H, c_double_p, C.c_double, C.c_int, C.c_int, C.c_int, c_double_p,
                                     C.POINTER(LinResult)]
    L.ax1gpu_newton.argtypes = [H, C.POINTER(NewtonOpts), c_double_p, C.POINTER(NewtonResult)]
    L.ax1gpu_upload_u.argtypes = [H, c_double_p]
    L.ax1gpu_download_u.argtypes = [H, c_double_p]
    L.ax1gpu_set_uold_dev.argtypes = [H]
    L.ax1gpu_newton_dev.argtypes = [H, C.POINTER(NewtonOpts), C.POINTER(NewtonResult)]
    L.ax1gpu_timeloop_opts_default.argtypes = [C.POINTER(TimeLoopOpts)]
    L.ax1gpu_timeloop_opts_default.restype = None
    L.ax1gpu_time_step.argtypes = [H, C.POINTER(TimeLoopOpts), C.POINTER(NewtonOpts), C.POINTER(TimeLoopState),
                                   C.POINTER(StepResult)]
    L.ax1gpu_membrane_potential_range.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_membrane_potential_dev.argtypes = [H, c_double_p]
    L.ax1gpu_membrane_flux_terms.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_mori_step.argtypes = [H, C.POINTER(NewtonOpts), c_double_p, C.POINTER(MoriResult)]
    L.ax1gpu_mori_assemble.argtypes = [H, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
    L.ax1gpu_mori_update_state.argtypes = [H, c_double_p]
    L.ax1gpu_mori_get_gamma.argtypes = [H, c_double_p]
    L.ax1gpu_spmv.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_ilu_factor.argtypes = [H, C.c_int]
    L.ax1gpu_ilu_factor_n.argtypes = [H, C.c_int, C.c_int]
    L.ax1gpu_set_amg.argtypes = [H, C.c_int, C.c_int]
    L.ax1gpu_ilu_apply.argtypes = [H, c_double_p, c_double_p]
    L.ax1gpu_set_ilu_split.argtypes = [H, C.c_int]
    L.ax1gpu_get_ilu_split.argtypes = [H, c_int_p]
    L.ax1gpu_set_spmv_compact.argtypes = [H, C.c_int]
    L.ax1gpu_get_spmv_compact.argtypes = [H, c_int_p]
    L.ax1gpu_rownorm.argtypes = [H, c_double_p]
    L.ax1gpu_time_kernel.argtypes = [H, C.c_char_p, C.c_int, c_double_p]
    L.ax1gpu_launch_count.argtypes = [H]
    L.ax1gpu_launch_count.restype = C.c_longlong
    L.ax1gpu_stream.argtypes = [H]
    L.ax1gpu_stream.restype = C.c_void_p
    L.ax1gpu_nccl_unique_id.argtypes = [C.c_char_p, C.c_char_p]
    L.ax1gpu_comm_init.argtypes = [H, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.ax1gpu_comm_mode.argtypes = [H]
    L.ax1gpu_p2p_export.argtypes = [H, C.c_int, C.c_int, C.c_char_p]
    L.ax1gpu_p2p_attach.argtypes = [H, C.c_char_p]
    L.ax1host_generate_y.argtypes = [C.c_double] * 6 + [C.c_int, C.c_double, C.c_double, C.c_int, c_double_p, C.c_int]
    L.ax1host_physics_maps.argtypes = [C.POINTER(Grid), C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_double, C.c_int, C.c_double, C.POINTER(HostBC), c_ubyte_p, c_double_p,
                                       c_ubyte_p]
    L.ax1host_slab.argtypes = [C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]
    GP = C.POINTER(MembraneGroupC)
    L.ax1host_group_partition.argtypes = [GP, C.c_int, C.c_int, C.c_double, c_int_p, c_double_p, c_double_p, C.c_int]
    L.ax1host_generate_x_groups.argtypes = [GP, C.c_int, C.c_int, c_int_p, c_double_p, c_double_p, C.c_double,
                                            C.c_double, c_double_p, C.c_int]
    L.ax1host_membrane_arrays.argtypes = [GP, C.c_int, C.c_int, c_int_p, c_double_p, c_double_p, C.c_int, C.c_double,
                                          C.c_double, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                          c_double_p, c_int_p]
    L.ax1host_check_processor_boundary.argtypes = [GP, C.c_int, c_int_p, c_double_p, c_double_p, C.c_int, C.c_double,
                                                   C.c_double, C.c_double]
    L.ax1host_state_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.ax1host_state_free.argtypes = [C.c_void_p]
    L.ax1host_state_info.argtypes = [C.c_void_p, c_int_p, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p, c_int_p]
    L.ax1host_state_vector.argtypes = [C.c_void_p, C.c_char_p, c_double_p, C.c_int]
    L.ax1host_state_save.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_char_p), C.POINTER(c_double_p), c_int_p]
    L.ax1host_state_extrapolate.argtypes = [C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, C.c_int, c_double_p,
                                            C.c_int, c_double_p, c_double_p]
    assert L.ax1gpu_sizeof_params() == C.sizeof(Params), "ax1gpu_params layout mismatch"
    _lib = L
    return L


def _err():
    return lib().ax1gpu_last_error().decode()


def _chk(rc):
    if rc != 0:
        raise Ax1GpuError("libax1gpu error %d: %s" % (rc, _err()))


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def nccl_library_path():
    """The libnccl.so.2 torch uses (pip package nvidia-nccl), so both share one NCCL instance."""
    try:
        import nvidia.nccl as n  # noqa
        base = list(n.__path__)[0]
        p = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    except Exception:
        pass
    return "libnccl.so.2"


# ------------------------------------------------------------------------------------------------
# host model (no GPU needed)
def generate_y(ymin=0.0, ymax=10.0e6, ymemb=500.0, dmemb=5.0, dy_cell=100.0, dy_cell_min=0.5,
               n_dy_cell_min=10, dy=100.0e3, dy_min=0.5, n_dy_min=10):
    """Radial grid vector; defaults = src/config/square10mm.config of the reference."""
    L = lib()
    n = L.ax1host_generate_y(ymin, ymax, ymemb, dmemb, dy_cell, dy_cell_min, n_dy_cell_min, dy, dy_min, n_dy_min,
                             None, 0)
    if n < 0:
        raise Ax1GpuError(_err())
    out = np.zeros(n)
    L.ax1host_generate_y(ymin, ymax, ymemb, dmemb, dy_cell, dy_cell_min, n_dy_cell_min, dy, dy_min, n_dy_min,
                         _dp(out), n)
    return out


def slab(nx_global, nranks, rank):
    """z-slab partition: (c0, c1) interior cell columns, (l0, l1) local columns incl. overlap."""
    v = [C.c_int() for _ in range(4)]
    _chk(lib().ax1host_slab(nx_global, nranks, rank, *[C.byref(a) for a in v]))
    return tuple(a.value for a in v)


class BoundaryConfig:
    """[boundary.concentration] / [boundary.potential] of the INI file; True = Dirichlet."""

    def __init__(self, top=(True, True), bottom=(False, False), left_cytosol=(False, False),
                 right_cytosol=(False, False), left_debye=(False, False), right_debye=(False, False),
                 left_extra=(False, False), right_extra=(False, False), debye_layer_width=0.0):
        self.c = HostBC()
        for name, val in (("top", top), ("bottom", bottom), ("left_cytosol", left_cytosol),
                          ("right_cytosol", right_cytosol), ("left_debye", left_debye),
                          ("right_debye", right_debye), ("left_extra", left_extra), ("right_extra", right_extra)):
            getattr(self.c, name)[0] = int(val[0])
            getattr(self.c, name)[1] = int(val[1])
        self.c.debye_layer_width = debye_layer_width


def physics_maps(x, y, ymemb=500.0, dmemb=5.0, xmin_global=None, xmax_global=None, ymin=None, ymax=None,
                 do_volume_scaling=True, volume_scaling_threshold=-1.0, bc=None, left_pb=False, right_pb=False):
    """cell_subdomain, node_volume, dirichlet_mask of one rank's local grid (host model of Physics)."""
    x = f64(x)
    y = f64(y)
    nx, ny = len(x) - 1, len(y) - 1
    g = Grid(nx, ny, _dp(x), _dp(y), int(left_pb), int(right_pb))
    bc = bc or BoundaryConfig()
    cs = np.zeros(nx * ny, dtype=np.uint8)
    nvol = np.zeros((nx + 1) * (ny + 1))
    dm = np.zeros((nx + 1) * (ny + 1), dtype=np.uint8)
    _chk(lib().ax1host_physics_maps(
        C.byref(g), ymemb, dmemb, x[0] if xmin_global is None else xmin_global,
        x[-1] if xmax_global is None else xmax_global, y[0] if ymin is None else ymin,
        y[-1] if ymax is None else ymax, int(do_volume_scaling), volume_scaling_threshold, C.byref(bc.c),
        cs.ctypes.data_as(c_ubyte_p), _dp(nvol), dm.ctypes.data_as(c_ubyte_p)))
    return cs, nvol, dm


class MembraneGroups:
    """[membrane.<group>] sections of the INI file (myelin / nodes of Ranvier): x partition, grid vector,
    per-column permittivities and channel densities (host model of Ax1GridGenerator / Physics)."""

    def __init__(self, groups, default_group, xmax, dx_default, smooth_permittivities=True,
                 default_permittivity=2.0, default_dmemb=5.0):
        """groups: ordered dict name -> dict(start, width, stride, dx, smoothTransition, geometricRefinement,
        dx_transition, n_transition, permittivity, d_memb, leak, Nav, Kv) using the INI key names."""
        self.names = list(groups)
        self.xmax, self.dx_default = float(xmax), float(dx_default)
        self.smooth, self.def_perm, self.def_dmemb = int(smooth_permittivities), default_permittivity, default_dmemb
        self.c = (MembraneGroupC * len(self.names))()
        for k, n in enumerate(self.names):
            g = groups[n]
            c = self.c[k]
            c.start, c.width, c.stride = g.get("start", -1.0), g.get("width", -1.0), g.get("stride", -1.0)
            c.dx = g.get("dx", -1.0)
            c.smooth_transition = int(g.get("smoothTransition", False))
            c.geometric_refinement = int(g.get("geometricRefinement", False))
            c.dx_transition, c.n_transition = g.get("dx_transition", -1.0), int(g.get("n_transition", 10))
            c.permittivity, c.d_memb = g.get("permittivity", -1.0), g.get("d_memb", -1.0)
            c.g_leak, c.g_nav, c.g_kv = g.get("leak", 0.0), g.get("Nav", 0.0), g.get("Kv", 0.0)
        self.default = self.names.index(default_group)
        L = lib()
        n = L.ax1host_group_partition(self.c, len(self.names), self.default, self.xmax, None, None, None, 0)
        if n < 0:
            raise Ax1GpuError(_err())
        self.ival_group = np.zeros(n, dtype=np.int32)
        self.ival_start, self.ival_end = np.zeros(n), np.zeros(n)
        L.ax1host_group_partition(self.c, len(self.names), self.default, self.xmax,
                                  self.ival_group.ctypes.data_as(c_int_p), _dp(self.ival_start), _dp(self.ival_end), n)

    def _iv(self):
        return (len(self.ival_group), self.ival_group.ctypes.data_as(c_int_p), _dp(self.ival_start),
                _dp(self.ival_end))

    def x_vector(self):
        L = lib()
        n = L.ax1host_generate_x_groups(self.c, len(self.names), *self._iv(), self.dx_default, self.xmax, None, 0)
        if n < 0:
            raise Ax1GpuError(_err())
        x = np.zeros(n)
        L.ax1host_generate_x_groups(self.c, len(self.names), *self._iv(), self.dx_default, self.xmax, _dp(x), n)
        return x

    def membrane_arrays(self, x):
        """eps_memb, g_leak, g_nav, g_kv, group index of the membrane cell columns of the local grid x."""
        x = f64(x)
        nx = len(x) - 1
        eps, gl, gn, gk = np.zeros(nx), np.zeros(nx), np.zeros(nx), np.zeros(nx)
        grp = np.zeros(nx, dtype=np.int32)
        _chk(lib().ax1host_membrane_arrays(self.c, len(self.names), *self._iv(), self.smooth, self.def_perm,
                                           self.def_dmemb, nx, _dp(x), _dp(eps), _dp(gl), _dp(gn), _dp(gk),
                                           grp.ctypes.data_as(c_int_p)))
        return eps, gl, gn, gk, grp

    def check_processor_boundary(self, xmax_rank, forbidden="node_of_ranvier"):
        fg = self.names.index(forbidden) if forbidden in self.names else -1
        _chk(lib().ax1host_check_processor_boundary(self.c, *self._iv(), fg, self.dx_default, float(xmax_rank),
                                                    self.xmax))


# ------------------------------------------------------------------------------------------------
# simulation-state files (Ax1SimulationState format) and x-extrapolation of a loaded state
def load_state(path, names=("x", "y", "channelStates", "uold", "unew")):
    L = lib()
    h = C.c_void_p()
    _chk(L.ax1host_state_load(path.encode(), C.byref(h)))
    try:
        ne, ts, rk, sz, nvec = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        t, dt = C.c_double(), C.c_double()
        _chk(L.ax1host_state_info(h, C.byref(ne), C.byref(t), C.byref(dt), C.byref(ts), C.byref(rk), C.byref(sz),
                                  C.byref(nvec)))
        out = dict(elements=ne.value, time=t.value, dt=dt.value, timestep=ts.value, mpi_rank=rk.value,
                   mpi_size=sz.value, nvectors=nvec.value)
        for n in names:
            ln = L.ax1host_state_vector(h, n.encode(), None, 0)
            if ln >= 0:
                a = np.zeros(ln)
                L.ax1host_state_vector(h, n.encode(), _dp(a), ln)
                out[n] = a
        return out
    finally:
        L.ax1host_state_free(h)


def save_state(path, nelements, time, dt, vectors, timestep=0, mpi_rank=0, mpi_size=1):
    """vectors: ordered dict name -> array (the reference writes x, y, channelStates, uold, unew)."""
    names = list(vectors)
    arrs = [f64(vectors[n]) for n in names]
    cn = (C.c_char_p * len(names))(*[n.encode() for n in names])
    cd = (c_double_p * len(names))(*[_dp(a) for a in arrs])
    cs = (C.c_int * len(names))(*[len(a) for a in arrs])
    _chk(lib().ax1host_state_save(path.encode(), int(nelements), int(timestep), float(time), float(dt), int(mpi_rank),
                                  int(mpi_size), len(names), cn, cd, cs))


def extrapolate_state(xs, ys, us, x, y):
    """doExtrapolateXData: evaluate the loaded state as a Q1 grid function at the vertices of (x, y)."""
    xs, ys, us, x, y = f64(xs), f64(ys), f64(us), f64(x), f64(y)
    u = np.zeros(4 * len(x) * len(y))
    _chk(lib().ax1host_state_extrapolate(len(xs), _dp(xs), len(ys), _dp(ys), _dp(us), len(x), _dp(x), len(y), _dp(y),
                                         _dp(u)))
    return u


def default_params(**kw):
    p = Params()
    lib().ax1gpu_params_default(C.byref(p))
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            for i, vi in enumerate(v):
                getattr(p, k)[i] = vi
        else:
            setattr(p, k, v)
    return p


def default_newton_opts(**kw):
    o = NewtonOpts()
    lib().ax1gpu_newton_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# ------------------------------------------------------------------------------------------------
class Ax1Gpu:
    """Device-resident PNP problem of one rank (thin wrapper over the C ABI)."""

    def __init__(self, x, y, cell_subdomain, memb_perm, node_volume, dirichlet_mask, params, g_leak, g_nav,
                 g_kv, left_pb=False, right_pb=False, device=0, stream=None):
        L = lib()
        self.x = f64(x)
        self.y = f64(y)
        self.nx, self.ny = len(self.x) - 1, len(self.y) - 1
        self.nv = (self.nx + 1) * (self.ny + 1)
        self.n = 4 * self.nv
        self.params = params
        g = Grid(self.nx, self.ny, _dp(self.x), _dp(self.y), int(left_pb), int(right_pb))
        bcast = lambda a: f64(np.broadcast_to(a, (self.nx,)))  # noqa: E731
        cs = np.ascontiguousarray(cell_subdomain, dtype=np.uint8)
        dm = np.ascontiguousarray(dirichlet_mask, dtype=np.uint8)
        self.h = C.c_void_p()
        rc = L.ax1gpu_create(C.byref(g), cs.ctypes.data_as(c_ubyte_p), _dp(bcast(memb_perm)), _dp(f64(node_volume)),
                             dm.ctypes.data_as(c_ubyte_p), C.byref(params), _dp(bcast(g_leak)), _dp(bcast(g_nav)),
                             _dp(bcast(g_kv)), int(device), C.c_void_p(stream) if stream else None, C.byref(self.h))
        if rc != 0:
            self.h = None
            raise Ax1GpuError("ax1gpu_create failed (%d): %s" % (rc, _err()))

    def close(self):
        if getattr(self, "h", None):
            lib().ax1gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- pattern / time / state
    def pattern(self):
        nb = C.c_int()
        _chk(lib().ax1gpu_pattern(self.h, None, None, C.byref(nb)))
        rowptr = np.zeros(self.nv + 1, dtype=np.int32)
        col = np.zeros(nb.value, dtype=np.int32)
        _chk(lib().ax1gpu_pattern(self.h, rowptr.ctypes.data_as(c_int_p), col.ctypes.data_as(c_int_p), C.byref(nb)))
        return rowptr, col

    def set_time(self, t, dt):
        _chk(lib().ax1gpu_set_time(self.h, float(t), float(dt)))

    def set_uold(self, u):
        _chk(lib().ax1gpu_set_uold(self.h, _dp(f64(u))))

    def update_state(self, u=None):
        _chk(lib().ax1gpu_update_state(self.h, _dp(f64(u)) if u is not None else None))

    def accept_state(self):
        _chk(lib().ax1gpu_accept_state(self.h))

    def channel_state(self, new=False):
        s = np.zeros(5 * self.nx)
        v = C.c_double()
        _chk(lib().ax1gpu_get_channel_state(self.h, _dp(s), int(new), C.byref(v)))
        return s, v.value

    def set_channel_state(self, s, v_rest, initialized=True):
        _chk(lib().ax1gpu_set_channel_state(self.h, _dp(f64(s)), float(v_rest), int(initialized)))

    def membrane_potential(self, u=None):
        vm = np.zeros(self.nx)
        _chk(lib().ax1gpu_membrane_potential(self.h, _dp(f64(u)) if u is not None else None, _dp(vm)))
        return vm

    # -- assembly
    def residual(self, u=None, want_r=True):
        r = np.zeros(self.n) if want_r else None
        nrm = C.c_double()
        _chk(lib().ax1gpu_residual(self.h, _dp(f64(u)) if u is not None else None, _dp(r), C.byref(nrm)))
        return r, nrm.value

    def jacobian(self, u=None, download=True):
        blocks = None
        if download:
            nb = C.c_int()
            _chk(lib().ax1gpu_pattern(self.h, None, None, C.byref(nb)))
            blocks = np.zeros(nb.value * 16)
        _chk(lib().ax1gpu_jacobian(self.h, _dp(f64(u)) if u is not None else None, _dp(blocks)))
        return blocks

    def set_jacobian(self, blocks):
        _chk(lib().ax1gpu_set_jacobian(self.h, _dp(f64(blocks))))

    def rownorm(self, r):
        r = f64(r).copy()
        _chk(lib().ax1gpu_rownorm(self.h, _dp(r)))
        nb = C.c_int()
        _chk(lib().ax1gpu_pattern(self.h, None, None, C.byref(nb)))
        blocks = np.zeros(nb.value * 16)
        _chk(lib().ax1gpu_get_jacobian(self.h, _dp(blocks)))
        return blocks, r

    # -- linear algebra
    def spmv(self, x):
        y = np.zeros(self.n)
        _chk(lib().ax1gpu_spmv(self.h, _dp(f64(x)), _dp(y)))
        return y

    def ilu_factor(self, slab_w=0, fill=-1):
        """SeqILUn(A, n=fill, 1.0) on the device Jacobian; fill < 0 keeps the handle's current level."""
        _chk(lib().ax1gpu_ilu_factor_n(self.h, int(slab_w), int(fill)))

    def set_amg(self, on=True, cx=0):
        """Preconditioner = one V(1,1) cycle of the x-aggregation AMG (block-ILU smoothers) instead of the ILU alone."""
        _chk(lib().ax1gpu_set_amg(self.h, int(on), int(cx)))

    def set_ilu_split(self, jcut=-1):
        """Radial split of the ILU0 sub-slabs: -1 automatic (far-field cut on latency-bound grids), 0 off, > 0 row."""
        _chk(lib().ax1gpu_set_ilu_split(self.h, int(jcut)))

    def set_spmv_compact(self, mode=-1):
        """Solver SpMV operand in 64-byte blocks: -1 automatic (grids of >= 2^20 vertices), 0 never, 1 always."""
        _chk(lib().ax1gpu_set_spmv_compact(self.h, int(mode)))

    def spmv_compact(self):
        """Whether the last factorisation's 64-byte-block copy of the Jacobian is what the solver multiplies with."""
        v = C.c_int()
        _chk(lib().ax1gpu_get_spmv_compact(self.h, C.byref(v)))
        return bool(v.value)

    def ilu_split(self):
        """The radial cut the last factorisation used (0: none)."""
        v = C.c_int()
        _chk(lib().ax1gpu_get_ilu_split(self.h, C.byref(v)))
        return v.value

    def ilu_apply(self, d):
        v = np.zeros(self.n)
        _chk(lib().ax1gpu_ilu_apply(self.h, _dp(f64(d)), _dp(v)))
        return v

    def solve(self, r, reduction, maxit=5000, slab_w=0, fill=-1):
        r = f64(r).copy()
        z = np.zeros(self.n)
        res = LinResult()
        if fill >= 0:  # SeqILUn is built inside the backend's apply: factor with the requested n first
            self.ilu_factor(slab_w, fill)
            slab_w = 0
        _chk(lib().ax1gpu_solve(self.h, _dp(r), float(reduction), int(maxit), int(slab_w), _dp(z), C.byref(res)))
        return z, r, res

    def solve_gmres(self, r, reduction, restart=100, maxit=5000, slab_w=0, fill=-1):
        """Dune::RestartedGMResSolver with the handle's preconditioner (ISTLBackend_GMRES_AMG::apply)."""
        r = f64(r).copy()
        z = np.zeros(self.n)
        res = LinResult()
        if fill >= 0:
            self.ilu_factor(slab_w, fill)
            slab_w = 0
        _chk(lib().ax1gpu_solve_gmres(self.h, _dp(r), float(reduction), int(restart), int(maxit), int(slab_w),
                                      _dp(z), C.byref(res)))
        return z, r, res

    # -- Newton
    def newton(self, u, opts=None):
        u = f64(u).copy()
        opts = opts or default_newton_opts()
        res = NewtonResult()
        _chk(lib().ax1gpu_newton(self.h, C.byref(opts), _dp(u), C.byref(res)))
        return u, res

    def time_step(self, tl_opts, newton_opts, state):
        """One adaptive time step on the device-resident iterate (ax1gpu_time_step); state is updated in place."""
        res = StepResult()
        _chk(lib().ax1gpu_time_step(self.h, C.byref(tl_opts), C.byref(newton_opts), C.byref(state), C.byref(res)))
        return res

    def membrane_flux_terms(self, u=None):
        """[5 terms][3 ions][nx] output terms of the membrane flux (total, diffusion, drift, leak, voltage gated)."""
        out = np.zeros(15 * self.nx)
        _chk(lib().ax1gpu_membrane_flux_terms(self.h, _dp(f64(u)) if u is not None else None, _dp(out)))
        return out.reshape(5, 3, self.nx)

    def membrane_potential_range(self):
        lo, hi = C.c_double(), C.c_double()
        _chk(lib().ax1gpu_membrane_potential_range(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def membrane_potential_dev(self):
        vm = np.zeros(self.nx)
        _chk(lib().ax1gpu_membrane_potential_dev(self.h, _dp(vm)))
        return vm

    # -- electroneutral (Mori) operator-split model (csrc/mori.cu)
    def mori_step(self, u=None, opts=None):
        """One operator-split step (potential solve, charge-layer update, concentration solve). u: host iterate
        (returned updated) or None for the device iterate."""
        opts = opts or default_newton_opts()
        res = MoriResult()
        uu = f64(u).copy() if u is not None else None
        _chk(lib().ax1gpu_mori_step(self.h, C.byref(opts), _dp(uu), C.byref(res)))
        return uu, res

    def mori_assemble(self, which, u, ulast, jacobian=True):
        r = np.zeros(self.n)
        blocks = None
        if jacobian:
            nb = C.c_int()
            _chk(lib().ax1gpu_pattern(self.h, None, None, C.byref(nb)))
            blocks = np.zeros(nb.value * 16)
        _chk(lib().ax1gpu_mori_assemble(self.h, int(which), _dp(f64(u)), _dp(f64(ulast)), _dp(r), _dp(blocks)))
        return r, blocks

    def mori_update_state(self, u=None):
        _chk(lib().ax1gpu_mori_update_state(self.h, _dp(f64(u)) if u is not None else None))

    def mori_gamma(self):
        out = np.zeros(12 * self.nx)
        _chk(lib().ax1gpu_mori_get_gamma(self.h, _dp(out)))
        return out.reshape(2, 2, 3, self.nx)

    def upload_u(self, u):
        _chk(lib().ax1gpu_upload_u(self.h, _dp(f64(u))))

    def download_u(self):
        u = np.zeros(self.n)
        _chk(lib().ax1gpu_download_u(self.h, _dp(u)))
        return u

    def set_uold_dev(self):
        _chk(lib().ax1gpu_set_uold_dev(self.h))

    def newton_dev(self, opts=None):
        opts = opts or default_newton_opts()
        res = NewtonResult()
        _chk(lib().ax1gpu_newton_dev(self.h, C.byref(opts), C.byref(res)))
        return res

    def time_kernel(self, name, reps=10):
        ms = C.c_double()
        _chk(lib().ax1gpu_time_kernel(self.h, name.encode(), int(reps), C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().ax1gpu_launch_count(self.h))

    def stream(self):
        return lib().ax1gpu_stream(self.h)

    def comm_init(self, unique_id, rank, nranks, libnccl=None):
        _chk(lib().ax1gpu_comm_init(self.h, (libnccl or nccl_library_path()).encode(), unique_id, rank, nranks))

    def p2p_export(self, rank, nranks):
        """This rank's peer-window IPC handle (64 bytes); gather over all ranks, then p2p_attach."""
        buf = C.create_string_buffer(64)
        _chk(lib().ax1gpu_p2p_export(self.h, int(rank), int(nranks), buf))
        return buf.raw

    def p2p_attach(self, handles):
        """handles: list of the nranks 64-byte IPC handles in rank order."""
        _chk(lib().ax1gpu_p2p_attach(self.h, b"".join(handles)))

    def comm_mode(self):
        """0 single rank, 1 NCCL send/recv + allreduce, 2 NVLink peer windows (p2p.cuh)."""
        return int(lib().ax1gpu_comm_mode(self.h))


def nccl_unique_id(libnccl=None):
    buf = C.create_string_buffer(128)
    _chk(lib().ax1gpu_nccl_unique_id((libnccl or nccl_library_path()).encode(), buf))
    return buf.raw


# ------------------------------------------------------------------------------------------------
# Mirror of the reference's operator / backend / solver concepts
class GridOperator:
    """IGO: OneStepGridOperator over (spatial GO0, temporal GO1), implicit Euler."""

    def __init__(self, dev):
        self.dev = dev

    def preStep(self, time, dt, uold):
        """OneStepMethod::apply prologue: stage setup, r0 = -M(uold), LOP time = time+dt."""
        self.dev.set_time(time, dt)
        self.dev.set_uold(uold)

    def residual(self, x, r=None):
        out, _ = self.dev.residual(x)
        if r is not None:
            r[:] = out
            return r
        return out

    def jacobian(self, x, A=None):
        blocks = self.dev.jacobian(x)
        if A is not None:
            A[:] = blocks
            return A
        return blocks


class ISTLBackend_BCGS_ILU0:
    """LS: BiCGStab + block ILU0 over x-sub-slabs (ISTLBackend_{SEQ,OVLP}_BCGS_ILUn stand-in)."""

    def __init__(self, dev, maxiter=5000, slab_w=0, verbose=0):
        self.dev, self.maxiter, self.slab_w = dev, maxiter, slab_w
        self.res = LinResult()

    def norm(self, v):
        return float(np.sqrt(np.dot(v, v)))

    def apply(self, A, z, r, reduction):
        if A is not None:
            self.dev.set_jacobian(A)
        zz, rr, self.res = self.dev.solve(r, reduction, self.maxiter, self.slab_w)
        z[:] = zz
        r[:] = rr

    def result(self):
        return self.res


class ISTLBackend_SEQ_BCGS_ILUn(ISTLBackend_BCGS_ILU0):
    """LS of the reference's default build: ISTLBackend_SEQ_BCGS_ILUn(n, w, maxiter, verbose)
    (acme2_cyl_setup.hh:797-798: n = 1, w = 1.0). n = 0 or 1 on the device; w must be 1."""

    def __init__(self, dev, n=1, w=1.0, maxiter=5000, verbose=0, slab_w=0):
        if w != 1.0:
            raise ValueError("relaxation factor w != 1 is not supported")
        super().__init__(dev, maxiter, slab_w, verbose)
        self.n = int(n)

    def apply(self, A, z, r, reduction):
        if A is not None:
            self.dev.set_jacobian(A)
        zz, rr, self.res = self.dev.solve(r, reduction, self.maxiter, self.slab_w, fill=self.n)
        z[:] = zz
        r[:] = rr


class ISTLBackend_GMRES_AMG_ILU0(ISTLBackend_BCGS_ILU0):
    """LS of the myelin build: ISTLBackend_GMRES_AMG_ILU0(gfs, maxiter, verbose, reuse, usesuperlu)
    (ax1_solverbackend.hh:279-302, acme2_cyl_setup.hh:749-760): RestartedGMRes(100) preconditioned by one
    multigrid cycle with ILU0 smoothers. On the device the aggregates are `cx` vertex columns (csrc/amg.cu)."""

    def __init__(self, dev, maxiter=5000, verbose=0, reuse=False, usesuperlu=True, slab_w=0, restart=100, cx=4):
        super().__init__(dev, maxiter, slab_w, verbose)
        self.restart, self.cx = int(restart), int(cx)

    def apply(self, A, z, r, reduction):
        if A is not None:
            self.dev.set_jacobian(A)
        self.dev.set_amg(True, self.cx)
        try:
            zz, rr, self.res = self.dev.solve_gmres(r, reduction, self.restart, self.maxiter, self.slab_w, fill=0)
        finally:
            self.dev.set_amg(False)
        z[:] = zz
        r[:] = rr


class MembraneFlux:
    """gfMembFlux: implicit membrane flux grid function state (channels)."""

    def __init__(self, dev):
        self.dev = dev

    def updateState(self, time=None, dt=None, u=None):
        if time is not None:
            self.dev.set_time(time, dt)
        self.dev.update_state(u)

    def acceptTimeStep(self):
        self.dev.accept_state()


class Ax1Newton:
    """PDESOLVER: damped Newton with the Hackbusch-Reusken line search of Ax1Newton."""

    def __init__(self, dev):
        self.dev = dev
        self.opts = default_newton_opts()
        self.res = NewtonResult()

    def setReduction(self, v): self.opts.reduction = v
    def setAbsoluteLimit(self, v): self.opts.abs_limit = v
    def setMinLinearReduction(self, v): self.opts.min_lin_reduction = v
    def setFixedLinearReduction(self, v): self.opts.fixed_lin_reduction = int(v)
    def setMaxIterations(self, v): self.opts.maxit = int(v)
    def setForceIteration(self, v): self.opts.force_iteration = int(v)
    def setLineSearchMaxIterations(self, v): self.opts.line_search_maxit = int(v)
    def setLinearSolverMaxIterations(self, v): self.opts.lin_maxit = int(v)
    def setSlabWidth(self, v): self.opts.slab_w = int(v)
    def setIluFill(self, n): self.opts.ilu_fill = int(n)
    def setAmg(self, on): self.opts.amg = int(on)
    def setKrylov(self, gmres, restart=100): self.opts.krylov, self.opts.restart = int(gmres), int(restart)

    def apply(self, u):
        unew, self.res = self.dev.newton(u, self.opts)
        u[:] = unew
        if self.res.status != 0:
            raise NewtonError(self.res.status, self.res)
        return u

    def result(self):
        return self.res
