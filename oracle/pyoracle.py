"""ctypes binding of the CPU oracle (oracle/libax1oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (dune_ax1_b200/) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libax1oracle.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ubyte_p = C.POINTER(C.c_ubyte)


class Config(C.Structure):
    _fields_ = [
        ("temperature", C.c_double), ("eps_elec", C.c_double), ("length_scale", C.c_double),
        ("time_scale", C.c_double), ("diff_water", C.c_double * 3), ("valence", C.c_int * 3),
        ("pot_residual_scale", C.c_double),
        ("ymemb", C.c_double), ("dmemb", C.c_double), ("xmin_global", C.c_double),
        ("xmax_global", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
        ("do_volume_scaling", C.c_int), ("volume_scaling_threshold", C.c_double),
        ("top_dirichlet", C.c_int * 2), ("bottom_dirichlet", C.c_int * 2),
        ("left_cytosol_dirichlet", C.c_int * 2), ("right_cytosol_dirichlet", C.c_int * 2),
        ("left_debye_dirichlet", C.c_int * 2), ("right_debye_dirichlet", C.c_int * 2),
        ("left_extra_dirichlet", C.c_int * 2), ("right_extra_dirichlet", C.c_int * 2),
        ("debye_layer_width", C.c_double),
        ("do_stimulation", C.c_int), ("tinj_start", C.c_double), ("tinj_end", C.c_double),
        ("stim_x", C.c_double), ("stim_y", C.c_double), ("I_inj", C.c_double),
        ("do_equilibration", C.c_int), ("t_equilibrium", C.c_double), ("dummy_value", C.c_double),
        ("membrane_active", C.c_int), ("automatic_channel_init", C.c_int),
        ("channel_resting_potential", C.c_double),
        ("analytic_volume_jacobian", C.c_int), ("use_rownorm_preconditioner", C.c_int),
        ("leak_ratio_na", C.c_double), ("leak_ratio_k", C.c_double),
        ("do_memb_flux_stimulation", C.c_int), ("memb_flux_ion", C.c_int), ("memb_flux_inj", C.c_double),
        ("mori_membrane_neumann", C.c_int),
    ]


class NewtonOpts(C.Structure):
    _fields_ = [
        ("reduction", C.c_double), ("abs_limit", C.c_double), ("min_lin_reduction", C.c_double),
        ("fixed_lin_reduction", C.c_int), ("maxit", C.c_int), ("line_search_maxit", C.c_int),
        ("lin_maxit", C.c_int), ("force_iteration", C.c_int), ("disable_line_search", C.c_int),
        ("ilu_fill", C.c_int), ("slab_w", C.c_int), ("fixed_work", C.c_int),
        ("krylov", C.c_int), ("restart", C.c_int), ("amg", C.c_int), ("amg_cx", C.c_int),
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


class LinResult(C.Structure):
    _fields_ = [("converged", C.c_int), ("iterations", C.c_int), ("reduction", C.c_double),
                ("it_half", C.c_double)]


class MoriResult(C.Structure):
    _fields_ = [("status", C.c_int), ("pot_iterations", C.c_int), ("con_iterations", C.c_int),
                ("pot_first_defect", C.c_double), ("pot_defect", C.c_double), ("con_first_defect", C.c_double),
                ("con_defect", C.c_double)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


HALO_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, c_double_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_int, C.c_double)

_lib = None


_REF_LIB = os.path.join(_HERE, "_ref", "libax1ref.so")


def build_ref(reference="/root/reference"):
    """Compile oracle/_ref/libax1ref.so from the reference's own headers (grid-vector primitives, HH channel
    classes) where they lie under `reference`; returns None when the reference tree is not there (GPU box)."""
    if not os.path.isdir(os.path.join(reference, "dune", "ax1")):
        return _REF_LIB if os.path.exists(_REF_LIB) else None
    subprocess.check_call(["make", "-C", _HERE, "_ref", "REFERENCE=" + reference], stdout=subprocess.DEVNULL)
    return _REF_LIB


def ref_operator_lib():
    """ctypes handle of oracle/_ref/libax1refop.so (the reference's local operators on single cells), or None."""
    path = os.path.join(_HERE, "_ref", "libax1refop.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    R.ax1ref_local_operator.argtypes = [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p, c_double_p,
                                        C.c_int, c_double_p, c_double_p]
    R.ax1ref_rownorm.argtypes = [C.c_int, c_int_p, c_int_p, c_double_p, c_double_p]
    R.ax1ref_membrane_flux_terms.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double, c_double_p]
    R.ax1ref_elec_alpha_boundary.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p, c_int_p, c_double_p, C.c_int,
                                             c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_double,
                                             c_double_p]
    return R


def ref_mori_lib():
    """ctypes handle of oracle/_ref/libax1refmori.so (the reference's Mori operators / parameter classes / charge
    layer on single cells), or None."""
    path = os.path.join(_HERE, "_ref", "libax1refmori.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    R.ax1ref_mori_gamma.argtypes = [c_double_p, C.c_int, C.c_double, c_int_p, c_double_p]
    R.ax1ref_mori_operator.argtypes = [C.c_int, C.c_int, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p,
                                       c_double_p, c_int_p, c_double_p, C.c_int, c_double_p, c_double_p, c_double_p,
                                       c_double_p, c_double_p, c_double_p]
    return R


DEFECT_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.c_int)


def ref_newton_lib():
    """ctypes handle of oracle/_ref/libax1refnewton.so (the reference's Ax1Newton overrides), or None."""
    path = os.path.join(_HERE, "_ref", "libax1refnewton.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    R.ax1ref_line_search.argtypes = [C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_double,
                                     C.c_int, C.c_int, DEFECT_CB, C.c_void_p, c_double_p, C.c_int, c_double_p]
    R.ax1ref_prepare_step.argtypes = [C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, C.c_int]
    return R


def ref_lib():
    """ctypes handle of oracle/_ref/libax1ref.so, or None if it was not built."""
    if not os.path.exists(_REF_LIB):
        return None
    R = C.CDLL(_REF_LIB)
    R.ax1ref_gridvector.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, c_double_p, C.c_int]
    R.ax1ref_cyl_geometry.argtypes = [C.c_int] + [C.c_double] * 6 + [c_double_p]
    R.ax1ref_cyl_geometry.restype = None
    R.ax1ref_electrolyte.argtypes = [C.c_double] * 4 + [c_double_p]
    R.ax1ref_electrolyte.restype = None
    R.ax1ref_channelset.argtypes = [C.c_double] * 7 + [c_double_p, C.c_int, c_double_p]
    return R


def build(force=False):
    """Compile oracle/libax1oracle.so (g++ only; building the checker is not using it)."""
    if force or not os.path.exists(_LIB) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB)
            for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libax1oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        build()
    L = C.CDLL(_LIB)
    L.ax1o_last_error.restype = C.c_char_p
    L.ax1o_create.restype = C.c_void_p
    L.ax1o_create.argtypes = [C.POINTER(Config), C.c_int, C.c_int, c_double_p, c_double_p, C.c_int,
                              C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
    L.ax1o_destroy.argtypes = [C.c_void_p]
    L.ax1o_membrane_flux_terms.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.ax1o_line_search.argtypes = [C.c_int, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, DEFECT_CB,
                                   C.c_void_p, c_double_p, C.c_int, c_int_p]
    L.ax1o_line_search.restype = C.c_double
    L.ax1o_gridvector.argtypes = [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double, c_double_p, C.c_int]
    L.ax1o_generate_y.argtypes = [C.c_double] * 6 + [C.c_int, C.c_double, C.c_double, C.c_int,
                                                      c_double_p, C.c_int]
    for name in ("ax1o_num_vertices", "ax1o_membrane_row"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.ax1o_set_ilu_split.argtypes = [C.c_void_p, C.c_int]
    L.ax1o_set_ilu_split.restype = None
    L.ax1o_mori_assemble.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]
    L.ax1o_mori_get_gamma.argtypes = [C.c_void_p, c_double_p]
    L.ax1o_mori_get_gamma.restype = None
    L.ax1o_mori_update_state.argtypes = [C.c_void_p, c_double_p]
    L.ax1o_mori_update_state.restype = None
    L.ax1o_mori_step.argtypes = [C.c_void_p, c_double_p, C.POINTER(NewtonOpts), C.POINTER(MoriResult)]
    L.ax1o_get_maps.argtypes = [C.c_void_p, c_ubyte_p, c_double_p, c_ubyte_p]
    L.ax1o_pattern.argtypes = [C.c_void_p, c_int_p, c_int_p]
    L.ax1o_constants.argtypes = [C.c_void_p, c_double_p]
    L.ax1o_set_time.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.ax1o_set_uold.argtypes = [C.c_void_p, c_double_p]
    L.ax1o_update_state.argtypes = [C.c_void_p, c_double_p]
    L.ax1o_accept_state.argtypes = [C.c_void_p]
    L.ax1o_get_channel_state.argtypes = [C.c_void_p, c_double_p, C.c_int]
    L.ax1o_set_channel_state.argtypes = [C.c_void_p, c_double_p, C.c_double, C.c_int]
    L.ax1o_get_v_rest.argtypes = [C.c_void_p]
    L.ax1o_get_v_rest.restype = C.c_double
    L.ax1o_membrane_potential.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.ax1o_cell_kernel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, C.c_double, c_double_p]
    L.ax1o_cell_jacobian.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_double, c_double_p]
    L.ax1o_residual.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
    L.ax1o_jacobian.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.ax1o_rownorm.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.ax1o_spmv.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
    L.ax1o_ilu_apply.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]
    L.ax1o_ilu_num_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ax1o_solve.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int,
                             C.c_int, C.c_int, C.POINTER(LinResult)]
    L.ax1o_solve_ex.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_double, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(LinResult)]
    L.ax1o_amg_apply.argtypes = [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]
    L.ax1o_newton.argtypes = [C.c_void_p, c_double_p, C.POINTER(NewtonOpts), C.POINTER(NewtonResult)]
    L.ax1o_newton_ext.argtypes = [C.c_void_p, c_double_p, C.POINTER(NewtonOpts),
                                  C.POINTER(NewtonResult), C.c_int, HALO_FN, ALLREDUCE_FN, C.c_void_p]
    L.ax1o_newton_mt.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(c_double_p),
                                 C.POINTER(NewtonOpts), C.POINTER(NewtonResult)]
    L.ax1o_solve_mt.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(c_double_p),
                                C.POINTER(c_double_p), C.POINTER(c_double_p), C.c_double, C.c_int,
                                C.c_int, C.c_int, C.POINTER(LinResult)]
    assert L.ax1o_sizeof_config() == C.sizeof(Config), "ax1o_config layout mismatch"
    _lib = L
    return L


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_config(**kw):
    c = Config()
    lib().ax1o_config_default(C.byref(c))
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            for i, vi in enumerate(v):
                getattr(c, k)[i] = vi
        else:
            setattr(c, k, v)
    return c


def default_newton_opts(**kw):
    o = NewtonOpts()
    lib().ax1o_newton_opts_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def generate_y(ymin=0.0, ymax=10.0e6, ymemb=500.0, dmemb=5.0, dy_cell=100.0, dy_cell_min=0.5,
               n_dy_cell_min=10, dy=100.0e3, dy_min=0.5, n_dy_min=10):
    """Defaults = square10mm.config (src/config/square10mm.config)."""
    L = lib()
    n = L.ax1o_generate_y(ymin, ymax, ymemb, dmemb, dy_cell, dy_cell_min, n_dy_cell_min, dy, dy_min,
                          n_dy_min, None, 0)
    if n < 0:
        raise RuntimeError(L.ax1o_last_error().decode())
    out = np.zeros(n)
    L.ax1o_generate_y(ymin, ymax, ymemb, dmemb, dy_cell, dy_cell_min, n_dy_cell_min, dy, dy_min,
                      n_dy_min, _dp(out), n)
    return out


class Problem:
    """One rank's local PNP problem in the oracle."""

    def __init__(self, cfg, x, y, left_pb=False, right_pb=False, eps_memb=None, g_leak=None,
                 g_nav=None, g_kv=None):
        L = lib()
        self.cfg = cfg
        self.x = f64(x)
        self.y = f64(y)
        self.nx = len(self.x) - 1
        self.ny = len(self.y) - 1
        arrs = [None if a is None else f64(np.broadcast_to(a, (self.nx,)))
                for a in (eps_memb, g_leak, g_nav, g_kv)]
        self._keep = arrs
        self.h = L.ax1o_create(C.byref(cfg), self.nx, self.ny, _dp(self.x), _dp(self.y),
                               int(left_pb), int(right_pb), *[_dp(a) for a in arrs])
        if not self.h:
            raise RuntimeError(L.ax1o_last_error().decode())
        self.nv = L.ax1o_num_vertices(self.h)
        self.n = 4 * self.nv
        self.jm = L.ax1o_membrane_row(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ax1o_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_ilu_split(self, jcut):
        """Factorise rows j < jcut and j >= jcut of every ILU sub-slab separately (0: off)."""
        lib().ax1o_set_ilu_split(self.h, int(jcut))

    # ---- electroneutral (Mori) operator-split model (oracle/mori.cpp)
    def mori_assemble(self, which, u, ulast, jacobian=True):
        """which 0: potential, 1: concentration sub-problem -> (r, BCRS values or None)"""
        r = np.zeros(self.n)
        vals = np.zeros(lib().ax1o_pattern(self.h, None, None) * 16) if jacobian else None
        if lib().ax1o_mori_assemble(self.h, int(which), _dp(f64(u)), _dp(f64(ulast)), _dp(r), _dp(vals)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return r, vals

    def mori_gamma(self):
        """[side (0 cytosol, 1 ES)][old / new][ion][column] charge-layer states"""
        out = np.zeros(12 * self.nx)
        lib().ax1o_mori_get_gamma(self.h, _dp(out))
        return out.reshape(2, 2, 3, self.nx)

    def mori_update_state(self, u):
        lib().ax1o_mori_update_state(self.h, _dp(f64(u)))

    def mori_step(self, u, opts=None):
        u = f64(u).copy()
        opts = opts or default_newton_opts()
        res = MoriResult()
        if lib().ax1o_mori_step(self.h, _dp(u), C.byref(opts), C.byref(res)) < 0:
            raise RuntimeError(lib().ax1o_last_error().decode())
        return u, res

    def maps(self):
        cs = np.zeros(self.nx * self.ny, dtype=np.uint8)
        nvol = np.zeros(self.nv)
        dm = np.zeros(self.nv, dtype=np.uint8)
        lib().ax1o_get_maps(self.h, cs.ctypes.data_as(c_ubyte_p), _dp(nvol), dm.ctypes.data_as(c_ubyte_p))
        return cs, nvol, dm

    def pattern(self):
        nb = lib().ax1o_pattern(self.h, None, None)
        rowptr = np.zeros(self.nv + 1, dtype=np.int32)
        col = np.zeros(nb, dtype=np.int32)
        lib().ax1o_pattern(self.h, rowptr.ctypes.data_as(c_int_p), col.ctypes.data_as(c_int_p))
        return rowptr, col

    def constants(self):
        out = np.zeros(5)
        lib().ax1o_constants(self.h, _dp(out))
        return dict(PC=out[0], kT_e=out[1], D=out[2:5].copy())

    def set_time(self, t, dt):
        lib().ax1o_set_time(self.h, float(t), float(dt))

    def set_uold(self, u):
        lib().ax1o_set_uold(self.h, _dp(f64(u)))

    def update_state(self, u):
        if lib().ax1o_update_state(self.h, _dp(f64(u))):
            raise RuntimeError(lib().ax1o_last_error().decode())

    def accept_state(self):
        lib().ax1o_accept_state(self.h)

    def channel_state(self, new=False):
        s = np.zeros(5 * self.nx)
        lib().ax1o_get_channel_state(self.h, _dp(s), int(new))
        return s

    def set_channel_state(self, s, v_rest, initialized=True):
        lib().ax1o_set_channel_state(self.h, _dp(f64(s)), float(v_rest), int(initialized))

    def v_rest(self):
        return lib().ax1o_get_v_rest(self.h)

    def membrane_potential(self, u):
        vm = np.zeros(self.nx)
        lib().ax1o_membrane_potential(self.h, _dp(f64(u)), _dp(vm))
        return vm

    def cell_kernel(self, i, j, which, xl, u=None, weight=1.0):
        out = np.zeros(16)
        uu = f64(u) if u is not None else None
        if lib().ax1o_cell_kernel(self.h, int(i), int(j), int(which), _dp(f64(xl)), _dp(uu), float(weight), _dp(out)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return out

    def cell_jacobian(self, i, j, which, xl, weight=1.0):
        out = np.zeros(256)
        if lib().ax1o_cell_jacobian(self.h, int(i), int(j), int(which), _dp(f64(xl)), float(weight), _dp(out)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return out.reshape(16, 16)

    def membrane_flux_terms(self, u):
        """[5 terms][3 ions][nx]: total, diffusion term, drift term, leak, voltage gated (see ax1_oracle.h)."""
        out = np.zeros(15 * self.nx)
        if lib().ax1o_membrane_flux_terms(self.h, _dp(f64(u)), _dp(out)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return out.reshape(5, 3, self.nx)

    def residual(self, u, with_abs=False):
        r = np.zeros(self.n)
        a = np.zeros(self.n) if with_abs else None
        if lib().ax1o_residual(self.h, _dp(f64(u)), _dp(r), _dp(a)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return (r, a) if with_abs else r

    def jacobian(self, u):
        nb = lib().ax1o_pattern(self.h, None, None)
        vals = np.zeros(nb * 16)
        if lib().ax1o_jacobian(self.h, _dp(f64(u)), _dp(vals)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return vals

    def rownorm(self, vals, r):
        vals = f64(vals).copy()
        r = f64(r).copy()
        lib().ax1o_rownorm(self.h, _dp(vals), _dp(r))
        return vals, r

    def spmv(self, vals, x):
        y = np.zeros(self.n)
        lib().ax1o_spmv(self.h, _dp(f64(vals)), _dp(f64(x)), _dp(y))
        return y

    def ilu_apply(self, vals, d, ilu_fill=0, slab_w=0):
        v = np.zeros(self.n)
        if lib().ax1o_ilu_apply(self.h, _dp(f64(vals)), ilu_fill, slab_w, _dp(f64(d)), _dp(v)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return v

    def solve(self, vals, r, reduction, maxit=5000, ilu_fill=1, slab_w=0):
        r = f64(r).copy()
        z = np.zeros(self.n)
        res = LinResult()
        if lib().ax1o_solve(self.h, _dp(f64(vals)), _dp(r), _dp(z), reduction, maxit, ilu_fill, slab_w,
                            C.byref(res)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return z, r, res

    def solve_ex(self, vals, r, reduction, maxit=5000, ilu_fill=0, slab_w=0, krylov=0, restart=100, amg=0,
                 amg_cx=4):
        """krylov 0 = BiCGStab, 1 = RestartedGMRes(restart); amg 1 = multigrid cycle as preconditioner."""
        r = f64(r).copy()
        z = np.zeros(self.n)
        res = LinResult()
        if lib().ax1o_solve_ex(self.h, _dp(f64(vals)), _dp(r), _dp(z), reduction, maxit, ilu_fill, slab_w,
                               krylov, restart, amg, amg_cx, C.byref(res)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return z, r, res

    def amg_apply(self, vals, d, ilu_fill=0, slab_w=0, amg_cx=4):
        v = np.zeros(self.n)
        if lib().ax1o_amg_apply(self.h, _dp(f64(vals)), ilu_fill, slab_w, amg_cx, _dp(f64(d)), _dp(v)):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return v

    def newton(self, u, opts=None):
        u = f64(u).copy()
        opts = opts or default_newton_opts()
        res = NewtonResult()
        st = lib().ax1o_newton(self.h, _dp(u), C.byref(opts), C.byref(res))
        if st < 0:
            raise RuntimeError(lib().ax1o_last_error().decode())
        return u, res

    def newton_ext(self, u, opts, rank, halo_add, allreduce):
        u = f64(u).copy()
        res = NewtonResult()
        hcb = HALO_FN(halo_add)
        acb = ALLREDUCE_FN(allreduce)
        st = lib().ax1o_newton_ext(self.h, _dp(u), C.byref(opts), C.byref(res), rank, hcb, acb, None)
        if st < 0:
            raise RuntimeError(lib().ax1o_last_error().decode())
        return u, res


def newton_mt(problems, us, opts):
    """Overlapping multi-rank Newton, one thread per rank (stand-in for MPI ranks)."""
    n = len(problems)
    us = [f64(u).copy() for u in us]
    hs = (C.c_void_p * n)(*[p.h for p in problems])
    ups = (c_double_p * n)(*[_dp(u) for u in us])
    res = (NewtonResult * n)()
    st = lib().ax1o_newton_mt(hs, n, ups, C.byref(opts), res)
    return us, list(res), st
