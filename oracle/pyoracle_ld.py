"""ctypes binding of the EXTENDED-PRECISION build of the CPU oracle (oracle/libax1oracle_ld.so).

TEST INFRASTRUCTURE ONLY. The same sources as libax1oracle.so compiled with `double` re-defined to the x87 80-bit
`long double` (oracle/ld_prelude.hpp): every struct field, array and intermediate of the restatement carries a
64-bit mantissa (eps 1.1e-19 instead of 2.2e-16). It answers one question -- when the device and the double-precision
oracle disagree in the worst-conditioned unknown after Newton has converged to the rounding floor of the double
residual, where does the truth lie? (tests/test_precision_floor.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import pyoracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libax1oracle_ld.so")
LD = np.longdouble
c_ld_p = C.POINTER(C.c_longdouble)


def _ld_fields(struct):
    out = []
    for name, t in struct._fields_:
        if t is C.c_double:
            t = C.c_longdouble
        elif hasattr(t, "_type_") and hasattr(t, "_length_") and t._type_ is C.c_double:
            t = C.c_longdouble * t._length_
        out.append((name, t))
    return out


class Config(C.Structure):
    _fields_ = _ld_fields(O.Config)


class NewtonOpts(C.Structure):
    _fields_ = _ld_fields(O.NewtonOpts)


class NewtonResult(C.Structure):
    _fields_ = _ld_fields(O.NewtonResult)


def _convert(src, cls):
    dst = cls()
    for name, t in src._fields_:
        v = getattr(src, name)
        if hasattr(v, "__len__"):
            for i in range(len(v)):
                getattr(dst, name)[i] = v[i]
        else:
            setattr(dst, name, v)
    return dst


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".hpp", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libax1oracle_ld.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.ax1o_last_error.restype = C.c_char_p
        L.ax1o_create.restype = C.c_void_p
        L.ax1o_create.argtypes = [C.POINTER(Config), C.c_int, C.c_int, c_ld_p, c_ld_p, C.c_int, C.c_int, c_ld_p, c_ld_p,
                                  c_ld_p, c_ld_p]
        L.ax1o_destroy.argtypes = [C.c_void_p]
        L.ax1o_num_vertices.argtypes = [C.c_void_p]
        L.ax1o_set_time.argtypes = [C.c_void_p, C.c_longdouble, C.c_longdouble]
        L.ax1o_set_uold.argtypes = [C.c_void_p, c_ld_p]
        L.ax1o_update_state.argtypes = [C.c_void_p, c_ld_p]
        L.ax1o_accept_state.argtypes = [C.c_void_p]
        L.ax1o_membrane_potential.argtypes = [C.c_void_p, c_ld_p, c_ld_p]
        L.ax1o_residual.argtypes = [C.c_void_p, c_ld_p, c_ld_p, c_ld_p]
        L.ax1o_newton.argtypes = [C.c_void_p, c_ld_p, C.POINTER(NewtonOpts), C.POINTER(NewtonResult)]
        assert L.ax1o_sizeof_config() == C.sizeof(Config), "long double ax1o_config layout mismatch"
        _lib = L
    return _lib


def ld(a):
    return np.ascontiguousarray(a, dtype=LD)


def _p(a):
    return a.ctypes.data_as(c_ld_p) if a is not None else None


class Problem:
    """One rank's local PNP problem in the extended-precision oracle; takes the double-precision Config of
    pyoracle and float64 arrays, returns numpy longdouble arrays."""

    def __init__(self, cfg, x, y, left_pb=False, right_pb=False, eps_memb=None, g_leak=None, g_nav=None, g_kv=None):
        L = lib()
        self.cfg = _convert(cfg, Config)
        self.x, self.y = ld(x), ld(y)
        self.nx, self.ny = len(self.x) - 1, len(self.y) - 1
        arrs = [None if a is None else ld(np.broadcast_to(a, (self.nx,))) for a in (eps_memb, g_leak, g_nav, g_kv)]
        self._keep = arrs
        self.h = L.ax1o_create(C.byref(self.cfg), self.nx, self.ny, _p(self.x), _p(self.y), int(left_pb), int(right_pb),
                               *[_p(a) for a in arrs])
        if not self.h:
            raise RuntimeError(L.ax1o_last_error().decode())
        self.nv = L.ax1o_num_vertices(self.h)
        self.n = 4 * self.nv

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ax1o_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_time(self, t, dt):
        lib().ax1o_set_time(self.h, t, dt)

    def set_uold(self, u):
        lib().ax1o_set_uold(self.h, _p(ld(u)))

    def update_state(self, u):
        if lib().ax1o_update_state(self.h, _p(ld(u))):
            raise RuntimeError(lib().ax1o_last_error().decode())

    def accept_state(self):
        lib().ax1o_accept_state(self.h)

    def membrane_potential(self, u):
        vm = np.zeros(self.nx, dtype=LD)
        lib().ax1o_membrane_potential(self.h, _p(ld(u)), _p(vm))
        return vm

    def residual(self, u):
        r = np.zeros(self.n, dtype=LD)
        if lib().ax1o_residual(self.h, _p(ld(u)), _p(r), None):
            raise RuntimeError(lib().ax1o_last_error().decode())
        return r

    def newton(self, u, opts):
        """opts: a pyoracle.NewtonOpts (double); returns (u as longdouble array, NewtonResult)"""
        u = ld(u).copy()
        o = _convert(opts, NewtonOpts)
        res = NewtonResult()
        if lib().ax1o_newton(self.h, _p(u), C.byref(o), C.byref(res)) < 0:
            raise RuntimeError(lib().ax1o_last_error().decode())
        return u, res
