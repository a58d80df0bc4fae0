"""ctypes binding of oracle/_build/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (sphexample_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from sphexample_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sph_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "sphb200.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "_build/liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, vp = C.POINTER(C.c_double), C.c_void_p
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(_abi.Params), C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.orc_destroy.argtypes = [vp]
        L.orc_simulation_loop.argtypes = [vp, C.c_double]
        L.orc_step.argtypes = [vp, C.c_int64, C.c_int]
        L.orc_update_neighbors.restype = C.c_int64
        L.orc_update_neighbors.argtypes = [vp]
        L.orc_pressure.argtypes = [vp, C.c_int]
        L.orc_neighbor_loop.argtypes = [vp, C.c_int]
        L.orc_delta_t.restype = C.c_double
        L.orc_delta_t.argtypes = [vp]
        L.orc_progress_motion.argtypes = [vp, C.c_double]
        L.orc_apply_mdbc.argtypes = [vp]
        L.orc_half_time_step.argtypes = [vp, C.c_double]
        L.orc_full_time_step.argtypes = [vp, C.c_double]
        L.orc_get.restype = C.c_int
        L.orc_get.argtypes = [vp, C.c_char_p, vp]
        L.orc_get_ids.argtypes = [vp, vp]
        L.orc_get_types.argtypes = [vp, vp]
        L.orc_get_cells.argtypes = [vp, vp]
        L.orc_get_cell_list.restype = C.c_int64
        L.orc_get_cell_list.argtypes = [vp, vp, vp]
        L.orc_get_report.argtypes = [vp, C.POINTER(_abi.Report)]
        L.orc_set_time.argtypes = [vp, C.c_double, C.c_int64]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """fp64 CPU restatement of the reference's step functions; one method per reference function."""

    _VEC = ("pos", "vel", "acc", "pos_h", "vel_h", "gradC", "kgrad", "ghost")
    _SCA = ("rho", "press", "drhodt", "rho_h", "divr", "kern")

    def __init__(self, params: _abi.Params, particles, nthreads: int = 1):
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.N, self.D = particles.Position.shape
        self.params = params
        args = [f64(particles.Position), f64(particles.Velocity), f64(particles.Acceleration),
                f64(particles.Density), np.ascontiguousarray(particles.Type, np.uint8),
                np.ascontiguousarray(particles.GroupMarker, np.uint64),
                np.ascontiguousarray(particles.ID, np.int64), f64(particles.GhostPoints),
                f64(particles.GhostNormals)]
        self._h = lib().orc_create(C.byref(params), self.N, *[_ptr(a) for a in args], int(nthreads))
        if not self._h:
            raise RuntimeError("orc_create failed")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    def simulation_loop(self, next_output_time): lib().orc_simulation_loop(self._h, float(next_output_time))
    def step(self, n=1, reset_delta_x=False): lib().orc_step(self._h, int(n), int(bool(reset_delta_x)))
    def update_neighbors(self): return int(lib().orc_update_neighbors(self._h))
    def pressure(self, half=0): lib().orc_pressure(self._h, int(half))
    def neighbor_loop(self, pass_=0): lib().orc_neighbor_loop(self._h, int(pass_))
    def delta_t(self): return float(lib().orc_delta_t(self._h))
    def progress_motion(self, dt2): lib().orc_progress_motion(self._h, float(dt2))
    def apply_mdbc(self): lib().orc_apply_mdbc(self._h)
    def half_time_step(self, dt2): lib().orc_half_time_step(self._h, float(dt2))
    def full_time_step(self, dt): lib().orc_full_time_step(self._h, float(dt))
    def set_time(self, t, it=0): lib().orc_set_time(self._h, float(t), int(it))

    def get(self, name):
        shape = (self.N, self.D) if name in self._VEC else (self.N,)
        out = np.empty(shape, np.float64)
        if lib().orc_get(self._h, name.encode(), _ptr(out)) != 0:
            raise KeyError(name)
        return out

    @property
    def ids(self):
        out = np.empty(self.N, np.int64)
        lib().orc_get_ids(self._h, _ptr(out))
        return out

    @property
    def types(self):
        out = np.empty(self.N, np.uint8)
        lib().orc_get_types(self._h, _ptr(out))
        return out

    @property
    def cells(self):
        out = np.empty((self.N, self.D), np.int64)
        lib().orc_get_cells(self._h, _ptr(out))
        return out

    def cell_list(self):
        n = int(lib().orc_get_cell_list(self._h, None, None))
        cells = np.empty((n, self.D), np.int64)
        start = np.empty(n + 1, np.int64)
        lib().orc_get_cell_list(self._h, _ptr(cells), _ptr(start))
        return cells, start

    def report(self):
        r = _abi.Report()
        lib().orc_get_report(self._h, C.byref(r))
        return r.as_dict()
