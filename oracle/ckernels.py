"""ctypes binding of ``oracle/ref_kernels.c`` (TEST INFRASTRUCTURE).

Exposes the C restatement of the reference's Fortran kernels both as plain
array functions (used by ``ref_modules.use_c_kernels``) and as
``KernelSpec(func, expects_transpose)`` objects with the f2py calling
convention of ``fortran_kernels/loader.py:15-20`` so they can be injected
through the reference's own monkeypatch seam
(``tests/test_surface_nocopy_guardrails.py:49-53``).
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass
from typing import Callable

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libms_oracle.so")
_lib = None

_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ref_kernels.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.oracle_surface_energy_and_gradient.restype = ctypes.c_double
        lib.oracle_surface_energy_and_gradient.argtypes = [
            ctypes.c_int32, ctypes.c_int32, _D, _I, _D, _D, ctypes.c_int32]
        lib.oracle_grad_cotan_batch.restype = None
        lib.oracle_grad_cotan_batch.argtypes = [ctypes.c_int32, _D, _D, _D, _D]
        lib.oracle_apply_beltrami_laplacian.restype = None
        lib.oracle_apply_beltrami_laplacian.argtypes = [
            ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _D, _I, _D, _D, ctypes.c_int32]
        lib.oracle_p1_triangle_divergence.restype = None
        lib.oracle_p1_triangle_divergence.argtypes = [
            ctypes.c_int32, ctypes.c_int32, _D, _D, _I, _D, _D, _D, _D, _D, ctypes.c_int32]
        lib.oracle_compute_curvature_data.restype = None
        lib.oracle_compute_curvature_data.argtypes = [
            ctypes.c_int32, ctypes.c_int32, _D, _I, _D, _D, _D, ctypes.c_int32, _D, _D, _D]
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- plain array API (row-major (n,3)) -----------------------------------
def surface_energy_and_gradient(pos, tri, gamma, grad) -> float:
    lib = _load()
    pos, tri, gamma = _f64(pos), _i32(tri), _f64(gamma)
    assert grad.dtype == np.float64 and grad.flags["C_CONTIGUOUS"]
    return float(lib.oracle_surface_energy_and_gradient(
        pos.shape[0], tri.shape[0], _d(pos), _i(tri), _d(gamma), _d(grad), 1))


def grad_cotan_batch(u, v):
    lib = _load()
    u, v = _f64(u), _f64(v)
    gu, gv = np.empty_like(u), np.empty_like(v)
    lib.oracle_grad_cotan_batch(u.shape[0], _d(u), _d(v), _d(gu), _d(gv))
    return gu, gv


def apply_beltrami_laplacian(weights, tri, field):
    lib = _load()
    weights, tri, field = _f64(weights), _i32(tri), _f64(field)
    out = np.empty_like(field)
    dim = 1 if field.ndim == 1 else field.shape[1]
    lib.oracle_apply_beltrami_laplacian(dim, field.shape[0], tri.shape[0], _d(weights), _i(tri),
                                        _d(field), _d(out), 1)
    return out


def p1_triangle_divergence(pos, tilts, tri):
    lib = _load()
    pos, tilts, tri = _f64(pos), _f64(tilts), _i32(tri)
    nf = tri.shape[0]
    div, area = np.empty(nf), np.empty(nf)
    g0, g1, g2 = np.empty((nf, 3)), np.empty((nf, 3)), np.empty((nf, 3))
    lib.oracle_p1_triangle_divergence(pos.shape[0], nf, _d(pos), _d(tilts), _i(tri), _d(div),
                                      _d(area), _d(g0), _d(g1), _d(g2), 1)
    return div, area, g0, g1, g2


def compute_curvature_data(pos, tri):
    lib = _load()
    pos, tri = _f64(pos), _i32(tri)
    nv, nf = pos.shape[0], tri.shape[0]
    k = np.empty((nv, 3))
    a = np.empty(nv)
    w = np.empty((nf, 3))
    va0, va1, va2 = np.empty(nf), np.empty(nf), np.empty(nf)
    lib.oracle_compute_curvature_data(nv, nf, _d(pos), _i(tri), _d(k), _d(a), _d(w), 1,
                                      _d(va0), _d(va1), _d(va2))
    return k, a, w, va0, va1, va2


# ---- f2py-convention KernelSpec seam (fortran_kernels/loader.py:15-20) ----
@dataclass(frozen=True)
class KernelSpec:
    func: Callable
    expects_transpose: bool


def _as_rows(a_t):
    """(3,n) Fortran-ordered view -> (n,3) C-ordered view without copying."""
    a = a_t.T
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("expected the transpose of a C-contiguous (n,3) array")
    return a


def f2py_surface_kernel() -> KernelSpec:
    def surface_energy_and_gradient_t(pos_t, tri_t, gamma, grad_t, zero_based=1):
        lib = _load()
        pos, tri, grad = _as_rows(pos_t), _as_rows(tri_t), _as_rows(grad_t)
        return float(lib.oracle_surface_energy_and_gradient(
            pos.shape[0], tri.shape[0], _d(pos), _i(tri), _d(gamma), _d(grad), int(zero_based)))

    return KernelSpec(surface_energy_and_gradient_t, True)
