"""KernelSpec getters backed by the B200 C ABI.

Drop-in for ``fortran_kernels/loader.py`` (reference ``loader.py:15-20,30,85,139,193,247``):
the same five getters return ``KernelSpec(func, expects_transpose=False)`` whose callables
take the reference's native-layout arrays -- ``(nv,3)`` C-order float64 positions,
``(nf,3)`` int32 triangle rows, in-place ``intent(inout)`` outputs -- and run on the GPU
through the stateless shims of ``include/ms_b200.h``.  A maintainer plugs them in exactly
where the reference's tests inject fake kernels
(``monkeypatch.setattr(surface, "get_surface_energy_kernel", ...)``,
``tests/test_surface_nocopy_guardrails.py:49-53``).

Unlike the reference there is no silent fallback: if the library or the device is missing
the getter raises, and a non-zero return code of a shim raises ``B200Error``.
Strict no-copy semantics (``MEMBRANE_FORTRAN_STRICT_NOCOPY``, ``surface.py:137-155``) are
always on: wrong dtype raises ``TypeError``, wrong layout raises ``ValueError``.
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable

import numpy as np

from . import _lib as L


@dataclass(frozen=True)
class KernelSpec:
    """Resolved kernel callable with metadata (mirrors ``loader.py:15-20``)."""

    func: Callable
    expects_transpose: bool


def _check(a, dtype, name, ndim=None):
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{name} must be a numpy array")
    if a.dtype != dtype:
        raise TypeError(f"{name} must have dtype {np.dtype(dtype).name}, got {a.dtype}")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError(f"{name} must be C-contiguous (native (n,3) layout)")
    if ndim is not None and a.ndim != ndim:
        raise ValueError(f"{name} must be {ndim}-dimensional")
    return a


def _surface_energy_and_gradient(pos, tri, gamma, grad, zero_based=1):
    """``surface_energy_and_gradient`` (``surface_energy.f90:27-99``): grad += dE/dx; returns E."""
    _check(pos, np.float64, "pos", 2)
    _check(tri, np.int32, "tri", 2)
    _check(gamma, np.float64, "gamma", 1)
    _check(grad, np.float64, "grad", 2)
    e = ctypes.c_double(0.0)
    L.check(L.lib().ms_surface_energy_and_gradient(pos.shape[0], tri.shape[0], L.dptr(pos), L.iptr(tri),
                                                   L.dptr(gamma), L.dptr(grad), ctypes.byref(e), int(zero_based)))
    return float(e.value)


def _grad_cotan_batch(u, v, grad_u, grad_v):
    """``grad_cotan_batch`` (``bending_kernels.f90:32-74``)."""
    for a, n in ((u, "u"), (v, "v"), (grad_u, "grad_u"), (grad_v, "grad_v")):
        _check(a, np.float64, n, 2)
    L.check(L.lib().ms_grad_cotan_batch(u.shape[0], L.dptr(u), L.dptr(v), L.dptr(grad_u), L.dptr(grad_v)))


def _apply_beltrami_laplacian(weights, tri, field, out, zero_based=1):
    """``apply_beltrami_laplacian`` (``bending_kernels.f90:87-131``): out is overwritten."""
    _check(weights, np.float64, "weights", 2)
    _check(tri, np.int32, "tri", 2)
    _check(field, np.float64, "field", 2)
    _check(out, np.float64, "out", 2)
    L.check(L.lib().ms_apply_beltrami_laplacian(field.shape[1], field.shape[0], tri.shape[0], L.dptr(weights),
                                                L.iptr(tri), L.dptr(field), L.dptr(out), int(zero_based)))


def _p1_triangle_divergence(pos, tilts, tri, div_tri, area, g0, g1, g2, zero_based=1):
    """``p1_triangle_divergence`` (``tilt_kernels.f90:26-86``)."""
    _check(pos, np.float64, "pos", 2)
    _check(tilts, np.float64, "tilts", 2)
    _check(tri, np.int32, "tri", 2)
    for a, n in ((div_tri, "div_tri"), (area, "area"), (g0, "g0"), (g1, "g1"), (g2, "g2")):
        _check(a, np.float64, n)
    L.check(L.lib().ms_p1_triangle_divergence(pos.shape[0], tri.shape[0], L.dptr(pos), L.dptr(tilts), L.iptr(tri),
                                              L.dptr(div_tri), L.dptr(area), L.dptr(g0), L.dptr(g1), L.dptr(g2),
                                              int(zero_based)))


def _compute_curvature_data(pos, tri, k_vecs, vertex_areas, weights, zero_based=1, va0_out=None, va1_out=None,
                            va2_out=None):
    """``compute_curvature_data`` (``tilt_kernels.f90:88-190``); the corner areas are optional."""
    _check(pos, np.float64, "pos", 2)
    _check(tri, np.int32, "tri", 2)
    _check(k_vecs, np.float64, "k_vecs", 2)
    _check(vertex_areas, np.float64, "vertex_areas", 1)
    _check(weights, np.float64, "weights", 2)
    for a, n in ((va0_out, "va0_out"), (va1_out, "va1_out"), (va2_out, "va2_out")):
        if a is not None:
            _check(a, np.float64, n, 1)
    L.check(L.lib().ms_compute_curvature_data(pos.shape[0], tri.shape[0], L.dptr(pos), L.iptr(tri), L.dptr(k_vecs),
                                              L.dptr(vertex_areas), L.dptr(weights), int(zero_based),
                                              L.dptr(va0_out), L.dptr(va1_out), L.dptr(va2_out)))


def _spec(fn) -> KernelSpec:
    L.lib()  # raises when the library is missing: no silent fallback
    return KernelSpec(func=fn, expects_transpose=False)


def get_surface_energy_kernel() -> KernelSpec:
    return _spec(_surface_energy_and_gradient)


def get_bending_grad_cotan_kernel() -> KernelSpec:
    return _spec(_grad_cotan_batch)


def get_bending_laplacian_kernel() -> KernelSpec:
    return _spec(_apply_beltrami_laplacian)


def get_tilt_divergence_kernel() -> KernelSpec:
    return _spec(_p1_triangle_divergence)


def get_tilt_curvature_kernel() -> KernelSpec:
    return _spec(_compute_curvature_data)
