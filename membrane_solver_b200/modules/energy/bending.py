"""Helfrich / Willmore bending plugin on the B200 path.

Twin of ``modules/energy/bending.py:32-181``.  Pass A of the patch kernels computes the
cotan-Laplacian curvature vectors, the mixed-Voronoi and effective vertex areas and the
vertex stage; pass B back-propagates the analytic shape gradient
(``bending_gradient.py:17-175``) or, in ``approx`` mode, applies only ``-L fK``.
``bending_gradient_mode = finite_difference`` is a debug mode of the reference
(``bending_diagnostics.py:12-44``) and is not offered here.
"""

from __future__ import annotations

import numpy as np

from . import _common as C

B200_MODULE = C.L.MOD_BENDING
USES_TILT = False


def _flags(global_params) -> int:
    flags = 0
    if C.energy_model(global_params) == "willmore":
        flags |= C.L.FLAG_WILLMORE
    mode = C.gradient_mode(global_params)
    if mode == "finite_difference":
        raise C.L.B200Error("bending_gradient_mode=finite_difference is a debug mode of the reference and is not "
                            "part of the B200 path; use 'analytic' or 'approx'")
    if mode == "approx":
        flags |= C.L.FLAG_APPROX
    return flags


def b200_configure(state, mesh, global_params, param_resolver) -> dict:
    kappa, c0 = C.per_vertex_bending_params(mesh, global_params, C.energy_model(global_params))
    state.set_bending(kappa, c0)
    return {"flags": _flags(global_params)}


def b200_energy(result) -> float:
    return result.e_bending


def _inactive(mesh, global_params) -> bool:
    kappa, _ = C.per_vertex_bending_params(mesh, global_params, C.energy_model(global_params))
    return C.max_abs(kappa) == 0.0 and float(np.max(kappa)) == 0.0


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr) -> float:
    flags = _flags(global_params)
    want_grad = grad_arr is not None
    tmp = C.scratch_like(positions) if want_grad else None
    st, res = C.device_eval(mesh, positions, B200_MODULE, flags=flags, want_grad=want_grad, grad=tmp,
                            configure=lambda s: b200_configure(s, mesh, global_params, param_resolver))
    if want_grad:
        C.accumulate(grad_arr, tmp)
        if (flags & C.L.FLAG_APPROX) and st.boundary is not None:
            # bending.py:161-165: the approx mode zeroes the boundary rows of the caller's array
            grad_arr[st.boundary.astype(bool)] = 0.0
    return res.e_bending


def compute_energy_array(mesh, global_params, positions, index_map) -> np.ndarray:
    """Per-vertex bending energy, ``(N_vertices,)`` (``bending.py:62-87``)."""
    n = len(mesh.vertex_ids)
    if _inactive(mesh, global_params):
        return np.zeros(n, dtype=float)
    flags = _flags(global_params) & C.L.FLAG_WILLMORE
    st, _ = C.device_eval(mesh, positions, B200_MODULE, flags=flags, want_grad=False, diagnostics=True,
                          configure=lambda s: b200_configure(s, mesh, global_params, None))
    return st.dm.download(C.L.ARR_E_VERTEX)


def compute_total_energy(mesh, global_params, positions, index_map) -> float:
    """``bending.compute_total_energy`` (``bending.py:32-59``)."""
    if _inactive(mesh, global_params):
        return 0.0
    flags = _flags(global_params) & C.L.FLAG_WILLMORE
    _, res = C.device_eval(mesh, positions, B200_MODULE, flags=flags, want_grad=False,
                           configure=lambda s: b200_configure(s, mesh, global_params, None))
    return res.e_bending


compute_energy_and_gradient = C.dict_api(compute_energy_and_gradient_array)

__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_total_energy",
           "compute_energy_and_gradient"]
