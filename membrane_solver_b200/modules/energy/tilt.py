"""Tilt-magnitude plugin on the B200 path.

Twin of ``modules/energy/tilt.py:99-219`` (lumped mass): ``E = sum_f k_t/2 (sum_k |t_k|^2 / 3) A_f``,
shape gradient ``coeff_f dA_f/dx`` and tilt gradient ``k_t t_v A_bary(v)``.
"""

from __future__ import annotations

import numpy as np

from . import _common as C

B200_MODULE = C.L.MOD_TILT
USES_TILT = True


def _k_tilt(param_resolver) -> float:
    return float(param_resolver.get(None, "tilt_rigidity") or 0.0)


def _tilts(mesh, tilts) -> np.ndarray:
    if tilts is None:
        tilts = mesh.tilts_view()
    tilts = np.asarray(tilts, dtype=float)
    if tilts.shape != (len(mesh.vertex_ids), 3):
        raise ValueError("tilts must have shape (N_vertices, 3)")
    return tilts


def b200_configure(state, mesh, global_params, param_resolver, tilts=None) -> dict:
    state.set_tilts(_tilts(mesh, tilts), _k_tilt(param_resolver))
    return {}


def b200_energy(result) -> float:
    return result.e_tilt


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr,
                                      tilts=None, tilt_grad_arr=None) -> float:
    if _k_tilt(param_resolver) == 0.0:
        return 0.0
    tri, _ = mesh.triangle_row_cache()
    if tri is None or len(tri) == 0:
        return 0.0
    if tilt_grad_arr is not None:
        tilt_grad_arr = np.asarray(tilt_grad_arr)
        if tilt_grad_arr.shape != (len(mesh.vertex_ids), 3):
            raise ValueError("tilt_grad_arr must have shape (N_vertices, 3)")
    want_grad = grad_arr is not None or tilt_grad_arr is not None
    tmp = C.scratch_like(positions) if want_grad else None
    tmp_t = C.scratch_like(positions) if tilt_grad_arr is not None else None
    _, res = C.device_eval(mesh, positions, B200_MODULE, want_grad=want_grad, grad=tmp, tilt_grad=tmp_t,
                           configure=lambda s: b200_configure(s, mesh, global_params, param_resolver, tilts))
    if grad_arr is not None:
        C.accumulate(grad_arr, tmp)
    if tilt_grad_arr is not None:
        C.accumulate(tilt_grad_arr, tmp_t)
    return res.e_tilt


def compute_energy_array(mesh, global_params, param_resolver, *, positions, index_map, tilts=None) -> float:
    if _k_tilt(param_resolver) == 0.0:
        return 0.0
    tri, _ = mesh.triangle_row_cache()
    if tri is None or len(tri) == 0:
        return 0.0
    _, res = C.device_eval(mesh, positions, B200_MODULE, want_grad=False,
                           configure=lambda s: b200_configure(s, mesh, global_params, param_resolver, tilts))
    return res.e_tilt


def compute_energy_and_gradient(mesh, global_params, param_resolver, *, compute_gradient: bool = True):
    """Legacy dict API: ``(E, shape_grad, tilt_grad)`` (``tilt.py:28-96``)."""
    positions = mesh.positions_view()
    if not compute_gradient:
        return compute_energy_array(mesh, global_params, param_resolver, positions=positions,
                                    index_map=mesh.vertex_index_to_row), {}, {}
    g = np.zeros_like(positions)
    tg = np.zeros_like(positions)
    e = compute_energy_and_gradient_array(mesh, global_params, param_resolver, positions=positions,
                                          index_map=mesh.vertex_index_to_row, grad_arr=g, tilt_grad_arr=tg)
    rows = list(enumerate(mesh.vertex_ids))
    return float(e), {int(v): g[r].copy() for r, v in rows}, {int(v): tg[r].copy() for r, v in rows}


__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient"]
