"""Bending energy with tilt-splay coupling (single tilt field) on the B200 path.

Twin of ``modules/energy/bending_tilt.py:151-482``:
``E = 1/2 sum_f sum_k kappa_k (2H_k - c0_k + div_f t)^2 va_eff,k`` with the P1 divergence of
``geometry/tilt_operators.py:158-175``.  The shape gradient treats ``div t`` as constant
(``bending_tilt.py:14-19``) -- the bending back-propagation with ``term = base + div_eff`` -- and
the tilt gradient is exact.  ``grad_arr=None`` requests the tilt-only evaluation of the inner tilt
solve (``evaluation_manager.py:693-698``).  The leaflet variants (``bending_tilt_leaflet.py``) live in
``bending_tilt_in.py`` / ``bending_tilt_out.py``.
"""

from __future__ import annotations

import numpy as np

from . import _common as C
from .tilt import _tilts

B200_MODULE = C.L.MOD_BENDING_TILT
USES_TILT = True


def _flags(global_params) -> int:
    mode = C.gradient_mode(global_params)
    if mode == "finite_difference":
        raise C.L.B200Error("bending_gradient_mode=finite_difference is a debug mode of the reference and is not "
                            "part of the B200 path; use 'analytic' or 'approx'")
    return C.L.FLAG_APPROX if mode == "approx" else 0


def b200_configure(state, mesh, global_params, param_resolver, tilts=None) -> dict:
    # the coupling is defined for Helfrich-like models (bending_tilt.py:202-206)
    kappa, c0 = C.per_vertex_bending_params(mesh, global_params, "helfrich")
    state.set_bending(kappa, c0)
    state.set_tilts(_tilts(mesh, tilts), state._k_tilt or 0.0)
    return {"flags": _flags(global_params)}


def b200_energy(result) -> float:
    return float(result.scalars[C.L.SC_E_BENDING_TILT])


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr,
                                      ctx=None, tilts=None, tilt_grad_arr=None) -> float:
    tri, _ = mesh.triangle_row_cache()
    if tri is None or len(tri) == 0:
        return 0.0
    if tilt_grad_arr is not None:
        tilt_grad_arr = np.asarray(tilt_grad_arr)
        if tilt_grad_arr.shape != (len(mesh.vertex_ids), 3):
            raise ValueError("tilt_grad_arr must have shape (N_vertices, 3)")
    flags = _flags(global_params)
    want_grad = grad_arr is not None
    tmp = C.scratch_like(positions) if want_grad else None
    tmp_t = C.scratch_like(positions) if tilt_grad_arr is not None else None
    pos = C.positions_array(positions)
    st = C.get_state(mesh, pos)
    b200_configure(st, mesh, global_params, param_resolver, tilts)
    opts = st.dm.options(B200_MODULE, flags=flags, want_grad=want_grad, want_tilt_grad=tilt_grad_arr is not None)
    res = st.dm.eval_host(opts, pos, grad=tmp, tilt_grad=tmp_t)
    if want_grad:
        C.accumulate(grad_arr, tmp)
        if (flags & C.L.FLAG_APPROX) and st.boundary is not None:
            grad_arr[st.boundary.astype(bool)] = 0.0  # bending_tilt.py:321-323
    if tilt_grad_arr is not None:
        C.accumulate(tilt_grad_arr, tmp_t)
    return b200_energy(res)


def compute_energy_array(mesh, global_params, param_resolver=None, *, positions, index_map, tilts=None) -> float:
    return compute_energy_and_gradient_array(mesh, global_params, param_resolver, positions=positions,
                                             index_map=index_map, grad_arr=None, tilts=tilts)


def compute_energy_and_gradient(mesh, global_params, param_resolver, *, compute_gradient: bool = True):
    """Legacy dict API: ``(E, shape_grad, tilt_grad)`` (``bending_tilt.py:485-527``)."""
    positions = mesh.positions_view()
    idx = mesh.vertex_index_to_row
    if not compute_gradient:
        return compute_energy_array(mesh, global_params, param_resolver, positions=positions, index_map=idx), {}, {}
    g, tg = np.zeros_like(positions), np.zeros_like(positions)
    e = compute_energy_and_gradient_array(mesh, global_params, param_resolver, positions=positions, index_map=idx,
                                          grad_arr=g, tilt_grad_arr=tg)
    rows = list(enumerate(mesh.vertex_ids))
    return float(e), {int(v): g[r].copy() for r, v in rows}, {int(v): tg[r].copy() for r, v in rows}


__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient"]
