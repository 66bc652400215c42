"""Tilt smoothness (Dirichlet) energy of the single tilt field on the B200 path.

Twin of ``modules/energy/tilt_smoothness.py:219-300``: ``E = k_s/4 sum_f (c0 |t1-t2|^2 + c1 |t2-t0|^2 + c2 |t0-t1|^2)``
over every facet with ``k_s = tilt_smoothness_rigidity``; exact tilt gradient, no shape gradient (``:22-24``);
``ambient_v1`` transport only.  Evaluated by the leaflet sweeps of ``csrc/ms_leaflet.cuh`` on their third slot
(``MS_LEAFLET_FIELD``), which has its own tilt and tilt-gradient arrays.
"""

from __future__ import annotations

import numpy as np

from . import _common as C

L = C.L
USES_TILT = True


def _evaluate(mesh, global_params, param_resolver, *, positions, tilts, tilt_grad_arr) -> float:
    k_smooth = float((param_resolver.get(None, "tilt_smoothness_rigidity") if param_resolver is not None
                      else C.gp_get(global_params, "tilt_smoothness_rigidity", 0.0)) or 0.0)
    if k_smooth == 0.0:
        return 0.0
    tri, _ = mesh.triangle_row_cache()
    if tri is None or len(tri) == 0:
        return 0.0
    mode = str(C.gp_get(global_params, "tilt_transport_model", "ambient_v1") or "ambient_v1").strip().lower()
    if mode != "ambient_v1":
        raise L.B200Error("tilt_transport_model=connection_v1 is not available on the B200 path")
    t = mesh.tilts_view() if tilts is None else tilts
    t = np.ascontiguousarray(np.asarray(t, dtype=float))
    if t.shape != (len(mesh.vertex_ids), 3):
        raise ValueError("tilts must have shape (N_vertices, 3)")
    if tilt_grad_arr is not None:
        tilt_grad_arr = np.asarray(tilt_grad_arr)
        if tilt_grad_arr.shape != t.shape:
            raise ValueError("tilt_grad_arr must have shape (N_vertices, 3)")
    pos = C.positions_array(positions)
    st = C.get_state(mesh, pos)
    st.set_leaflet("field", L.MOD_TILT_SMOOTHNESS, dict(k_smooth=k_smooth), 1.0)
    st.dm.set_positions(pos)
    st.dm.upload(L.ARR_TILTS_FIELD, t)
    _, _, e = st.dm.eval_leaflet(L.LEAFLET_FIELD, L.MOD_TILT_SMOOTHNESS, want_grad=False,
                                 want_tilt_grad=tilt_grad_arr is not None)
    if tilt_grad_arr is not None:
        C.accumulate(tilt_grad_arr, st.dm.download(L.ARR_TILT_GRAD_FIELD))
    return float(e)


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr,
                                      tilts=None, tilt_grad_arr=None, ctx=None) -> float:
    return _evaluate(mesh, global_params, param_resolver, positions=positions, tilts=tilts, tilt_grad_arr=tilt_grad_arr)


def compute_energy_array(mesh, global_params, param_resolver=None, *, positions, index_map, tilts=None, ctx=None) -> float:
    return _evaluate(mesh, global_params, param_resolver, positions=positions, tilts=tilts, tilt_grad_arr=None)


def compute_energy_and_gradient(mesh, global_params, param_resolver, *, compute_gradient: bool = True):
    """Legacy dict API ``(E, shape_grad[, tilt_grad])`` (``tilt_smoothness.py:184-216``)."""
    positions = mesh.positions_view()
    tg = np.zeros_like(positions) if compute_gradient else None
    e = _evaluate(mesh, global_params, param_resolver, positions=positions, tilts=None, tilt_grad_arr=tg)
    if not compute_gradient:
        return float(e), {}
    return float(e), {}, {int(v): tg[r].copy() for r, v in enumerate(mesh.vertex_ids) if np.any(tg[r])}


__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient"]
