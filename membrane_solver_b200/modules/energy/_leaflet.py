"""Shared body of the four leaflet plugins (``tilt_in``, ``tilt_out``, ``bending_tilt_in``,
``bending_tilt_out``) on the B200 path.

Numerics: ``csrc/ms_leaflet.cuh`` through ``ms_ctx_set_leaflet`` / ``ms_ctx_eval_leaflet``
(twin of ``modules/energy/tilt_leaflet.py:26-169`` and ``modules/energy/bending_tilt_leaflet.py:231-758``).

Selections: WHICH facets belong to a leaflet, WHICH rows carry a base term and the per-vertex leaflet
parameters are mesh-option bookkeeping of the reference (``leaflet_presence.py``, ``bt_selection.py``,
``bt_params.py``, ``tilt_params.py``, ``tilt_utils.py``), not arithmetic.  They reach the device as plain
masks (``struct ms_leaflet_desc``) from one of three sources:

* a mesh that answers ``leaflet_selection(leaflet)`` with ready-made arrays (``ArrayMesh(leaflets=...)``);
* ``leaflet_selection.py``: this package's own derivation from the vertex options and the global parameters
  (presets, base-term boundary group, assume-J0 rows, region modes, per-vertex moduli) -- the path of
  BASELINE configs[3] and of any ``ArrayMesh(vertex_options=...)``;
* for the reference's experimental rim / shell controls only (``leaflet_selection.needs_reference_helpers``):
  the reference's own helpers, imported under their ``modules.energy.*`` names inside the reference process.

The reference's experimental switches that change the arithmetic are refused loudly (no silent
approximation): recovered / trace-reconstructed divergence, stage-A lanes, inner update modes, the
scaffold stencil, the flat reference base term, ``connection_v1`` transport, non-analytic gradient modes.
"""

from __future__ import annotations

import importlib

import numpy as np

from . import _common as C

L = C.L
SIGN = {"in": -1.0, "out": 1.0}        # bending_tilt_in.py:46, bending_tilt_out.py:46
WHICH = {"in": L.LEAFLET_IN, "out": L.LEAFLET_OUT}
ARR_TILTS = {"in": L.ARR_TILTS_IN, "out": L.ARR_TILTS_OUT}
ARR_TILT_GRAD = {"in": L.ARR_TILT_GRAD_IN, "out": L.ARR_TILT_GRAD_OUT}
ENERGY_SLOT = {L.MOD_BENDING_TILT: 0, L.MOD_TILT: 1, L.MOD_TILT_SMOOTHNESS: 2}   # ms_ctx_eval_leaflet energies3


def _txt(global_params, key, default=""):
    return str(C.gp_get(global_params, key, default) or default).strip().lower()


def refuse_unsupported(global_params, leaflet: str, *, bending_tilt: bool) -> None:
    """Raise for switches whose arithmetic is not on the B200 path (bt_params.py:13-222)."""
    bad = []
    if bending_tilt:
        if _txt(global_params, "theory_parity_lane"):
            bad.append("theory_parity_lane (recovered divergence / stage-A operators)")
        if _txt(global_params, "bending_tilt_in_update_mode", "off") != "off":
            bad.append("bending_tilt_in_update_mode")
        for key in (f"bending_tilt_interface_divergence_mode_{leaflet}", "bending_tilt_out_interface_divergence_mode",
                    "bending_tilt_interface_divergence_mode"):
            if _txt(global_params, key, "p1_triangle") != "p1_triangle":
                bad.append(key)
        if _txt(global_params, "bending_tilt_in_scaffold_shape_stencil_mode", "off") != "off":
            bad.append("bending_tilt_in_scaffold_shape_stencil_mode")
        for key in (f"bending_tilt_base_term_reference_mode_{leaflet}", "bending_tilt_base_term_reference_mode"):
            if _txt(global_params, key, "current_geometry") != "current_geometry":
                bad.append(key)
        if _txt(global_params, "tilt_transport_model", "ambient_v1") != "ambient_v1":
            bad.append("tilt_transport_model")
        if C.gradient_mode(global_params) != "analytic":
            bad.append("bending_gradient_mode (only 'analytic' for the leaflet coupling)")
    if bad:
        raise L.B200Error("leaflet module option(s) not available on the B200 path: " + ", ".join(bad) +
                          "; there is no CPU fallback")


def _ref(name: str):
    try:
        return importlib.import_module(f"modules.energy.{name}")
    except ImportError as exc:  # not inside the reference process and the mesh carries no selections
        raise L.B200Error("the mesh does not provide leaflet_selection() and the reference's selection helper "
                          f"modules.energy.{name} is not importable") from exc


def _selection_from_reference(mesh, global_params, param_resolver, leaflet: str) -> dict:
    """Masks and parameters as the reference derives them from mesh options."""
    tri, _ = mesh.triangle_row_cache()
    tri = np.asarray(tri, dtype=np.int32)
    nv = len(mesh.vertex_ids)
    index_map = mesh.vertex_index_to_row
    presence = _ref("leaflet_presence")
    sel = _ref("bt_selection")
    par = _ref("bt_params")
    tpar = _ref("tilt_params")
    tut = _ref("tilt_utils")
    absent = presence.leaflet_absent_vertex_mask(mesh, global_params, leaflet=leaflet)
    keep = presence.leaflet_present_triangle_mask(mesh, tri, absent_vertex_mask=absent)
    keep = np.ones(len(tri), bool) if np.size(keep) == 0 else np.asarray(keep, bool).copy()
    keep_tilt = keep.copy()
    transition = sel._shared_rim_support_transition_triangle_mask(mesh, global_params, tri, keep_physical_outer_edge=True)
    if transition is not None:
        keep &= ~np.asarray(transition, bool)        # bt_payload.py:131-144 (coupling module only)
    interior = np.asarray(sel._interior_mask_leaflet(mesh, global_params, cache_tag=leaflet, index_map=index_map), bool)
    kappa, c0 = par._per_vertex_params_leaflet(mesh, global_params, model="helfrich",
                                               kappa_key=f"bending_modulus_{leaflet}", cache_tag=leaflet)
    base_zero = np.zeros(nv, bool)
    presets = par._assume_J0_presets(global_params, cache_tag=leaflet)
    if presets:
        rows = sel._collect_preset_rows(mesh, presets=presets, cache_tag=leaflet, index_map=index_map,
                                        radius_max=par._assume_J0_radius_max(global_params, cache_tag=leaflet),
                                        center_xy=par._assume_J0_center_xy(global_params))
        base_zero[np.asarray(rows, dtype=np.int64)] = True
    rows = sel._base_term_region_zero_rows(mesh, global_params, cache_tag=leaflet, index_map=index_map)
    if np.size(rows):
        base_zero[np.asarray(rows, dtype=np.int64)] = True
    out = dict(keep_bt=keep, keep_tilt=keep_tilt, interior=interior, base_zero=base_zero,
               kappa=np.asarray(kappa, float), c0=np.asarray(c0, float))
    if param_resolver is not None:
        k_s = param_resolver.get(None, f"bending_modulus_{leaflet}")        # tilt_smoothness_utils.py:77-84
        if k_s is None:
            k_s = param_resolver.get(None, "bending_modulus")
        out["k_smooth"] = float(k_s or 0.0)
        out["k_tilt"] = float(tpar._resolve_tilt_modulus(param_resolver, leaflet))
        mode = tpar._resolve_tilt_mass_mode(param_resolver, leaflet)
        out["consistent"] = mode == "consistent"
        out["row_weight"] = tut._active_row_weights(mesh, param_resolver, leaflet)
        shell_mode = tut._resolve_shared_rim_outer_shell_mass_mode(param_resolver, leaflet)
        if shell_mode is not None:
            rows_eff = tri[keep_tilt]
            support = tut._shared_rim_outer_support_triangle_mask(mesh, rows_eff, leaflet)
            if support is not None:                  # tilt_leaflet.py:102-111: per-facet mass mode
                per_facet = np.full(len(tri), mode == "consistent")
                per_facet[np.flatnonzero(keep_tilt)[np.asarray(support, bool)]] = shell_mode == "consistent"
                out["facet_consistent"] = per_facet
    return out


def selection(mesh, global_params, param_resolver, leaflet: str) -> dict:
    from . import leaflet_selection as LS

    own = getattr(mesh, "leaflet_selection", None)
    given = getattr(mesh, "leaflets", None)       # ArrayMesh: ready-made arrays per leaflet, possibly none
    if own is not None and (given is None or leaflet in given):
        return own(leaflet)
    if LS.needs_reference_helpers(global_params):
        return _selection_from_reference(mesh, global_params, param_resolver, leaflet)
    key = (int(getattr(mesh, "_vertex_ids_version", 0) or 0), int(getattr(mesh, "_topology_version", 0) or 0),
           int(getattr(mesh, "_facet_loops_version", 0) or 0), int(getattr(mesh, "_version", 0) or 0), leaflet,
           id(global_params))
    cache = getattr(mesh, "_b200_leaflet_selection_cache", None)
    if cache is None:
        cache = {}
        try:
            setattr(mesh, "_b200_leaflet_selection_cache", cache)
        except AttributeError:
            pass
    hit = cache.get(leaflet)
    if hit is not None and hit[0] == key:
        return hit[1]
    spec = LS.leaflet_selection(mesh, global_params, param_resolver, leaflet)
    cache[leaflet] = (key, spec)
    return spec


def _leaflet_tilts(mesh, leaflet: str, tilts) -> np.ndarray:
    if tilts is None:
        tilts = mesh.tilts_in_view() if leaflet == "in" else mesh.tilts_out_view()
    t = np.asarray(tilts, dtype=float)
    if t.shape != (len(mesh.vertex_ids), 3):
        raise ValueError(f"tilts for leaflet '{leaflet}' must have shape (N_vertices, 3)")
    return np.ascontiguousarray(t)


def configure(state, mesh, global_params, param_resolver, leaflet: str, module_bit: int) -> dict | None:
    """Bring the device's description of the leaflet up to date; None = the module contributes nothing."""
    refuse_unsupported(global_params, leaflet, bending_tilt=bool(module_bit & L.MOD_BENDING_TILT))
    if (module_bit & L.MOD_TILT_SMOOTHNESS) and _txt(global_params, "tilt_transport_model", "ambient_v1") != "ambient_v1":
        raise L.B200Error("tilt_transport_model=connection_v1 is not available on the B200 path")
    spec = selection(mesh, global_params, param_resolver, leaflet)
    if (module_bit & L.MOD_TILT) and float(spec.get("k_tilt", 0.0)) == 0.0:
        return None                                   # tilt_leaflet.py:41-43
    if (module_bit & L.MOD_TILT_SMOOTHNESS) and float(spec.get("k_smooth", 0.0)) == 0.0:
        return None                                   # tilt_smoothness_leaflet.py:33-35
    state.set_leaflet(leaflet, module_bit, spec, SIGN[leaflet])
    return spec


def evaluate(mesh, global_params, param_resolver, *, leaflet: str, module_bit: int, positions, grad_arr, tilts,
             tilt_grad_arr) -> float:
    """One leaflet module through the C ABI with host buffers (accumulating contract of the reference)."""
    tri, _ = mesh.triangle_row_cache()
    if tri is None or len(tri) == 0:
        return 0.0
    if tilt_grad_arr is not None:
        tilt_grad_arr = np.asarray(tilt_grad_arr)
        if tilt_grad_arr.shape != (len(mesh.vertex_ids), 3):
            raise ValueError("tilt_grad_arr must have shape (N_vertices, 3)")
    pos = C.positions_array(positions)
    st = C.get_state(mesh, pos)
    if configure(st, mesh, global_params, param_resolver, leaflet, module_bit) is None:
        return 0.0
    t = _leaflet_tilts(mesh, leaflet, tilts)
    st.dm.set_positions(pos)
    st.dm.upload(ARR_TILTS[leaflet], t)
    energies = st.dm.eval_leaflet(WHICH[leaflet], module_bit, want_grad=grad_arr is not None and module_bit != L.MOD_TILT_SMOOTHNESS,
                                  want_tilt_grad=tilt_grad_arr is not None)
    if grad_arr is not None and module_bit != L.MOD_TILT_SMOOTHNESS:     # the smoothness term has no shape gradient
        C.accumulate(grad_arr, st.dm.download(L.ARR_GRAD))
    if tilt_grad_arr is not None:
        C.accumulate(tilt_grad_arr, st.dm.download(ARR_TILT_GRAD[leaflet]))
    return float(energies[ENERGY_SLOT[module_bit]])


def make_module(leaflet: str, module_bit: int):
    """The three contract functions of one leaflet plugin."""

    def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr,
                                          ctx=None, tilts_in=None, tilts_out=None, tilt_in_grad_arr=None,
                                          tilt_out_grad_arr=None) -> float:
        return evaluate(mesh, global_params, param_resolver, leaflet=leaflet, module_bit=module_bit,
                        positions=positions, grad_arr=grad_arr, tilts=tilts_in if leaflet == "in" else tilts_out,
                        tilt_grad_arr=tilt_in_grad_arr if leaflet == "in" else tilt_out_grad_arr)

    def compute_energy_array(mesh, global_params, param_resolver=None, *, positions, index_map, tilts_in=None,
                             tilts_out=None) -> float:
        return evaluate(mesh, global_params, param_resolver, leaflet=leaflet, module_bit=module_bit,
                        positions=positions, grad_arr=None, tilts=tilts_in if leaflet == "in" else tilts_out,
                        tilt_grad_arr=None)

    def compute_energy_and_gradient(mesh, global_params, param_resolver, *, compute_gradient: bool = True):
        """Legacy dict API: ``(E, shape_grad, tilt_grad)`` (tilt_leaflet.py:196-228)."""
        positions = mesh.positions_view()
        idx = mesh.vertex_index_to_row
        if not compute_gradient:
            return compute_energy_array(mesh, global_params, param_resolver, positions=positions, index_map=idx), {}
        g, tg = np.zeros_like(positions), np.zeros_like(positions)
        kw = {"tilt_in_grad_arr": tg} if leaflet == "in" else {"tilt_out_grad_arr": tg}
        e = compute_energy_and_gradient_array(mesh, global_params, param_resolver, positions=positions,
                                              index_map=idx, grad_arr=g, **kw)
        rows = list(enumerate(mesh.vertex_ids))
        return float(e), {int(v): g[r].copy() for r, v in rows}, {int(v): tg[r].copy() for r, v in rows}

    return compute_energy_and_gradient_array, compute_energy_array, compute_energy_and_gradient
