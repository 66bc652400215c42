"""Leaflet selections and per-vertex leaflet parameters derived from mesh options -- on arrays.

Which facets belong to a leaflet, which rows carry a base term and which moduli a vertex uses are
bookkeeping over ``vertex.options`` / global parameters in the reference:

* leaflet presence         ``modules/energy/leaflet_presence.py:34-170``  (``leaflet_<l>_absent_presets``)
* interior rows            ``modules/energy/bt_selection.py:289-330``     (boundary vertices and the
                           ``bending_tilt_base_term_boundary_group_<l>`` ring carry no base term)
* assume-J0 rows           ``bt_selection.py:140-201``, ``bt_params.py:40-92``
* base-term region modes   ``bt_selection.py:229-287``
* per-vertex kappa / c0    ``bt_params.py:233-318``
* tilt modulus, mass mode  ``tilt_params.py:6-24``; smoothness rigidity ``tilt_smoothness_utils.py:77-84``

Here the options of all vertices are read ONCE into columns (``VertexOptionTable``: preset labels, group tags,
numeric overrides), cached on the vertex-id version, and every selection is a vectorised expression over those
columns, so that a million-vertex ``ArrayMesh`` pays numpy time, not a Python loop per evaluation.  The same
code serves a mesh with per-vertex option dicts (the reference's ``Mesh``, ``ArrayMesh(vertex_options=...)``)
and a mesh that answers ``vertex_option_columns()`` with ready-made arrays.

The reference's experimental rim modes (``rim_slope_match_mode`` = ``shared_rim_staggered_v1`` /
``physical_edge_staggered_v1``: transition-triangle masks, shell row weights, per-facet mass modes) are not
re-derived here: ``needs_reference_helpers`` reports them and ``_leaflet.py`` then takes the masks from the
reference's own helpers (inside the reference process) or refuses.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# option keys whose value tags a vertex as member of a base-term boundary group (bt_selection.py:22-30)
GROUP_KEYS = ("rim_slope_match_group", "tilt_thetaB_group", "tilt_thetaB_group_in", "tilt_thetaB_group_out")
_NUMERIC_KEYS = ("bending_modulus", "bending_modulus_in", "bending_modulus_out", "spontaneous_curvature",
                 "spontaneous_curvature_in", "spontaneous_curvature_out", "intrinsic_curvature")


@dataclass
class VertexOptionTable:
    """Per-vertex options as columns aligned with ``mesh.vertex_ids``."""

    preset: np.ndarray                                   # (nv,) object: the ``preset`` label ('' = none)
    groups: dict = field(default_factory=dict)           # key -> (nv,) object array of group tags ('' = none)
    numeric: dict = field(default_factory=dict)          # key -> (rows int64, values float64) overrides


def _gp(global_params, key, default=None):
    if global_params is None:
        return default
    val = global_params.get(key)
    return default if val is None else val


def option_table(mesh) -> VertexOptionTable:
    """Columns of the vertex options, cached on ``mesh._vertex_ids_version``."""
    own = getattr(mesh, "vertex_option_columns", None)
    if own is not None:
        return own()
    key = (int(getattr(mesh, "_vertex_ids_version", 0) or 0), len(mesh.vertex_ids))
    cached = getattr(mesh, "_b200_option_table", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    build = getattr(mesh, "build_position_cache", None)
    if build is not None:
        build()
    nv = len(mesh.vertex_ids)
    preset = np.full(nv, "", dtype=object)
    groups = {k: np.full(nv, "", dtype=object) for k in GROUP_KEYS}
    rows = {k: [] for k in _NUMERIC_KEYS}
    vals = {k: [] for k in _NUMERIC_KEYS}
    vertices = getattr(mesh, "vertices", {}) or {}
    for row, vid in enumerate(mesh.vertex_ids):
        v = vertices.get(int(vid)) if hasattr(vertices, "get") else vertices[int(vid)]
        opts = (getattr(v, "options", None) or {}) if v is not None else {}
        if not opts:
            continue
        p = opts.get("preset")
        if p is not None:
            preset[row] = p
        for k in GROUP_KEYS:
            g = opts.get(k)
            if g is not None:
                groups[k][row] = g
        for k in _NUMERIC_KEYS:
            if k in opts and opts[k] is not None:
                try:
                    x = float(opts[k])
                except (TypeError, ValueError):  # the reference skips values it cannot convert (bt_params.py:283-296)
                    continue
                rows[k].append(row)
                vals[k].append(x)
    numeric = {k: (np.asarray(rows[k], dtype=np.int64), np.asarray(vals[k], dtype=np.float64)) for k in _NUMERIC_KEYS}
    table = VertexOptionTable(preset=preset, groups=groups, numeric=numeric)
    try:
        setattr(mesh, "_b200_option_table", (key, table))
    except AttributeError:
        pass
    return table


def _names(raw) -> list[str]:
    """Normalised list of preset names (leaflet_presence.py:16-31, bt_params.py:40-62)."""
    if raw is None:
        return []
    if isinstance(raw, str):
        raw = [raw]
    try:
        items = list(raw)
    except TypeError:
        items = [raw]
    return [s for s in (str(x).strip() for x in items if x is not None) if s]


_SHELL_KEYS = ("tilt_in_exclude_shared_rim_outer_rows", "tilt_out_exclude_shared_rim_outer_rows",
               "tilt_exclude_shared_rim_outer_rows_in", "tilt_exclude_shared_rim_outer_rows_out",
               "tilt_out_exclude_shared_rim_rows", "tilt_exclude_shared_rim_rows_out", "tilt_in_exclude_shared_rim_rows",
               "tilt_exclude_shared_rim_rows_in", "tilt_in_shared_rim_outer_row_energy_weight",
               "tilt_in_shared_rim_outer_shell_mass_mode", "tilt_out_shared_rim_outer_shell_mass_mode",
               "parity_trace_layer_radius")


def needs_reference_helpers(global_params) -> list[str]:
    """Experimental rim / shell controls (tilt_utils.py:28-260, bt_selection.py:103-137, leaflet_presence.py:128-153)
    whose masks and row weights are not derived here."""
    found = []
    mode = str(_gp(global_params, "rim_slope_match_mode", "") or "").strip().lower()
    if mode in ("shared_rim_staggered_v1", "physical_edge_staggered_v1"):
        found.append(f"rim_slope_match_mode={mode}")
    found += [k for k in _SHELL_KEYS if _gp(global_params, k) is not None]
    return found


def absent_vertex_mask(mesh, global_params, leaflet: str) -> np.ndarray:
    """Vertices whose preset lists the leaflet as absent (leaflet_presence.py:34-125)."""
    nv = len(mesh.vertex_ids)
    if leaflet not in ("in", "out"):
        return np.zeros(nv, dtype=bool)
    absent = set(_names(_gp(global_params, f"leaflet_{leaflet}_absent_presets")))
    if not absent:
        return np.zeros(nv, dtype=bool)
    return np.isin(option_table(mesh).preset, list(absent))


def present_triangle_mask(tri: np.ndarray, absent: np.ndarray) -> np.ndarray:
    """Triangles that touch no absent vertex (leaflet_presence.py:156-167)."""
    tri = np.asarray(tri)
    if tri.size == 0:
        return np.zeros(0, dtype=bool)
    if absent.size == 0 or not absent.any():
        return np.ones(len(tri), dtype=bool)
    return ~absent[tri].any(axis=1)


def boundary_rows_mask(mesh) -> np.ndarray:
    nv = len(mesh.vertex_ids)
    out = np.zeros(nv, dtype=bool)
    vids = getattr(mesh, "boundary_vertex_ids", None)
    if vids:
        idx = mesh.vertex_index_to_row
        rows = [idx[v] for v in vids if v in idx]
        if rows:
            out[np.asarray(rows, dtype=np.int64)] = True
    return out


def group_rows_mask(mesh, group: str) -> np.ndarray:
    """Rows tagged as members of ``group`` under any of GROUP_KEYS (bt_selection.py:204-227)."""
    t = option_table(mesh)
    out = np.zeros(len(mesh.vertex_ids), dtype=bool)
    for col in t.groups.values():
        out |= col == group
    return out


def interior_mask(mesh, global_params, leaflet: str) -> np.ndarray:
    """Rows that carry a base term: not on the boundary, not on the tagged interface ring (bt_selection.py:289-330)."""
    interior = ~boundary_rows_mask(mesh)
    group = _gp(global_params, f"bending_tilt_base_term_boundary_group_{leaflet}")
    group = None if group is None else str(group).strip()
    if group:
        interior &= ~group_rows_mask(mesh, group)
    return interior


def _center_xy(global_params) -> np.ndarray:
    raw = _gp(global_params, "tilt_thetaB_center")
    if raw is None:
        raw = _gp(global_params, "pin_to_circle_point")
    if raw is None:
        return np.zeros(2)
    arr = np.asarray(raw, dtype=float).reshape(-1)
    return arr[:2].copy() if arr.size >= 2 else np.zeros(2)


def base_zero_mask(mesh, global_params, leaflet: str, positions: np.ndarray) -> np.ndarray:
    """Rows whose Helfrich base term is set to zero: assume-J0 presets (optionally clipped to a radius) and the
    benchmark-scoped region modes (bt_selection.py:140-201, 229-287; bt_params.py:40-106, 163-173)."""
    nv = len(mesh.vertex_ids)
    out = np.zeros(nv, dtype=bool)
    center = _center_xy(global_params)
    radii = None
    raw = _gp(global_params, f"bending_tilt_assume_J0_presets_{leaflet}")
    if raw is None:
        raw = _gp(global_params, "bending_tilt_assume_J0_presets")
    presets = _names(raw)
    if presets:
        sel = np.isin(option_table(mesh).preset, presets)
        rmax = _gp(global_params, f"bending_tilt_assume_J0_presets_radius_max_{leaflet}")
        if rmax is None:
            rmax = _gp(global_params, "bending_tilt_assume_J0_presets_radius_max")
        if rmax is not None:
            rmax = float(rmax)
            if rmax < 0.0:
                raise ValueError("bending_tilt_assume_J0_presets_radius_max must be >= 0.")
            radii = np.linalg.norm(np.asarray(positions)[:, :2] - center[None, :], axis=1)
            sel &= ~(radii > rmax + 1.0e-12)
        out |= sel
    mode = str(_gp(global_params, "bending_tilt_base_term_region_mode", "off") or "off").strip().lower()
    if mode not in ("off", "physical_disk_split_v1", "disk_only_base_term_v1"):
        raise ValueError("bending_tilt_base_term_region_mode must be 'off' or 'physical_disk_split_v1' or "
                         "'disk_only_base_term_v1'.")
    if mode != "off":
        radius = _gp(global_params, "bending_tilt_base_term_region_radius")
        if radius is None:
            raise ValueError("bending_tilt_base_term_region_radius is required when "
                             "bending_tilt_base_term_region_mode is enabled.")
        radius = float(radius)
        if radius < 0.0:
            raise ValueError("bending_tilt_base_term_region_radius must be >= 0.")
        if radii is None:
            radii = np.linalg.norm(np.asarray(positions)[:, :2] - center[None, :], axis=1)
        if mode == "physical_disk_split_v1" and leaflet == "out":
            out |= radii <= radius + 1.0e-12
        elif mode == "disk_only_base_term_v1" and leaflet == "in":
            out |= radii > radius + 1.0e-12
    return out


def per_vertex_params(mesh, global_params, leaflet: str) -> tuple[np.ndarray, np.ndarray]:
    """(kappa, c0) per vertex with the leaflet defaults and the vertex overrides (bt_params.py:225-318, helfrich)."""
    nv = len(mesh.vertex_ids)
    t = option_table(mesh)
    k_key, c_key = f"bending_modulus_{leaflet}", f"spontaneous_curvature_{leaflet}"
    k_def = _gp(global_params, k_key)
    if k_def is None:
        k_def = _gp(global_params, "bending_modulus", 0.0)
    c_def = _gp(global_params, c_key)
    if c_def is None:
        c_def = _gp(global_params, "spontaneous_curvature")
        if c_def is None:
            c_def = _gp(global_params, "intrinsic_curvature", 0.0)
    kappa = np.full(nv, float(k_def or 0.0))
    c0 = np.full(nv, float(c_def or 0.0))
    # precedence: the generic key first, the leaflet-specific key last (it wins)
    for key in ("bending_modulus", k_key):
        rows, vals = t.numeric.get(key, (np.zeros(0, np.int64), np.zeros(0)))
        kappa[rows] = vals
    for key in ("intrinsic_curvature", "spontaneous_curvature", c_key):
        rows, vals = t.numeric.get(key, (np.zeros(0, np.int64), np.zeros(0)))
        c0[rows] = vals
    return kappa, c0


def tilt_parameters(param_resolver, leaflet: str) -> dict:
    """Tilt modulus (with the reference's legacy spelling), mass mode, smoothness rigidity."""
    out = {}
    if param_resolver is None:
        return out
    k = param_resolver.get(None, f"tilt_modulus_{leaflet}")
    if k is None:
        k = param_resolver.get(None, f"tilt_modolus_{leaflet}")   # tilt_params.py:9-11
    out["k_tilt"] = float(k or 0.0)
    mode = param_resolver.get(None, f"tilt_mass_mode_{leaflet}")
    if mode is None:
        mode = param_resolver.get(None, "tilt_mass_mode")
    txt = str(mode or "lumped").strip().lower()
    if txt not in ("lumped", "consistent"):
        raise ValueError(f"tilt_mass_mode_{leaflet} must be 'lumped' or 'consistent'.")
    out["consistent"] = txt == "consistent"
    k_s = param_resolver.get(None, f"bending_modulus_{leaflet}")     # tilt_smoothness_utils.py:77-84
    if k_s is None:
        k_s = param_resolver.get(None, "bending_modulus")
    out["k_smooth"] = float(k_s or 0.0)
    return out


def leaflet_selection(mesh, global_params, param_resolver, leaflet: str, positions=None) -> dict:
    """Everything ``struct ms_leaflet_desc`` needs, for meshes without the experimental rim modes."""
    tri, _ = mesh.triangle_row_cache()
    tri = np.zeros((0, 3), np.int32) if tri is None else np.asarray(tri, dtype=np.int32)
    if positions is None:
        positions = mesh.positions_view()
    absent = absent_vertex_mask(mesh, global_params, leaflet)
    keep = present_triangle_mask(tri, absent)
    kappa, c0 = per_vertex_params(mesh, global_params, leaflet)
    out = dict(keep_bt=keep, keep_tilt=keep.copy(), interior=interior_mask(mesh, global_params, leaflet),
               base_zero=base_zero_mask(mesh, global_params, leaflet, positions), kappa=kappa, c0=c0,
               row_weight=None)
    out.update(tilt_parameters(param_resolver, leaflet))
    return out
