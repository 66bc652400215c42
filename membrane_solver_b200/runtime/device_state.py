"""Device twin of one reference ``Mesh``: the B200 counterpart of ``GeometryCache``
(``runtime/energy_context.py:63-276``).

The reference keys its dense host caches on version counters:
topology <- ``(id(mesh), _facet_loops_version, _vertex_ids_version, _topology_version)``,
positions <- ``_version``.  ``DeviceState`` mirrors that: the packed topology (patches,
records, masks) is rebuilt on the device only when the topology key changes (after
``r`` / ``u`` / a new ``Mesh``); positions are uploaded per evaluation (the line search
evaluates at arrays that are not the mesh cache, ``line_search.py:358-382``); per-entity
parameters are re-sent only when their values change.

The state object is attached to the mesh as ``mesh._b200_state`` -- the same way the
reference's modules attach their memo attributes.
"""

from __future__ import annotations

import weakref

import numpy as np

from .. import _lib as L
from ..context import DeviceMesh

# Test seam: a callable returning an object with the DeviceMesh interface.
DEVICE_MESH_FACTORY = DeviceMesh
DEFAULT_DEVICE = 0


def _version(mesh, name):
    return int(getattr(mesh, name, 0) or 0)


def topology_key(mesh):
    return (id(mesh), _version(mesh, "_facet_loops_version"), _version(mesh, "_vertex_ids_version"),
            _version(mesh, "_topology_version"))


def triangle_rows(mesh) -> np.ndarray:
    """``mesh.triangle_row_cache()`` (``mesh.py:597-624``) as (nf,3) int32; raises for polygons."""
    tri, _ = mesh.triangle_row_cache()
    if tri is None:
        if len(getattr(mesh, "facets", ())) == 0:
            return np.zeros((0, 3), dtype=np.int32)
        raise L.B200Error("the B200 path needs a pure triangle mesh (triangle_row_cache() is None); "
                          "there is no CPU fallback")
    return np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)


def boundary_mask(mesh, nv: int) -> np.ndarray | None:
    """Rows of ``mesh.boundary_vertex_ids`` (``mesh.py:304-319``); None for a closed mesh."""
    vids = getattr(mesh, "boundary_vertex_ids", None)
    if not vids:
        return None
    idx = mesh.vertex_index_to_row
    rows = [idx[v] for v in vids if v in idx]
    if not rows:
        return None
    m = np.zeros(nv, dtype=np.uint8)
    m[np.asarray(rows, dtype=np.int64)] = 1
    return m


def fixed_mask(mesh, nv: int) -> np.ndarray | None:
    fm = getattr(mesh, "fixed_mask", None)
    if fm is None:
        return None
    fm = np.asarray(fm() if callable(fm) else fm, dtype=bool)
    if fm.shape != (nv,) or not fm.any():
        return None
    return fm.astype(np.uint8)


def body_entries(mesh):
    """[(body, facet rows, target volume or None)] for every body of the mesh."""
    out = []
    for body in getattr(mesh, "bodies", {}).values():
        rows = body._get_triangle_rows(mesh)
        target = getattr(body, "target_volume", None)
        if target is None:
            target = (getattr(body, "options", None) or {}).get("target_volume")
        out.append((body, None if rows is None else np.asarray(rows, dtype=np.int64), target))
    return out


class DeviceState:
    """Packed topology + parameters of one mesh on one GPU."""

    def __init__(self, device: int | None = None):
        self.dm = DEVICE_MESH_FACTORY(DEFAULT_DEVICE if device is None else device)
        self.key = None
        self.nv = 0
        self.nf = 0
        self.has_boundary = False
        self.boundary = None
        self.body_rows = None      # facet rows flagged REC_BODY on the device
        self._gamma_key = None
        self._bend_key = None
        self._tilt_key = None
        self._k_tilt = None
        self._leaflet_key = {}     # leaflet -> hash of the description held by the device
        self._fixed_key = None     # (Mesh._fixed_flags_version, hash of the mask) the device holds
        self.uploads = 0           # topology uploads (tests assert residency with this)
        self.fixed_uploads = 0     # fixed-mask uploads outside a topology upload

    # -- topology -------------------------------------------------------------
    def sync_topology(self, mesh, positions: np.ndarray) -> None:
        key = topology_key(mesh)
        nv = int(positions.shape[0])
        if key == self.key and nv == self.nv:
            self._sync_fixed(mesh, nv)
            return
        tri = triangle_rows(mesh)
        bodies = body_entries(mesh)
        body = None
        self.body_rows = None
        if len(bodies) == 1 and bodies[0][1] is not None:
            body = np.zeros(tri.shape[0], dtype=np.uint8)
            body[bodies[0][1]] = 1
            self.body_rows = bodies[0][1]
        self.boundary = boundary_mask(mesh, nv)
        self.has_boundary = self.boundary is not None
        fm = fixed_mask(mesh, nv)
        self.dm.set_topology(nv, tri, is_boundary=self.boundary, body_mask=body, fixed_mask=fm, order_hint=positions)
        self._fixed_key = self._fixed_signature(mesh, fm)
        self.key, self.nv, self.nf = key, nv, int(tri.shape[0])
        self._gamma_key = self._bend_key = self._tilt_key = self._k_tilt = None
        self._leaflet_key = {}
        self.uploads += 1

    # -- fixed vertices: own counter in the reference (geometry/mesh.py:211-231) ----------------------
    @staticmethod
    def _fixed_signature(mesh, fm):
        return (_version(mesh, "_fixed_flags_version"), None if fm is None else hash(fm.tobytes()))

    def _sync_fixed(self, mesh, nv: int) -> None:
        """Vertices fixed or released without a topology change: only the mask travels."""
        ver = _version(mesh, "_fixed_flags_version")
        if self._fixed_key is not None and hasattr(mesh, "_fixed_flags_version") and self._fixed_key[0] == ver:
            return  # the reference bumps the counter on every change of a fixed flag
        fm = fixed_mask(mesh, nv)
        sig = self._fixed_signature(mesh, fm)
        if sig != self._fixed_key:
            if self._fixed_key is None or sig[1] != self._fixed_key[1]:
                self.dm.set_fixed_mask(fm)
                self.fixed_uploads += 1
            self._fixed_key = sig

    # -- parameters (sent only when they change) ------------------------------
    def set_gamma(self, gamma) -> None:
        g = np.asarray(gamma, dtype=np.float64)
        key = (g.shape, hash(g.tobytes()))
        if key != self._gamma_key:
            self.dm.set_surface_tension(g if g.ndim else float(g))
            self._gamma_key = key

    def set_bending(self, kappa, c0) -> None:
        k = np.asarray(kappa, dtype=np.float64)
        c = np.asarray(c0, dtype=np.float64)
        key = (k.shape, hash(k.tobytes()), c.shape, hash(c.tobytes()))
        if key != self._bend_key:
            self.dm.set_bending_params(k if k.ndim else float(k), c if c.ndim else float(c))
            self._bend_key = key

    def set_tilts(self, tilts, k_tilt: float) -> None:
        t = np.ascontiguousarray(tilts, dtype=np.float64)  # tilts_view() is F-ordered (mesh.py:407)
        if t.shape != (self.nv, 3):
            raise ValueError("tilts must have shape (N_vertices, 3)")
        self.dm.set_tilts(t)
        if k_tilt != self._k_tilt:
            self.dm.set_tilt_rigidity(float(k_tilt))
            self._k_tilt = float(k_tilt)


    def set_leaflet(self, leaflet: str, module_bits: int, spec: dict, div_sign: float) -> None:
        """Send one leaflet's selections / parameters (``struct ms_leaflet_desc``) when they changed.
        The coupling module may drop more facets than the tilt-magnitude module (transition triangles,
        ``bt_payload.py:131-144``): the mask of the module about to run is the one held by the device."""
        keep = spec.get("keep_bt") if (module_bits & L.MOD_BENDING_TILT) else spec.get("keep_tilt")
        if (module_bits & L.MOD_BENDING_TILT) and (module_bits & (L.MOD_TILT | L.MOD_TILT_SMOOTHNESS)):
            a, b = spec.get("keep_bt"), spec.get("keep_tilt")
            if (a is None) != (b is None) or (a is not None and not np.array_equal(a, b)):
                raise L.B200Error("the leaflet's two modules use different facet selections: evaluate them separately")
        parts = []
        for name, val in (("keep", keep), ("interior", spec.get("interior")), ("base_zero", spec.get("base_zero")),
                          ("kappa", spec.get("kappa", 0.0)), ("c0", spec.get("c0", 0.0)),
                          ("row_weight", spec.get("row_weight")), ("facet_consistent", spec.get("facet_consistent"))):
            parts.append((name, None if val is None else (np.shape(val), hash(np.asarray(val).tobytes()))))
        key = (tuple(parts), float(spec.get("k_tilt", 0.0)), float(spec.get("k_smooth", 0.0)),
               bool(spec.get("consistent", False)), float(div_sign))
        if self._leaflet_key.get(leaflet) == key:
            return

        def uniform(x):
            a = np.asarray(x, dtype=np.float64)
            return float(a.flat[0]) if a.ndim == 0 or (a.size and np.all(a == a.flat[0])) else a

        def mask(x):
            if x is None:
                return None
            m = np.asarray(x, dtype=bool)
            return m.astype(np.uint8)

        bz = mask(spec.get("base_zero"))
        slot = {"in": L.LEAFLET_IN, "out": L.LEAFLET_OUT, "field": L.LEAFLET_FIELD}[leaflet]
        self.dm.set_leaflet(slot, div_sign=div_sign,
                            kappa=uniform(spec.get("kappa", 0.0)), c0=uniform(spec.get("c0", 0.0)),
                            k_tilt=float(spec.get("k_tilt", 0.0)), k_smooth=float(spec.get("k_smooth", 0.0)),
                            facet_keep=None if keep is None or np.all(keep) else mask(keep),
                            interior=mask(spec.get("interior")), base_zero=None if bz is None or not bz.any() else bz,
                            tilt_row_weight=spec.get("row_weight"), facet_consistent=mask(spec.get("facet_consistent")),
                            consistent=bool(spec.get("consistent", False)))
        self._leaflet_key[leaflet] = key


def get_state(mesh, positions: np.ndarray) -> DeviceState:
    """The mesh's device state, with its topology brought up to date."""
    st = getattr(mesh, "_b200_state", None)
    if st is None:
        st = _SIDE_TABLE.get(id(mesh))
    if st is None:
        st = DeviceState()
        try:
            setattr(mesh, "_b200_state", st)
        except AttributeError:  # slotted mesh objects: keep a side table, emptied when the mesh dies
            key = id(mesh)
            _SIDE_TABLE[key] = st
            try:
                weakref.finalize(mesh, _drop_state, key)
            except TypeError:  # not weak-referenceable either: the entry lives as long as the process
                pass
    st.sync_topology(mesh, positions)
    return st


def _drop_state(key: int) -> None:
    st = _SIDE_TABLE.pop(key, None)
    if st is not None:
        st.dm.close()


_SIDE_TABLE: dict[int, DeviceState] = {}


def positions_array(positions) -> np.ndarray:
    p = np.asarray(positions)
    if p.dtype != np.float64:
        raise TypeError("positions must be float64")
    if p.ndim != 2 or p.shape[1] != 3:
        raise ValueError("positions must have shape (N_vertices, 3)")
    return np.ascontiguousarray(p)
