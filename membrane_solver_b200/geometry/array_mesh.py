"""ArrayMesh: a mesh held as dense arrays, exposing exactly the ``Mesh`` surface the hot-path
plugins touch: ``vertex_ids``, ``vertex_index_to_row``, ``facets``,
``facet_vertex_loops``, ``triangle_row_cache()``, ``get_facet_parameter_array()``,
``boundary_vertex_ids``, ``bodies``, ``vertices``, ``fixed_mask``, ``positions_view()``,
``tilts_view()``, ``build_position_cache()`` and the version counters.

The reference's ``Mesh`` (``geometry/mesh.py``) is a dict of Python objects and cannot hold the
10^7-facet benchmark meshes (SURVEY.md section 7 hard part 4); this container can, and the
plugins cannot tell the difference.
"""

from __future__ import annotations

import numpy as np


class _IdentityIndex:
    """vertex id -> row for ids 0..n-1 without materialising a dict."""

    def __init__(self, n: int):
        self.n = int(n)

    def get(self, key, default=None):
        k = int(key)
        return k if 0 <= k < self.n else default

    def __getitem__(self, key):
        k = int(key)
        if not 0 <= k < self.n:
            raise KeyError(key)
        return k

    def __contains__(self, key):
        return 0 <= int(key) < self.n

    def __len__(self):
        return self.n


class ArrayBody:
    def __init__(self, facet_rows, target_volume=None, options=None):
        self.rows = np.asarray(facet_rows, dtype=np.int64)
        self.target_volume = target_volume
        self.options = dict(options or {})

    def _get_triangle_rows(self, mesh):
        return self.rows


class GlobalParams(dict):
    """dict with attribute access, like ``GlobalParameters.get`` / ``.volume_stiffness``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def set(self, name, value):
        """``GlobalParameters.set`` (``parameters/global_parameters.py``)."""
        self[name] = value


class ParamResolver:
    """``ParameterResolver.get(entity, name)``: entity options first, then the global value."""

    def __init__(self, global_params):
        self.global_params = global_params

    def get(self, entity, name):
        opts = getattr(entity, "options", None) or {}
        if name in opts:
            return opts[name]
        return self.global_params.get(name)


class ArrayMesh:
    def __init__(self, positions, tri, *, global_params=None, facet_params=None, bodies=None, fixed=None,
                 tilts=None, vertex_options=None, tilts_in=None, tilts_out=None, leaflets=None,
                 tilt_fixed_in=None, tilt_fixed_out=None):
        self._positions = np.ascontiguousarray(positions, dtype=np.float64)
        self._tri = np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
        nv = self._positions.shape[0]
        self.global_params = global_params if global_params is not None else GlobalParams()
        self.vertex_ids = np.arange(nv, dtype=np.int64)
        self.vertex_index_to_row = _IdentityIndex(nv)
        self.facets = range(self._tri.shape[0])
        self.facet_vertex_loops = True
        self.facet_params = dict(facet_params or {})
        self.bodies = dict(bodies or {})
        self.vertices = {}
        for vid, opts in (vertex_options or {}).items():
            self.vertices[int(vid)] = type("V", (), {"options": dict(opts)})()
        self._fixed = np.zeros(nv, dtype=bool) if fixed is None else np.asarray(fixed, dtype=bool)
        self._tilts = np.zeros((nv, 3)) if tilts is None else np.asarray(tilts, dtype=np.float64)
        self._tilts_in = np.zeros((nv, 3)) if tilts_in is None else np.asarray(tilts_in, dtype=np.float64)
        self._tilts_out = np.zeros((nv, 3)) if tilts_out is None else np.asarray(tilts_out, dtype=np.float64)
        # rows whose leaflet tilt is clamped (the ``tilt_fixed_in`` / ``tilt_fixed_out`` vertex flags)
        self.tilt_fixed_in_mask = np.zeros(nv, bool) if tilt_fixed_in is None else np.asarray(tilt_fixed_in, dtype=bool)
        self.tilt_fixed_out_mask = np.zeros(nv, bool) if tilt_fixed_out is None else np.asarray(tilt_fixed_out, dtype=bool)
        # leaflet -> selections / parameters as plain arrays (keys of modules/energy/_leaflet.py: keep_bt, keep_tilt,
        # interior, base_zero, kappa, c0, k_tilt, consistent, row_weight, facet_consistent)
        self.leaflets = {k: dict(v) for k, v in (leaflets or {}).items()}
        self._version = 0
        self._facet_loops_version = 0
        self._vertex_ids_version = 0
        self._topology_version = 0
        self._fixed_flags_version = 0
        self._boundary = None

    # -- the Mesh surface ----------------------------------------------------------
    def build_position_cache(self):
        return None

    def positions_view(self):
        return self._positions

    def tilts_view(self):
        return self._tilts

    def tilts_in_view(self):
        return self._tilts_in

    def tilts_out_view(self):
        return self._tilts_out

    def set_tilts_in_from_array(self, tilts):
        self._tilts_in = np.array(tilts, dtype=np.float64).reshape(self._positions.shape)

    def set_tilts_out_from_array(self, tilts):
        self._tilts_out = np.array(tilts, dtype=np.float64).reshape(self._positions.shape)

    def leaflet_selection(self, leaflet: str) -> dict:
        """What the reference derives from mesh options (leaflet presence, base-term rows, per-vertex
        leaflet parameters), given here as arrays."""
        if leaflet not in self.leaflets:
            raise KeyError(f"ArrayMesh has no description of leaflet '{leaflet}' (leaflets={{...}})")
        return self.leaflets[leaflet]

    def triangle_row_cache(self):
        return (self._tri if self._tri.shape[0] else None), None

    def get_facet_parameter_array(self, name):
        arr = self.facet_params.get(name)
        if arr is None:
            return np.full(self._tri.shape[0], float(self.global_params.get(name, 0.0) or 0.0))
        return np.asarray(arr, dtype=np.float64)

    @property
    def fixed_mask(self):
        return self._fixed

    def set_fixed(self, fixed) -> None:
        """Fix / release vertices without touching the topology: bumps ``_fixed_flags_version`` like the
        reference's ``Mesh`` does for its own fixed-mask cache (geometry/mesh.py:211-231)."""
        fixed = np.asarray(fixed, dtype=bool)
        if fixed.shape != self._fixed.shape:
            raise ValueError("fixed must have one flag per vertex")
        self._fixed = fixed.copy()
        self._fixed_flags_version = int(getattr(self, "_fixed_flags_version", 0)) + 1

    @property
    def boundary_vertex_ids(self):
        """Vertices of edges with exactly one incident facet (``mesh.py:304-319``)."""
        if self._boundary is None:
            t = self._tri.astype(np.int64)
            e = np.concatenate([t[:, [0, 1]], t[:, [1, 2]], t[:, [2, 0]]])
            e.sort(axis=1)
            nv = self._positions.shape[0]
            keys, counts = np.unique(e[:, 0] * nv + e[:, 1], return_counts=True)
            once = keys[counts == 1]
            self._boundary = set(np.unique(np.concatenate([once // nv, once % nv])).tolist()) if once.size else set()
        return self._boundary

    # -- mutation (what refine / equiangulate / vertex-average do to the version counters) --
    def set_positions(self, positions):
        self._positions = np.ascontiguousarray(positions, dtype=np.float64)
        self._version += 1

    def set_triangles(self, tri):
        self._tri = np.ascontiguousarray(tri, dtype=np.int32).reshape(-1, 3)
        self.facets = range(self._tri.shape[0])
        self._boundary = None
        self._facet_loops_version += 1
        self._topology_version += 1

    def equiangulate(self, max_iterations: int = 100) -> int:
        """The ``u`` command on arrays (``runtime/equiangulation.py:11-148``): Delaunay edge flips in place; facets
        keep their rows, edges between fixed vertices stay.  Returns the number of flips; a topology version bump
        (the device re-packs on the next evaluation) only when something flipped."""
        from .equiangulate import equiangulate_triangles

        tri, flips = equiangulate_triangles(self._positions, self._tri, self._fixed, max_iterations=max_iterations)
        if flips:
            self.set_triangles(tri)
        return flips
