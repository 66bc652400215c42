"""Array-native 1 -> 4 midpoint refinement of a triangle mesh.

The geometric core of ``refine_triangle_mesh`` (``runtime/refinement.py:287-1133``) for meshes
without per-entity options: every edge gets a midpoint vertex (position = mean of its endpoints,
``refinement.py:692-695``), every facet (v0, v1, v2) becomes the corner facets (v0, m01, m20),
(v1, m12, m01), (v2, m20, m12) and the centre facet (m01, m12, m20), keeping the orientation.
A midpoint is fixed when its edge is fixed, i.e. here when both endpoints are.  Per-facet values
(surface tension, body membership) are inherited by the four children.

The reference's version walks dict-of-objects meshes (51 s per level at 49 k facets, SURVEY.md
section 7 hard part 4); this one is a few NumPy sorts, so the 24 * 4^k facet hierarchy of
BASELINE.json configs[2] (``meshes/bending_cube.yaml`` refined to ~1 M facets) can be built.
It is a next-row item, not part of the hot path.
"""

from __future__ import annotations

import numpy as np


def refine_triangles(pos: np.ndarray, tri: np.ndarray, fixed: np.ndarray | None = None):
    """Returns ``(pos2, tri2, fixed2, parent)``: ``parent[j]`` is the facet row child ``j`` came from."""
    pos = np.asarray(pos, dtype=np.float64)
    tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
    nv, nf = pos.shape[0], tri.shape[0]
    # the three edges of every facet as sorted vertex pairs -> unique edge ids
    e = np.stack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=1).reshape(-1, 2)
    key = np.minimum(e[:, 0], e[:, 1]) * np.int64(nv) + np.maximum(e[:, 0], e[:, 1])
    uniq, inverse = np.unique(key, return_inverse=True)
    a, b = uniq // nv, uniq % nv
    mid = nv + inverse.reshape(nf, 3)               # midpoint vertex of edge k of facet f
    pos2 = np.concatenate([pos, 0.5 * (pos[a] + pos[b])])
    m01, m12, m20 = mid[:, 0], mid[:, 1], mid[:, 2]
    v0, v1, v2 = tri[:, 0], tri[:, 1], tri[:, 2]
    tri2 = np.concatenate([
        np.stack([v0, m01, m20], axis=1), np.stack([v1, m12, m01], axis=1),
        np.stack([v2, m20, m12], axis=1), np.stack([m01, m12, m20], axis=1)]).astype(np.int32)
    parent = np.tile(np.arange(nf, dtype=np.int64), 4)
    fixed2 = None
    if fixed is not None:
        fixed = np.asarray(fixed, dtype=bool)
        fixed2 = np.concatenate([fixed, fixed[a] & fixed[b]])
    return np.ascontiguousarray(pos2), np.ascontiguousarray(tri2), fixed2, parent


def cube_mesh():
    """The centroid-triangulated unit cube the reference loads from ``meshes/cube.json`` /
    ``meshes/bending_cube.yaml`` (``geometry/io_readers.py:928-938``): 14 vertices, 24 facets, outward
    orientation, volume 1."""
    c = np.array([[0, 0, 0], [1, 0, 0], [1, 0, 1], [0, 0, 1], [0, 1, 1], [0, 1, 0], [1, 1, 0], [1, 1, 1]], dtype=float)
    faces = [[0, 1, 2, 3], [5, 4, 7, 6], [1, 6, 7, 2], [0, 3, 4, 5], [3, 2, 7, 4], [0, 5, 6, 1]]
    pos, tri = list(c), []
    for f in faces:
        centre = len(pos)
        pos.append(np.mean(c[f], axis=0))
        for k in range(4):
            tri.append([f[k], f[(k + 1) % 4], centre])
    pos, tri = np.array(pos), np.array(tri, dtype=np.int32)
    # orient outwards: the signed volume of a closed, consistently oriented surface must be +1
    v = pos[tri]
    vol = np.einsum("ij,ij->i", np.cross(v[:, 1], v[:, 2]), v[:, 0]).sum() / 6.0
    if vol < 0:
        tri = tri[:, [0, 2, 1]]
    return pos, np.ascontiguousarray(tri)
