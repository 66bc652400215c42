"""Array-native Evolver-style vertex averaging.

The geometric core of ``vertex_average`` (``runtime/vertex_average.py:28-120``) for triangle meshes held as
arrays: every edge gets the weight ``w_e`` = sum of the areas of its incident facets (areas taken ONCE, before
any vertex moves), and every movable vertex with at least two usable edges moves to

    x_new = x + 1/4 * sum_e w_e^2 (x_neighbour - x) / sum_e w_e^2 .

Fixed vertices and vertices pinned to a circle stay; an edge is used only if both ends share the same pin group
(``group`` per vertex, -1 = none; ``vertex_average.py:79-93``).  The reference's optional area restoration for
meshes with explicit ``target_area`` options (``:125-167``) is not part of this function.

Like the refinement next to it this is a next-row item: mesh maintenance between
minimisation blocks, a few NumPy passes on the host.
"""

from __future__ import annotations

import numpy as np


def vertex_average_arrays(pos: np.ndarray, tri: np.ndarray, *, movable: np.ndarray | None = None,
                          group: np.ndarray | None = None) -> np.ndarray:
    """Returns the averaged positions (a new array).  ``movable``: (nv,) bool, default all."""
    pos = np.asarray(pos, dtype=np.float64)
    tri = np.asarray(tri, dtype=np.int64).reshape(-1, 3)
    nv = pos.shape[0]
    if tri.shape[0] == 0:
        return pos.copy()
    v0, v1, v2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    area = 0.5 * np.linalg.norm(np.cross(v1 - v0, v2 - v0), axis=1)
    e = np.stack([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]], axis=1).reshape(-1, 2)
    a, b = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
    uniq, inverse = np.unique(a * np.int64(nv) + b, return_inverse=True)
    weight = np.bincount(inverse, weights=np.repeat(area, 3), minlength=uniq.size)
    ea, eb = uniq // nv, uniq % nv
    ok = weight > 0.0
    if group is not None:
        grp = np.asarray(group)
        # an end with a pin group only accepts neighbours of the same group (checked from that end's side)
        use_from_a = ok & ((grp[ea] < 0) | (grp[ea] == grp[eb]))
        use_from_b = ok & ((grp[eb] < 0) | (grp[eb] == grp[ea]))
    else:
        use_from_a = use_from_b = ok
    w2 = weight * weight
    side = pos[eb] - pos[ea]
    xsum = np.zeros_like(pos)
    total = np.zeros(nv)
    used = np.zeros(nv, dtype=np.int64)
    for axis in range(3):
        xsum[:, axis] += np.bincount(ea[use_from_a], weights=(w2 * side[:, axis])[use_from_a], minlength=nv)
        xsum[:, axis] -= np.bincount(eb[use_from_b], weights=(w2 * side[:, axis])[use_from_b], minlength=nv)
    total += np.bincount(ea[use_from_a], weights=w2[use_from_a], minlength=nv)
    total += np.bincount(eb[use_from_b], weights=w2[use_from_b], minlength=nv)
    used += np.bincount(ea[use_from_a], minlength=nv) + np.bincount(eb[use_from_b], minlength=nv)
    degree = np.bincount(ea, minlength=nv) + np.bincount(eb, minlength=nv)
    move = (used > 1) & (total >= 1e-15) & (degree > 1)
    if movable is not None:
        move &= np.asarray(movable, dtype=bool)
    out = pos.copy()
    out[move] += 0.25 * xsum[move] / total[move][:, None]
    return out


__all__ = ["vertex_average_arrays"]
