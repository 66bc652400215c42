"""Array-native equiangulation (Delaunay edge flips) of a triangle mesh.

The geometric core of ``equiangulate_mesh`` (``runtime/equiangulation.py:11-148``): an interior edge shared by two
triangles is replaced by the other diagonal of their quadrilateral when the two angles opposite to it add up to
more than pi (plus the reference's anti-cycling margin of 1e-3), measured -- as in ``should_flip_edge``
(``equiangulation.py:151-233``) -- in the local tangent plane spanned by the edge and the averaged facet normals.
Edges whose two endpoints are fixed are left alone (the reference respects ``edge.fixed``), a flip that would
invert a facet normal (``dot < -0.5``), produce a degenerate triangle or duplicate an existing edge is not made
(``flip_edge_safe``, ``equiangulation.py:291-420``).  Sweeps repeat until no edge flips or ``max_iterations``.

The reference walks dict-of-objects meshes and rebuilds its connectivity maps after every flip; here one sweep
evaluates the criterion of ALL interior edges in a few NumPy expressions and flips the violating edges whose two
triangles have not been touched earlier in the same sweep (the others are looked at again in the next sweep).
Vertex positions, vertex count and the orientation of the surface are unchanged; a facet keeps its row, so
per-facet values (surface tension, body membership) stay attached -- both triangles of a flipped pair belong to
the same body on any consistently tagged mesh.  Like refinement and vertex averaging this is a next-row item
(SURVEY.md section 8f rank 4), not part of the hot path: it produces the new ``tri_rows`` the device re-packs.
"""

from __future__ import annotations

import numpy as np

DELAUNAY_MARGIN = 1.0e-3   # equiangulation.py:231


def _edge_table(tri: np.ndarray, nv: int):
    """Interior edges with exactly two incident facets: (a, b, f1, f2, c, d) where f1 holds the directed edge
    a -> b with apex c and f2 holds b -> a with apex d (consistently oriented neighbours)."""
    nf = tri.shape[0]
    t = tri.astype(np.int64)
    tail = t.reshape(-1)                                  # corner k of facet f: edge from t[f,k] to t[f,k+1]
    head = np.roll(t, -1, axis=1).reshape(-1)
    apex = np.roll(t, -2, axis=1).reshape(-1)
    facet = np.repeat(np.arange(nf, dtype=np.int64), 3)
    key = np.minimum(tail, head) * np.int64(nv) + np.maximum(tail, head)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(ks.size, dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    start = np.flatnonzero(first)
    count = np.diff(np.append(start, ks.size))
    pair = start[count == 2]
    h1, h2 = order[pair], order[pair + 1]
    # opposite directions = consistently oriented pair; anything else (same direction: orientation defect) is skipped
    ok = (tail[h1] == head[h2]) & (head[h1] == tail[h2])
    h1, h2 = h1[ok], h2[ok]
    return tail[h1], head[h1], facet[h1], facet[h2], apex[h1], apex[h2], ks


def _should_flip(pos, a, b, c, d, margin):
    """Vectorised ``should_flip_edge``: True where the opposite angles at c and d exceed pi + margin."""
    p1, p2, p3, p4 = pos[a], pos[b], pos[c], pos[d]
    n1 = np.cross(p2 - p1, p3 - p1)
    n2 = np.cross(p4 - p1, p2 - p1)
    n = n1 + n2
    nn = np.linalg.norm(n, axis=1)
    use1 = nn < 1e-12
    n = np.where(use1[:, None], n1, n)
    nn = np.where(use1, np.linalg.norm(n1, axis=1), nn)
    use2 = nn < 1e-12
    n = np.where(use2[:, None], n2, n)
    nn = np.where(use2, np.linalg.norm(n2, axis=1), nn)
    valid = nn >= 1e-12
    n = n / np.where(valid, nn, 1.0)[:, None]
    e = p2 - p1
    en = np.linalg.norm(e, axis=1)
    valid &= en >= 1e-12
    u = e / np.where(en >= 1e-12, en, 1.0)[:, None]
    v = np.cross(n, u)
    vn = np.linalg.norm(v, axis=1)
    valid &= vn >= 1e-12
    v = v / np.where(vn >= 1e-12, vn, 1.0)[:, None]

    def proj(p):
        rel = p - p1
        return np.stack([np.einsum("ij,ij->i", rel, u), np.einsum("ij,ij->i", rel, v)], axis=1)

    q1 = np.zeros((len(a), 2))
    q2, q3, q4 = proj(p2), proj(p3), proj(p4)

    def angle_at(p, x, y):
        va, vb = x - p, y - p
        na, nb = np.linalg.norm(va, axis=1), np.linalg.norm(vb, axis=1)
        good = (na >= 1e-12) & (nb >= 1e-12)
        cos = np.einsum("ij,ij->i", va, vb) / np.where(good, na * nb, 1.0)
        return np.arccos(np.clip(cos, -1.0, 1.0)), good

    t1, g1 = angle_at(q3, q1, q2)
    t2, g2 = angle_at(q4, q1, q2)
    return valid & g1 & g2 & ((t1 + t2) > (np.pi + margin))


def _unit_normals(pos, tri_rows):
    n = np.cross(pos[tri_rows[:, 1]] - pos[tri_rows[:, 0]], pos[tri_rows[:, 2]] - pos[tri_rows[:, 0]])
    m = np.linalg.norm(n, axis=1)
    return n / np.where(m > 0, m, 1.0)[:, None], m


def equiangulate_triangles(pos: np.ndarray, tri: np.ndarray, fixed: np.ndarray | None = None, *,
                           max_iterations: int = 100, margin: float = DELAUNAY_MARGIN):
    """Returns ``(tri2, n_flips)``; ``tri2`` is int32 with the rows (facets) of ``tri`` in place."""
    pos = np.asarray(pos, dtype=np.float64)
    tri = np.ascontiguousarray(np.asarray(tri).reshape(-1, 3), dtype=np.int32).copy()
    nv = pos.shape[0]
    fixed = None if fixed is None else np.asarray(fixed, dtype=bool)
    total = 0
    for _ in range(max_iterations):
        a, b, f1, f2, c, d, keys = _edge_table(tri, nv)
        if a.size == 0:
            break
        want = _should_flip(pos, a, b, c, d, margin)
        if fixed is not None:
            want &= ~(fixed[a] & fixed[b])
        want &= c != d
        # the new diagonal must not exist already (it would become an edge with four facets)
        new_key = np.minimum(c, d) * np.int64(nv) + np.maximum(c, d)
        want &= ~np.isin(new_key, keys)
        idx = np.flatnonzero(want)
        if idx.size == 0:
            break
        # candidate triangles (a, d, c) and (b, c, d): orientation preserved, normals not inverted, not degenerate
        t1 = np.stack([a[idx], d[idx], c[idx]], axis=1)
        t2 = np.stack([b[idx], c[idx], d[idx]], axis=1)
        n1_old, m1_old = _unit_normals(pos, tri[f1[idx]].astype(np.int64))
        n2_old, m2_old = _unit_normals(pos, tri[f2[idx]].astype(np.int64))
        n1_new, m1_new = _unit_normals(pos, t1)
        n2_new, m2_new = _unit_normals(pos, t2)
        good = (m1_old > 0) & (m2_old > 0) & (m1_new > 1e-300) & (m2_new > 1e-300)
        good &= np.einsum("ij,ij->i", n1_new, n1_old) >= -0.5
        good &= np.einsum("ij,ij->i", n2_new, n2_old) >= -0.5
        touched = np.zeros(tri.shape[0], dtype=bool)
        claimed_new = set()
        flips = 0
        for j in np.flatnonzero(good):   # ascending edge key: a fixed, reproducible order
            i = idx[j]
            fa, fb = int(f1[i]), int(f2[i])
            if touched[fa] or touched[fb]:
                continue                 # its quadrilateral changed in this sweep: looked at again in the next one
            nk = int(new_key[i])
            if nk in claimed_new:
                continue
            claimed_new.add(nk)
            tri[fa] = t1[j]
            tri[fb] = t2[j]
            # the four outer edges of the quadrilateral now belong to different triangles: their other neighbours'
            # criteria are stale too, but flipping those in this sweep stays valid only if their own two triangles
            # are untouched -- which the `touched` test guarantees
            touched[fa] = touched[fb] = True
            flips += 1
        if flips == 0:
            break
        total += flips
    return tri, total


def delaunay_violations(pos: np.ndarray, tri: np.ndarray, *, margin: float = DELAUNAY_MARGIN) -> int:
    """Interior edges that still violate the flip criterion (diagnostic, used by the tests)."""
    a, b, _, _, c, d, _ = _edge_table(np.asarray(tri).reshape(-1, 3), np.asarray(pos).shape[0])
    return int(np.count_nonzero(_should_flip(np.asarray(pos, dtype=np.float64), a, b, c, d, margin))) if a.size else 0
