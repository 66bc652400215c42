"""Synthetic triangle meshes in dense-array form.

The reference can only build meshes through its dict-of-objects ``Mesh`` and
``runtime/refinement.py`` (51 s per refinement level at 49 k facets), so the
large benchmark meshes are generated directly as arrays here: a class-I
geodesic icosphere of frequency ``n`` (``nf = 20 n^2``, ``nv = 10 n^2 + 2``)
with a smooth seeded radial perturbation, ordered along a space-filling curve
so that contiguous vertex ranges are compact surface patches.
"""

from __future__ import annotations

import numpy as np

_PHI = (1.0 + 5.0**0.5) / 2.0
_ICO_V = np.array(
    [
        [-1, _PHI, 0], [1, _PHI, 0], [-1, -_PHI, 0], [1, -_PHI, 0],
        [0, -1, _PHI], [0, 1, _PHI], [0, -1, -_PHI], [0, 1, -_PHI],
        [_PHI, 0, -1], [_PHI, 0, 1], [-_PHI, 0, -1], [-_PHI, 0, 1],
    ],
    dtype=np.float64,
)
_ICO_F = np.array(
    [
        [0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11],
        [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
        [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9],
        [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1],
    ],
    dtype=np.int64,
)


def icosphere_facet_count(n: int) -> int:
    return 20 * n * n


def frequency_for_facets(target_facets: int) -> int:
    """Smallest frequency whose icosphere has at least ``target_facets`` facets."""
    n = max(1, int(np.floor(np.sqrt(target_facets / 20.0))))
    while 20 * n * n < target_facets:
        n += 1
    return n


def _morton3(q: np.ndarray) -> np.ndarray:
    """Interleave three 21-bit integer coordinates into a 63-bit key."""

    def spread(x):
        x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
        x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
        x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
        x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
        x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
        x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
        return x

    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


def sfc_order(pos: np.ndarray, tri: np.ndarray):
    """Reorder vertices along a Morton curve and facets by their lowest vertex.

    Returns ``(pos, tri, vertex_perm)`` with ``pos_new = pos_old[vertex_perm]``.
    """
    lo = pos.min(axis=0)
    span = np.maximum(pos.max(axis=0) - lo, 1e-300)
    q = np.minimum(((pos - lo) / span * 2097151.0).astype(np.int64), 2097151)
    perm = np.argsort(_morton3(q), kind="stable")
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    tri_new = inv[tri].astype(np.int32)
    order = np.argsort(tri_new.min(axis=1), kind="stable")
    return np.ascontiguousarray(pos[perm]), np.ascontiguousarray(tri_new[order]), perm


def icosphere(n: int, *, perturb: bool = True, reorder: bool = True):
    """Return ``(pos (nv,3) f64, tri (nf,3) i32)`` of a frequency-``n`` icosphere.

    Orientation is outward (positive enclosed volume).  With ``perturb`` the
    radius follows SURVEY.md section 8d: ``r = 1 + 0.02 sum_j a_j sin(k_j.x + phi_j)``
    (``default_rng(0)``) plus ``1e-3 h`` uniform jitter (``default_rng(1)``).
    """
    if n < 1:
        raise ValueError("frequency must be >= 1")
    verts = _ICO_V / np.linalg.norm(_ICO_V[0])
    # barycentric lattice of one face: all (a,b,c), a+b+c = n
    a_idx, b_idx = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    keep = (a_idx + b_idx) <= n
    a = a_idx[keep].astype(np.int64)
    b = b_idx[keep].astype(np.int64)
    c = n - a - b
    m = a.size
    lattice_id = -np.ones((n + 1, n + 1), dtype=np.int64)
    lattice_id[a, b] = np.arange(m)

    # canonical keys: corners, edge points (shared by two faces), interior points
    edge_id = -np.ones((12, 12), dtype=np.int64)
    ne = 0
    for f in _ICO_F:
        for u, v in ((f[0], f[1]), (f[1], f[2]), (f[2], f[0])):
            if edge_id[u, v] < 0:
                edge_id[u, v] = edge_id[v, u] = ne
                ne += 1
    keys = np.empty((20, m), dtype=np.int64)
    pts = np.empty((20, m, 3), dtype=np.float64)
    base_edge = 12
    base_int = 12 + 30 * (n + 1)
    for fi, (A, B, C) in enumerate(_ICO_F):
        k = base_int + fi * (n + 1) * (n + 1) + a * (n + 1) + b
        w = np.stack([a, b, c], axis=1)
        ids = np.array([A, B, C])
        for i, j, z in ((0, 1, 2), (1, 2, 0), (2, 0, 1)):
            on_edge = w[:, z] == 0
            U, V = ids[i], ids[j]
            t = w[:, i] if U < V else w[:, j]  # weight of the lower-numbered end
            k = np.where(on_edge, base_edge + edge_id[U, V] * (n + 1) + t, k)
        for i in range(3):
            k = np.where(w[:, i] == n, ids[i], k)
        keys[fi] = k
        pts[fi] = (a[:, None] * verts[A] + b[:, None] * verts[B] + c[:, None] * verts[C]) / n
    uniq, first, inverse = np.unique(keys.ravel(), return_index=True, return_inverse=True)
    pos = pts.reshape(-1, 3)[first]
    pos /= np.linalg.norm(pos, axis=1)[:, None]
    gid = inverse.reshape(20, m)

    # facets of one face in lattice coordinates
    ia, ib = a[(a + b) <= n - 1], b[(a + b) <= n - 1]
    up = np.stack([lattice_id[ia, ib], lattice_id[ia + 1, ib], lattice_id[ia, ib + 1]], axis=1)
    ja, jb = a[(a + b) <= n - 2], b[(a + b) <= n - 2]
    down = np.stack([lattice_id[ja + 1, jb], lattice_id[ja + 1, jb + 1], lattice_id[ja, jb + 1]], axis=1)
    local = np.concatenate([up, down], axis=0)
    tri = np.concatenate([gid[fi][local] for fi in range(20)], axis=0)

    # outward orientation
    p0, p1, p2 = pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]
    flip = np.einsum("ij,ij->i", np.cross(p1 - p0, p2 - p0), p0 + p1 + p2) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    tri = tri.astype(np.int32)

    if perturb:
        rng = np.random.default_rng(0)
        amp = rng.uniform(0.5, 1.0, size=8)
        kvec = rng.normal(size=(8, 3)) * 3.0
        phase = rng.uniform(0.0, 2.0 * np.pi, size=8)
        radial = 1.0 + 0.02 * np.sum(amp * np.sin(pos @ kvec.T + phase), axis=1) / amp.sum()
        pos = pos * radial[:, None]
        h = np.sqrt(4.0 * np.pi / (tri.shape[0] * (np.sqrt(3.0) / 4.0)))
        pos = pos + 1e-3 * h * np.random.default_rng(1).uniform(-1.0, 1.0, size=pos.shape)
    if reorder:
        pos, tri, _ = sfc_order(pos, tri)
    return np.ascontiguousarray(pos, dtype=np.float64), np.ascontiguousarray(tri, dtype=np.int32)


def open_sheet(nx: int, ny: int, *, jitter: float = 0.0, seed: int = 3):
    """A triangulated rectangular sheet with a boundary (open mesh) for edge-case tests."""
    xs, ys = np.meshgrid(np.arange(nx + 1, dtype=float), np.arange(ny + 1, dtype=float), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel(), np.zeros(xs.size)], axis=1)
    if jitter:
        rng = np.random.default_rng(seed)
        pos += jitter * rng.normal(size=pos.shape)
    vid = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    a, b, c, d = vid[:-1, :-1].ravel(), vid[1:, :-1].ravel(), vid[1:, 1:].ravel(), vid[:-1, 1:].ravel()
    tri = np.concatenate([np.stack([a, b, c], axis=1), np.stack([a, c, d], axis=1)], axis=0)
    return np.ascontiguousarray(pos), np.ascontiguousarray(tri, dtype=np.int32)


def tangent_tilts(pos: np.ndarray, tri: np.ndarray, *, sigma: float = 0.1, seed: int = 2) -> np.ndarray:
    """Random tilt field projected on the vertex tangent planes."""
    rng = np.random.default_rng(seed)
    t = sigma * rng.normal(size=pos.shape)
    n = np.zeros_like(pos)
    fn = np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]])
    for k in range(3):
        np.add.at(n, tri[:, k], fn)
    n /= np.maximum(np.linalg.norm(n, axis=1), 1e-300)[:, None]
    return t - np.einsum("ij,ij->i", t, n)[:, None] * n
