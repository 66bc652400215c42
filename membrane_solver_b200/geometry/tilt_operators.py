"""P1 tilt operators on the B200 path.

Twins of ``geometry/tilt_operators.py:130-465`` with the same keyword signatures and return tuples, for the
``ambient_v1`` transport (``connection_v1`` is out of scope and raises).  Stateless: each call copies its
inputs to the device and the results back (``ms_p1_triangle_divergence``, ``ms_p1_vertex_divergence``); the
energy modules do not go through here -- they use the resident context.
"""

from __future__ import annotations

import numpy as np

from .. import _lib as L


def _ambient(transport_model: str) -> None:
    if str(transport_model or "ambient_v1").strip().lower() != "ambient_v1":
        raise L.B200Error("only the ambient_v1 tilt transport is available on the B200 path")


def _inputs(positions, tri_rows, tilts=None):
    pos = np.ascontiguousarray(positions, dtype=np.float64)
    tri = np.ascontiguousarray(tri_rows, dtype=np.int32).reshape(-1, 3)
    if pos.ndim != 2 or pos.shape[1] != 3:
        raise ValueError("positions must have shape (N_vertices, 3)")
    t = None
    if tilts is not None:
        t = np.ascontiguousarray(tilts, dtype=np.float64)
        if t.shape != pos.shape:
            raise ValueError("tilts must have shape (N_vertices, 3)")
    return pos, tri, t


def p1_triangle_divergence(*, mesh=None, positions, tilts, tri_rows, transport_model: str = "ambient_v1",
                           normals=None):
    """``(div_tri, area, g0, g1, g2)`` (``tilt_operators.py:191-330``)."""
    _ = (mesh, normals)
    _ambient(transport_model)
    pos, tri, t = _inputs(positions, tri_rows, tilts)
    nf = tri.shape[0]
    div, area = np.zeros(nf), np.zeros(nf)
    g0, g1, g2 = np.zeros((nf, 3)), np.zeros((nf, 3)), np.zeros((nf, 3))
    if nf:
        L.check(L.lib().ms_p1_triangle_divergence(pos.shape[0], nf, L.dptr(pos), L.dptr(t), L.iptr(tri), L.dptr(div),
                                                  L.dptr(area), L.dptr(g0), L.dptr(g1), L.dptr(g2), 1))
    return div, area, g0, g1, g2


def p1_triangle_shape_gradients(*, positions, tri_rows):
    """``(area, g0, g1, g2)`` (``tilt_operators.py:130-188``)."""
    pos, tri, _ = _inputs(positions, tri_rows)
    _, area, g0, g1, g2 = p1_triangle_divergence(positions=pos, tilts=np.zeros_like(pos), tri_rows=tri)
    return area, g0, g1, g2


def p1_vertex_divergence(*, n_vertices: int, mesh=None, positions, tilts, tri_rows,
                         transport_model: str = "ambient_v1", normals=None):
    """``(div_v, area_bary)``: barycentric-area average of the triangle divergences (``:414-465``)."""
    _ = (mesh, normals)
    _ambient(transport_model)
    if n_vertices <= 0:
        return np.zeros(0), np.zeros(0)
    pos, tri, t = _inputs(positions, tri_rows, tilts)
    if pos.shape[0] != int(n_vertices):
        raise ValueError("n_vertices does not match positions")
    div_v, area_v = np.zeros(int(n_vertices)), np.zeros(int(n_vertices))
    if tri.shape[0]:
        L.check(L.lib().ms_p1_vertex_divergence(pos.shape[0], tri.shape[0], L.dptr(pos), L.dptr(t), L.iptr(tri),
                                                L.dptr(div_v), L.dptr(area_v), 1))
    return div_v, area_v


__all__ = ["p1_triangle_divergence", "p1_triangle_shape_gradients", "p1_vertex_divergence"]
