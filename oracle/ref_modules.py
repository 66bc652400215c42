"""NumPy restatement of the reference's energy+gradient algorithm on plain arrays.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the
reference file:line it follows (paths relative to the reference root).  The
functions take dense arrays only -- ``pos (nv,3) f64``, ``tri (nf,3) int32`` --
so they run on meshes the reference's dict-of-objects ``Mesh`` cannot hold.

The scatter order (facet order; column 0, then 1, then 2; ``np.add.at``) is the
reference NumPy path's, so the oracle's rounding matches the reference's to the
last bits on the golden fixtures; that order is NOT part of the contract
(SURVEY.md Appendix A.10), the tolerance is 1e-12 relative.
"""

from __future__ import annotations

import numpy as np

# Optional C kernels (oracle/ckernels.py) injected through the same seam the
# reference uses for its f2py kernels (fortran_kernels/loader.py:15-20).
_C_KERNELS = None


def use_c_kernels(kernels) -> None:
    """Install (or clear, with None) the C restatement of the Fortran kernels."""
    global _C_KERNELS
    _C_KERNELS = kernels


def _cross(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    out = np.empty_like(a)
    out[:, 0] = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    out[:, 1] = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    out[:, 2] = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return out


def _dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.einsum("ij,ij->i", a, b)


def _scatter_vec(out: np.ndarray, tri: np.ndarray, a0, a1, a2) -> None:
    np.add.at(out, tri[:, 0], a0)
    np.add.at(out, tri[:, 1], a1)
    np.add.at(out, tri[:, 2], a2)


def _corners(pos: np.ndarray, tri: np.ndarray):
    return pos[tri[:, 0]], pos[tri[:, 1]], pos[tri[:, 2]]


# --------------------------------------------------------------------------
# Surface tension: modules/energy/surface.py:181-221, surface_energy.f90:51-98
# --------------------------------------------------------------------------
def surface_energy_and_gradient(pos, tri, gamma, grad) -> float:
    """E = sum gamma_f * T_f over facets with |n| >= 1e-12; grad += dE/dx."""
    if _C_KERNELS is not None:
        return _C_KERNELS.surface_energy_and_gradient(pos, tri, gamma, grad)
    v0, v1, v2 = _corners(pos, tri)
    n = _cross(v1 - v0, v2 - v0)
    twice_area = np.linalg.norm(n, axis=1)
    keep = twice_area >= 1e-12
    if not np.any(keep):
        return 0.0
    nhat = n[keep] / twice_area[keep][:, None]
    g = np.asarray(gamma, dtype=float)[keep]
    energy = float(np.dot(g, 0.5 * twice_area[keep]))
    a, b, c = v0[keep], v1[keep], v2[keep]
    half_g = (0.5 * g)[:, None]
    _scatter_vec(
        grad,
        tri[keep],
        half_g * _cross(b - c, nhat),
        half_g * _cross(c - a, nhat),
        half_g * _cross(a - b, nhat),
    )
    return energy


def triangle_areas(pos, tri) -> np.ndarray:
    v0, v1, v2 = _corners(pos, tri)
    return 0.5 * np.linalg.norm(_cross(v1 - v0, v2 - v0), axis=1)


# --------------------------------------------------------------------------
# Body volume: geometry/body.py:70-252, modules/constraints/volume.py:43-66,
# modules/energy/volume.py:94-128
# --------------------------------------------------------------------------
def body_volume(pos, tri_body) -> float:
    v0, v1, v2 = _corners(pos, tri_body)
    return float(_dot(_cross(v1, v2), v0).sum() / 6.0)


def accumulate_volume_gradient(pos, tri_body, grad, factor: float) -> None:
    v0, v1, v2 = _corners(pos, tri_body)
    s = factor / 6.0
    _scatter_vec(grad, tri_body, _cross(v1, v2) * s, _cross(v2, v0) * s, _cross(v0, v1) * s)


def volume_penalty_energy_and_gradient(pos, tri_body, k, v0_target, grad):
    """Penalty mode (body.py:192-252): returns (V, E); grad += k (V-V0) dV/dx."""
    vol = body_volume(pos, tri_body)
    delta = vol - v0_target
    if grad is not None:
        accumulate_volume_gradient(pos, tri_body, grad, k * delta)
    return vol, 0.5 * k * delta**2


def kkt_project_single(grad, g_c) -> float:
    """runtime/constraint_manager.py:294-301: g -= (<g,gC>/<gC,gC>) gC."""
    denom = float(np.sum(g_c * g_c))
    if denom <= 1e-18:
        return 0.0
    lam = float(np.sum(grad * g_c)) / denom
    grad -= lam * g_c
    return lam


# --------------------------------------------------------------------------
# Curvature data: geometry/curvature.py:254-332, tilt_kernels.f90:122-189
# --------------------------------------------------------------------------
def _corner_areas(l0, l1, l2, c0, c1, c2, tri_area):
    """Mixed-Voronoi corner areas with the reference's override order
    (curvature.py:300-315, bending_utils.py:107-117)."""
    ob0, ob1, ob2 = c0 < 0, c1 < 0, c2 < 0
    any_ob = ob0 | ob1 | ob2
    va0 = np.where(~any_ob, (l1 * c1 + l2 * c2) / 8.0, 0.0)
    va1 = np.where(~any_ob, (l2 * c2 + l0 * c0) / 8.0, 0.0)
    va2 = np.where(~any_ob, (l0 * c0 + l1 * c1) / 8.0, 0.0)
    va0 = np.where(ob0, tri_area / 2.0, va0)
    va0 = np.where(ob1 | ob2, tri_area / 4.0, va0)
    va1 = np.where(ob1, tri_area / 2.0, va1)
    va1 = np.where(ob0 | ob2, tri_area / 4.0, va1)
    va2 = np.where(ob2, tri_area / 2.0, va2)
    va2 = np.where(ob0 | ob1, tri_area / 4.0, va2)
    return va0, va1, va2


def curvature_data(pos, tri, nv=None):
    """Return (k_vecs (nv,3), vertex_areas (nv), weights (nf,3), va0, va1, va2)."""
    nv = pos.shape[0] if nv is None else nv
    if _C_KERNELS is not None:
        return _C_KERNELS.compute_curvature_data(pos, tri)
    v0, v1, v2 = _corners(pos, tri)
    e0, e1, e2 = v2 - v1, v0 - v2, v1 - v0
    l0, l1, l2 = _dot(e0, e0), _dot(e1, e1), _dot(e2, e2)
    d = np.maximum(np.linalg.norm(_cross(e1, e2), axis=1), 1e-12)
    c0 = _dot(-e1, e2) / d
    c1 = _dot(-e2, e0) / d
    c2 = _dot(-e0, e1) / d
    k_vecs = np.zeros((nv, 3))
    _scatter_vec(
        k_vecs,
        tri,
        0.5 * (c1[:, None] * -e1 + c2[:, None] * e2),
        0.5 * (c2[:, None] * -e2 + c0[:, None] * e0),
        0.5 * (c0[:, None] * -e0 + c1[:, None] * e1),
    )
    va0, va1, va2 = _corner_areas(l0, l1, l2, c0, c1, c2, 0.5 * d)
    vertex_areas = np.zeros(nv)
    _scatter_vec(vertex_areas, tri, va0, va1, va2)
    weights = np.stack([c0, c1, c2], axis=1)
    return k_vecs, vertex_areas, weights, va0, va1, va2


# --------------------------------------------------------------------------
# Effective areas / vertex normals: modules/energy/bending_utils.py:13-171
# --------------------------------------------------------------------------
def effective_areas(pos, tri, weights, is_boundary):
    """Return (A_eff (nv), va_eff (nf,3)) with boundary-corner redistribution."""
    nv = pos.shape[0]
    v0, v1, v2 = _corners(pos, tri)
    e0, e1, e2 = v2 - v1, v0 - v2, v1 - v0
    l0, l1, l2 = _dot(e0, e0), _dot(e1, e1), _dot(e2, e2)
    tri_area = np.maximum(0.5 * np.linalg.norm(_cross(v1 - v0, v2 - v0), axis=1), 1e-12)
    va0, va1, va2 = _corner_areas(
        l0, l1, l2, weights[:, 0], weights[:, 1], weights[:, 2], tri_area
    )
    va = np.stack([va0, va1, va2], axis=1)
    corner_b = np.asarray(is_boundary, dtype=bool)[tri]
    corner_i = ~corner_b
    n_int = corner_i.sum(axis=1)
    move = (n_int > 0) & corner_b.any(axis=1)
    if np.any(move):
        extra = np.zeros(len(tri))
        extra[move] = (va * corner_b).sum(axis=1)[move] / n_int[move]
        va[move] = va[move] * corner_i[move] + corner_i[move] * extra[move, None]
    a_eff = np.zeros(nv)
    _scatter_vec(a_eff, tri, va[:, 0], va[:, 1], va[:, 2])
    return a_eff, va


def vertex_normals(pos, tri):
    nv = pos.shape[0]
    v0, v1, v2 = _corners(pos, tri)
    n = np.cross(v1 - v0, v2 - v0)
    out = np.zeros((nv, 3))
    _scatter_vec(out, tri, n, n, n)
    mag = np.linalg.norm(out, axis=1)
    ok = mag > 1e-15
    out[ok] /= mag[ok, None]
    return out


# --------------------------------------------------------------------------
# Angle / area derivatives: geometry/bending_derivatives.py:48-102,
# bending_kernels.f90:32-74
# --------------------------------------------------------------------------
def grad_cotan(u, v):
    if _C_KERNELS is not None:
        return _C_KERNELS.grad_cotan_batch(u, v)
    c = _dot(u, v)
    w = _cross(u, v)
    s = np.linalg.norm(w, axis=1)
    ok = s > 1e-15
    gu = np.zeros_like(u)
    gv = np.zeros_like(v)
    if not np.any(ok):
        return gu, gv
    inv_s = 1.0 / s[ok]
    k = (c[ok] / (s[ok] * s[ok] * s[ok]))[:, None]
    gu[ok] = v[ok] * inv_s[:, None] - k * _cross(v[ok], w[ok])
    gv[ok] = u[ok] * inv_s[:, None] - k * _cross(w[ok], u[ok])
    return gu, gv


def grad_triangle_area(u, v):
    w = _cross(u, v)
    s = np.linalg.norm(w, axis=1)
    ok = s > 1e-15
    gu = np.zeros_like(u)
    gv = np.zeros_like(v)
    if np.any(ok):
        inv_s = (1.0 / s[ok])[:, None]
        gu[ok] = 0.5 * _cross(v[ok], w[ok]) * inv_s
        gv[ok] = 0.5 * _cross(w[ok], u[ok]) * inv_s
    return gu, gv


def beltrami_laplacian(weights, tri, field):
    """modules/energy/bending_math.py:111-118, bending_kernels.f90:87-131."""
    if _C_KERNELS is not None:
        return _C_KERNELS.apply_beltrami_laplacian(weights, tri, field)
    c0, c1, c2 = weights[:, 0:1], weights[:, 1:2], weights[:, 2:3]
    f0, f1, f2 = field[tri[:, 0]], field[tri[:, 1]], field[tri[:, 2]]
    out = np.zeros_like(field)
    _scatter_vec(
        out,
        tri,
        0.5 * (c1 * (f0 - f2) + c2 * (f0 - f1)),
        0.5 * (c2 * (f1 - f0) + c0 * (f1 - f2)),
        0.5 * (c0 * (f2 - f1) + c1 * (f2 - f0)),
    )
    return out


# --------------------------------------------------------------------------
# Bending: modules/energy/bending.py:90-181, bending_gradient.py:17-175
# --------------------------------------------------------------------------
def bending_vertex_stage(k_vecs, a_vor, a_eff, kappa, c0, is_boundary, model, tau=None):
    """Per-vertex densities (bending.py:112-144). Returns dict of arrays.

    ``tau`` overrides the curvature term (used by bending_tilt, which adds the
    area-averaged divergence: bending_tilt.py:239-258)."""
    safe = np.maximum(a_vor, 1e-12)
    k_mag = np.linalg.norm(k_vecs, axis=1)
    h = k_mag / (2.0 * safe)
    interior = ~np.asarray(is_boundary, dtype=bool)
    ratio = np.zeros_like(a_eff)
    ok = safe > 1e-15
    ratio[ok] = a_eff[ok] / safe[ok]
    if model == "helfrich":
        term = (2.0 * h) - c0 if tau is None else np.array(tau, dtype=float)
        term[~interior] = 0.0
        energy = float(0.5 * np.sum(kappa * term**2 * a_eff))
        scale_k = kappa * term * ratio
        f_eff = 0.5 * kappa * term**2
        f_vor = -2.0 * kappa * term * ratio * h
    else:
        he = h.copy()
        he[~interior] = 0.0
        term = he
        energy = float(np.sum(kappa * he**2 * a_eff))
        scale_k = kappa * he * ratio
        f_eff = kappa * he**2
        f_vor = -2.0 * kappa * he**2 * ratio
    return dict(h=h, k_mag=k_mag, term=term, energy=energy, scale_k=scale_k,
                f_eff=f_eff, f_vor=f_vor, interior=interior, ratio=ratio)


def _k_direction(k_vecs, k_mag, pos, tri):
    d = np.zeros_like(k_vecs)
    ok = k_mag > 1e-15
    d[ok] = k_vecs[ok] / k_mag[ok][:, None]
    if not np.all(ok):
        d[~ok] = vertex_normals(pos, tri)[~ok]
    else:
        # the reference always evaluates the normals (bending.py:154); no effect
        pass
    return d


def _angle_scatter(out, rows, coef0, coef1, coef2, s0u, s0p, s1u, s1p, s2u, s2p):
    r0, r1, r2 = rows
    k0, k1, k2 = coef0[:, None], coef1[:, None], coef2[:, None]
    np.add.at(out, r1, k0 * s0u)
    np.add.at(out, r2, k0 * s0p)
    np.add.at(out, r0, k0 * -(s0u + s0p))
    np.add.at(out, r2, k1 * s1u)
    np.add.at(out, r0, k1 * s1p)
    np.add.at(out, r1, k1 * -(s1u + s1p))
    np.add.at(out, r0, k2 * s2u)
    np.add.at(out, r1, k2 * s2p)
    np.add.at(out, r2, k2 * -(s2u + s2p))


def _corner_angle_gradients(pos, tri):
    i0, i1, i2 = tri[:, 0], tri[:, 1], tri[:, 2]
    v0, v1, v2 = pos[i0], pos[i1], pos[i2]
    return grad_cotan(v1 - v0, v2 - v0) + grad_cotan(v2 - v1, v0 - v1) + grad_cotan(v0 - v2, v1 - v2)


def backprop_operator_terms(pos, tri, weights, f_k):
    """bending_gradient.py:17-80: -L fK (cotangents frozen) and the cotangent variation of L."""
    i0, i1, i2 = tri[:, 0], tri[:, 1], tri[:, 2]
    v0, v1, v2 = pos[i0], pos[i1], pos[i2]
    g_lin = -beltrami_laplacian(weights, tri, f_k)
    w0 = -0.5 * _dot(f_k[i1] - f_k[i2], v1 - v2)
    w1 = -0.5 * _dot(f_k[i2] - f_k[i0], v2 - v0)
    w2 = -0.5 * _dot(f_k[i0] - f_k[i1], v0 - v1)
    g_cot = np.zeros_like(pos)
    _angle_scatter(g_cot, (i0, i1, i2), w0, w1, w2, *_corner_angle_gradients(pos, tri))
    return g_lin, g_cot


def backprop_area_terms(pos, tri, weights, chi):
    """bending_gradient.py:82-175: Voronoi / effective-area variation with corner weights chi (nf,3)."""
    i0, i1, i2 = tri[:, 0], tri[:, 1], tri[:, 2]
    v0, v1, v2 = pos[i0], pos[i1], pos[i2]
    e0, e1, e2 = v2 - v1, v0 - v2, v1 - v0
    c0, c1, c2 = weights[:, 0], weights[:, 1], weights[:, 2]
    g_area = np.zeros_like(pos)
    obtuse = (c0 < 0) | (c1 < 0) | (c2 < 0)
    std = ~obtuse
    if np.any(std):
        s = std
        x0, x1, x2 = chi[s, 0], chi[s, 1], chi[s, 2]
        r0, r1, r2 = i0[s], i1[s], i2[s]
        for coef, edge, plus, minus in (
            (0.25 * c1[s] * x0, e1[s], r0, r2),
            (0.25 * c2[s] * x0, e2[s], r1, r0),
            (0.25 * c2[s] * x1, e2[s], r1, r0),
            (0.25 * c0[s] * x1, e0[s], r2, r1),
            (0.25 * c0[s] * x2, e0[s], r2, r1),
            (0.25 * c1[s] * x2, e1[s], r0, r2),
        ):
            np.add.at(g_area, plus, coef[:, None] * edge)
            np.add.at(g_area, minus, -coef[:, None] * edge)
        q0 = 0.125 * _dot(e0[s], e0[s]) * (x1 + x2)
        q1 = 0.125 * _dot(e1[s], e1[s]) * (x0 + x2)
        q2 = 0.125 * _dot(e2[s], e2[s]) * (x0 + x1)
        sub = np.ascontiguousarray(tri[s])
        _angle_scatter(g_area, (r0, r1, r2), q0, q1, q2, *_corner_angle_gradients(pos, sub))
    if np.any(obtuse):
        for k, at_k in enumerate((c0 < 0, c1 < 0, c2 < 0)):
            m = at_k & obtuse
            if not np.any(m):
                continue
            gu, gp = grad_triangle_area(pos[i1[m]] - pos[i0[m]], pos[i2[m]] - pos[i0[m]])
            others = [j for j in range(3) if j != k]
            phi = (0.5 * chi[m, k] + 0.25 * chi[m, others[0]] + 0.25 * chi[m, others[1]])[:, None]
            np.add.at(g_area, i1[m], phi * gu)
            np.add.at(g_area, i2[m], phi * gp)
            np.add.at(g_area, i0[m], phi * -(gu + gp))
    return g_area


def corner_weights(tri, interior, corner_f_eff, f_vor):
    """chi_k = (interior corner ? fA_eff,k : mean of fA_eff over the facet's interior corners) + fA_vor."""
    corner_int = interior[tri]
    n_int = corner_int.sum(axis=1)
    mean_int = np.zeros(len(tri))
    has = n_int > 0
    mean_int[has] = (corner_f_eff * corner_int).sum(axis=1)[has] / n_int[has]
    return np.where(corner_int, corner_f_eff, mean_int[:, None]) + f_vor[tri]


def bending_backprop(pos, tri, weights, interior, f_eff, f_vor, f_k, grad):
    """bending_gradient.py:17-175: three-term analytic shape gradient."""
    g_lin, g_cot = backprop_operator_terms(pos, tri, weights, f_k)
    g_area = backprop_area_terms(pos, tri, weights, corner_weights(tri, interior, f_eff[tri], f_vor))
    grad += g_lin + g_cot + g_area


def bending_energy_and_gradient(pos, tri, kappa, c0, is_boundary, grad,
                                model="helfrich", mode="analytic"):
    """modules/energy/bending.py:90-181 (finite_difference mode excluded)."""
    nv = pos.shape[0]
    if len(tri) == 0:
        return 0.0
    kappa = np.broadcast_to(np.asarray(kappa, dtype=float), (nv,))
    c0 = np.broadcast_to(np.asarray(c0 if model == "helfrich" else 0.0, dtype=float), (nv,))
    k_vecs, a_vor, weights, _, _, _ = curvature_data(pos, tri)
    a_eff, _ = effective_areas(pos, tri, weights, is_boundary)
    st = bending_vertex_stage(k_vecs, a_vor, a_eff, kappa, c0, is_boundary, model)
    if grad is None:
        return st["energy"]
    f_k = _k_direction(k_vecs, st["k_mag"], pos, tri) * st["scale_k"][:, None]
    if mode == "approx":
        grad -= beltrami_laplacian(weights, tri, f_k)
        b = np.asarray(is_boundary, dtype=bool)
        if b.any():
            grad[b] = 0.0
        return st["energy"]
    bending_backprop(pos, tri, weights, st["interior"], st["f_eff"], st["f_vor"], f_k, grad)
    return st["energy"]


def bending_energy_per_vertex(pos, tri, kappa, c0, is_boundary, model="helfrich"):
    """modules/energy/bending.py:62-87 (compute_energy_array)."""
    nv = pos.shape[0]
    kappa = np.broadcast_to(np.asarray(kappa, dtype=float), (nv,))
    c0 = np.broadcast_to(np.asarray(c0 if model == "helfrich" else 0.0, dtype=float), (nv,))
    k_vecs, a_vor, weights, _, _, _ = curvature_data(pos, tri)
    a_eff, _ = effective_areas(pos, tri, weights, is_boundary)
    safe = np.maximum(a_vor, 1e-12)
    h = np.linalg.norm(k_vecs / (2.0 * safe[:, None]), axis=1)
    dens = 0.5 * (2.0 * h - c0) ** 2 if model == "helfrich" else h**2
    dens[np.asarray(is_boundary, dtype=bool)] = 0.0
    return kappa * dens * a_eff


# --------------------------------------------------------------------------
# Tilt magnitude: modules/energy/tilt.py:99-219 (lumped mass, single field)
# --------------------------------------------------------------------------
def tilt_energy_and_gradient(pos, tri, tilts, k_tilt, grad=None, tilt_grad=None):
    if k_tilt == 0.0 or len(tri) == 0:
        return 0.0
    v0, v1, v2 = _corners(pos, tri)
    n = _cross(v1 - v0, v2 - v0)
    nn = np.linalg.norm(n, axis=1)
    keep = nn >= 1e-12
    if not np.any(keep):
        return 0.0
    t2 = _dot(tilts, tilts)
    areas = 0.5 * nn[keep]
    coeff = 0.5 * k_tilt * (t2[tri[keep]].sum(axis=1) / 3.0)
    energy = float(np.dot(coeff, areas))
    if grad is not None:
        nhat = n[keep] / nn[keep][:, None]
        c = coeff[:, None]
        _scatter_vec(
            grad,
            tri[keep],
            c * (0.5 * _cross(nhat, v2[keep] - v1[keep])),
            c * (0.5 * _cross(nhat, v0[keep] - v2[keep])),
            c * (0.5 * _cross(nhat, v1[keep] - v0[keep])),
        )
    if tilt_grad is not None:
        # geometry/mesh.py barycentric_vertex_areas: sum of area/3 over kept facets
        a_v = np.zeros(pos.shape[0])
        third = areas / 3.0
        _scatter_vec(a_v, tri[keep], third, third, third)
        tilt_grad += k_tilt * tilts * a_v[:, None]
    return energy


# --------------------------------------------------------------------------
# P1 divergence: geometry/tilt_operators.py:130-175,306-330, tilt_kernels.f90:26-86
# --------------------------------------------------------------------------
def p1_triangle_divergence(pos, tilts, tri):
    """Return (div (nf), area (nf), g0, g1, g2 (nf,3)); ambient_v1 transport."""
    if _C_KERNELS is not None:
        return _C_KERNELS.p1_triangle_divergence(pos, tilts, tri)
    v0, v1, v2 = _corners(pos, tri)
    n = _cross(v1 - v0, v2 - v0)
    n2 = _dot(n, n)
    den = np.maximum(n2, 1e-20)[:, None]
    g0 = _cross(n, v2 - v1) / den
    g1 = _cross(n, v0 - v2) / den
    g2 = _cross(n, v1 - v0) / den
    div = _dot(tilts[tri[:, 0]], g0) + _dot(tilts[tri[:, 1]], g1) + _dot(tilts[tri[:, 2]], g2)
    return div, 0.5 * np.sqrt(np.maximum(n2, 0.0)), g0, g1, g2


# --------------------------------------------------------------------------
# Bending-tilt coupling (single field): modules/energy/bending_tilt.py:151-482
# --------------------------------------------------------------------------
def bending_tilt_energy_and_gradient(pos, tri, tilts, kappa, c0, is_boundary,
                                     grad=None, tilt_grad=None, mode="analytic", sign=1.0):
    """E = 1/2 sum_f sum_k kappa_k (base_k + sign*div_f)^2 va_eff,k.

    ``sign`` is +1 for the single-field module (bending_tilt.py:233); the
    shape gradient treats div as constant (bending_tilt.py:14-19)."""
    nv = pos.shape[0]
    if len(tri) == 0:
        return 0.0
    kappa = np.broadcast_to(np.asarray(kappa, dtype=float), (nv,))
    c0 = np.broadcast_to(np.asarray(c0, dtype=float), (nv,))
    k_vecs, a_vor, weights, _, _, _ = curvature_data(pos, tri)
    div, _, g0, g1, g2 = p1_triangle_divergence(pos, tilts, tri)
    div = sign * div
    a_eff, va = effective_areas(pos, tri, weights, is_boundary)
    safe = np.maximum(a_vor, 1e-12)
    k_mag = np.linalg.norm(k_vecs, axis=1)
    h = k_mag / (2.0 * safe)
    interior = ~np.asarray(is_boundary, dtype=bool)
    base = (2.0 * h) - c0
    base[~interior] = 0.0
    term_tri = base[tri] + div[:, None]
    kap_tri = kappa[tri]
    energy = float(0.5 * np.sum(kap_tri * term_tri**2 * va))

    def add_tilt_grad():
        fac = (sign * np.sum(kap_tri * term_tri * va, axis=1))[:, None]
        _scatter_vec(tilt_grad, tri, fac * g0, fac * g1, fac * g2)

    if grad is None:
        if tilt_grad is not None:
            add_tilt_grad()
        return energy

    num = np.zeros(nv)
    num += np.bincount(tri[:, 0], weights=va[:, 0] * div, minlength=nv)
    num += np.bincount(tri[:, 1], weights=va[:, 1] * div, minlength=nv)
    num += np.bincount(tri[:, 2], weights=va[:, 2] * div, minlength=nv)
    div_eff = np.zeros(nv)
    ok = a_eff > 1e-20
    div_eff[ok] = num[ok] / a_eff[ok]
    tau = base + div_eff
    st = bending_vertex_stage(k_vecs, a_vor, a_eff, kappa, c0, is_boundary, "helfrich", tau=tau)
    f_k = _k_direction(k_vecs, k_mag, pos, tri) * st["scale_k"][:, None]
    if mode == "approx":
        grad -= beltrami_laplacian(weights, tri, f_k)
        b = ~interior
        if b.any():
            grad[b] = 0.0
    else:
        bending_backprop(pos, tri, weights, interior, st["f_eff"], st["f_vor"], f_k, grad)
    if tilt_grad is not None:
        add_tilt_grad()
    return energy


# --------------------------------------------------------------------------
# Fused evaluation used by the benchmarks (config 5): surface + bending +
# volume (lagrange constraint gradient or penalty energy).
# --------------------------------------------------------------------------
def boundary_mask_from_triangles(tri, nv) -> np.ndarray:
    """Vertices on an edge with fewer than two incident facets
    (geometry/mesh.py:304-319), for meshes given as triangle arrays."""
    tri = np.asarray(tri, dtype=np.int64)
    a = np.concatenate([tri[:, 0], tri[:, 1], tri[:, 2]])
    b = np.concatenate([tri[:, 1], tri[:, 2], tri[:, 0]])
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * nv + hi
    uniq, counts = np.unique(key, return_counts=True)
    open_edges = uniq[counts < 2]
    mask = np.zeros(nv, dtype=bool)
    mask[open_edges // nv] = True
    mask[open_edges % nv] = True
    return mask


def fused_surface_bending_volume(pos, tri, gamma, kappa, c0, is_boundary,
                                 model="helfrich", mode="analytic"):
    """Return dict(E_surface, E_bending, volume, area, grad, vol_grad)."""
    grad = np.zeros_like(pos)
    e_s = surface_energy_and_gradient(pos, tri, gamma, grad)
    e_b = bending_energy_and_gradient(pos, tri, kappa, c0, is_boundary, grad, model, mode)
    vol_grad = np.zeros_like(pos)
    accumulate_volume_gradient(pos, tri, vol_grad, 1.0)
    return dict(E_surface=e_s, E_bending=e_b, volume=body_volume(pos, tri),
                area=float(triangle_areas(pos, tri).sum()), grad=grad, vol_grad=vol_grad)


def p1_vertex_divergence(pos, tilts, tri):
    """geometry/tilt_operators.py:414-465: (div_v, area_bary), barycentric-area average of div_f."""
    nv = pos.shape[0]
    if len(tri) == 0:
        return np.zeros(nv), np.zeros(nv)
    div, area, _, _, _ = p1_triangle_divergence(pos, tilts, tri)
    w = area / 3.0
    num, den = np.zeros(nv), np.zeros(nv)
    _scatter_vec(num, tri, w * div, w * div, w * div)
    _scatter_vec(den, tri, w, w, w)
    out = np.zeros(nv)
    ok = den > 1e-20
    out[ok] = num[ok] / den[ok]
    return out, den
