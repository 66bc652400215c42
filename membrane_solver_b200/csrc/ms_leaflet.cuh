// Leaflet tilt modules on plain global arrays, shared by the device kernels and the test-only host
// emulator: tilt_in / tilt_out (tilt magnitude per leaflet) and bending_tilt_in / bending_tilt_out
// (Kozlov-Hamm splay coupling per leaflet).
//
//   E_tilt = sum_{kept f} q_f T_f,   q_f = k/6 sum|t_k|^2 (lumped) | k/12 (sum|t_k|^2 + t0.t1 + t1.t2 + t2.t0)
//   E_smooth = k_s/4 sum_{kept f} (c0 |t1-t2|^2 + c1 |t2-t0|^2 + c2 |t0-t1|^2)   (tilt gradient only)
//   E_bt   = 1/2 sum_{kept f} sum_k kappa_k (base_k + s div_f)^2 va_eff,k
//            base_v = 2 (K_v . n_v) / (2 A_vor,v) - c0_v   on interior rows, else 0   (s = -1 in, +1 out)
//
// Reference: modules/energy/tilt_leaflet.py:26-169; modules/energy/tilt_smoothness_leaflet.py:17-79 with
// tilt_smoothness.py:81-187 (cotangent Dirichlet energy, ambient_v1); modules/energy/bending_tilt_leaflet.py:231-758 with
// bt_payload.py:40-298 (curvature on the COMPLETE mesh, energy on the leaflet's facets),
// bt_gradient.py:89-389 (operator terms over all facets, area terms over kept facets with PER-CORNER
// dE/dA_eff), bt_gradient.py:20-64 (d div/dx for ambient tilts), bt_divergence.py:49-93.
// The selections the reference derives from mesh options arrive as masks (include/ms_b200.h,
// ms_leaflet_desc).  Three per-facet / per-vertex sweeps and fixed-order CSR gathers: no atomics.
#pragma once

#include "ms_math.cuh"

namespace ms {

struct LeafletMesh {
  int32_t nv, nf;
  const int32_t* tri;          // (nf,3), complete mesh
  const double* pos;           // (nv,3)
  const double* tilts;         // (nv,3) tilt field of this leaflet
  const uint8_t* keep;         // nf or nullptr (every facet belongs to the leaflet)
  const uint8_t* is_boundary;  // nv or nullptr: geometric boundary rows
  const uint8_t* interior;     // nv or nullptr (-> not boundary): rows that carry a base term
  const uint8_t* base_zero;    // nv or nullptr: rows whose base term is forced to zero
  const double* kappa;         // nv or nullptr -> kappa_u
  const double* c0;            // nv or nullptr -> c0_u
  double kappa_u, c0_u;
  const double* row_weight;    // nv or nullptr: active-row weights of the tilt magnitude module
  const uint8_t* consistent;   // nf or nullptr -> consistent_u: mass mode per facet
  int32_t consistent_u;
  double k_tilt;
  double k_smooth;             // tilt smoothness rigidity (bending_modulus_in / _out)
  double sign;                 // s
  const int32_t* csr_ptr;      // vertex -> corners (corner id = 3 f + k), facet-major order
  const int32_t* csr_idx;
};

constexpr int kLfCornerA = 9;  // doubles per corner from lf_facet_a: K(3), va, ve, ve*div, n(3)
constexpr int kLfVertex = 5;   // doubles per vertex from lf_vertex: base, fK(3), fA_vor

MS_HD d3 lf_row(const double* p, int i) { return make_d3(p[3 * size_t(i)], p[3 * size_t(i) + 1], p[3 * size_t(i) + 2]); }

MS_HD bool lf_facet_ok(const LeafletMesh& m, int f, int idx[3]) {
  idx[0] = m.tri[3 * size_t(f)];
  idx[1] = m.tri[3 * size_t(f) + 1];
  idx[2] = m.tri[3 * size_t(f) + 2];
  return idx[0] >= 0 && idx[0] < m.nv && idx[1] >= 0 && idx[1] < m.nv && idx[2] >= 0 && idx[2] < m.nv;
}

MS_HD bool lf_boundary(const LeafletMesh& m, int v) { return m.is_boundary && m.is_boundary[v]; }
MS_HD bool lf_kept(const LeafletMesh& m, int f) { return !m.keep || m.keep[f]; }

// Sweep 1 (per facet): curvature / Voronoi payload for every facet; effective areas, divergence and the
// area-weighted normal for the leaflet's facets.
MS_HD void lf_facet_a(const LeafletMesh& m, int f, double* corner) {
  double* o = corner + 3 * kLfCornerA * size_t(f);
  for (int k = 0; k < 3 * kLfCornerA; ++k) o[k] = 0.0;
  int idx[3];
  if (!lf_facet_ok(m, f, idx)) return;
  const FacetGeom g = facet_geom(lf_row(m.pos, idx[0]), lf_row(m.pos, idx[1]), lf_row(m.pos, idx[2]));
  const CornerA c = facet_pass_a(g, lf_boundary(m, idx[0]), lf_boundary(m, idx[1]), lf_boundary(m, idx[2]));
  const d3 K[3] = {c.K0, c.K1, c.K2};
  const double va[3] = {c.va0, c.va1, c.va2}, ve[3] = {c.ve0, c.ve1, c.ve2};
  const bool kept = lf_kept(m, f);
  double div = 0.0;
  if (kept) {
    const P1 p = facet_p1(g, lf_row(m.tilts, idx[0]), lf_row(m.tilts, idx[1]), lf_row(m.tilts, idx[2]));
    div = m.sign * p.div;
  }
  for (int k = 0; k < 3; ++k) {
    double* q = o + kLfCornerA * k;
    q[0] = K[k].x; q[1] = K[k].y; q[2] = K[k].z;
    q[3] = va[k];
    if (kept) {
      q[4] = ve[k];
      q[5] = ve[k] * div;
      q[6] = g.n.x; q[7] = g.n.y; q[8] = g.n.z;
    }
  }
}

// Sweep 2 (per vertex): signed curvature, base term, averaged divergence, back-propagation seeds.
MS_HD void lf_vertex(const LeafletMesh& m, int v, const double* corner, double* vbuf) {
  d3 K = make_d3(0, 0, 0), n = make_d3(0, 0, 0);
  double a_vor = 0.0, a_eff = 0.0, num = 0.0;
  for (int j = m.csr_ptr[v]; j < m.csr_ptr[v + 1]; ++j) {
    const double* q = corner + kLfCornerA * size_t(m.csr_idx[j]);
    K = K + make_d3(q[0], q[1], q[2]);
    a_vor += q[3];
    a_eff += q[4];
    num += q[5];
    n = n + make_d3(q[6], q[7], q[8]);
  }
  const double nm = sqrt(dot(n, n));
  if (nm > 1.0e-15) n = (1.0 / nm) * n;
  const double safe = fmax(a_vor, 1.0e-12);
  const double h = dot(K, n) / (2.0 * safe);
  const bool inter = m.interior ? (m.interior[v] != 0) : !lf_boundary(m, v);
  const double kap = m.kappa ? m.kappa[v] : m.kappa_u;
  const double c0 = m.c0 ? m.c0[v] : m.c0_u;
  const double base = (inter && !(m.base_zero && m.base_zero[v])) ? 2.0 * h - c0 : 0.0;
  const double ratio = (safe > 1.0e-15) ? a_eff / safe : 0.0;
  const double div_eff = a_eff > 1.0e-20 ? num / a_eff : 0.0;
  const double term = inter ? base + div_eff : 0.0;
  const double scale = kap * term * ratio;
  double* o = vbuf + kLfVertex * size_t(v);
  o[0] = base;
  o[1] = scale * n.x; o[2] = scale * n.y; o[3] = scale * n.z;   // d(K.n)/dK = n
  o[4] = -2.0 * kap * term * ratio * h;
}

// coefficient * d(div_P1)/dx of one facet: div = n.w/|n|^2, n = a x b, w = sum_k e_k x t_k
MS_HD void lf_div_shape_gradient(const FacetGeom& g, d3 t0, d3 t1, d3 t2, double coef, CornerG& out) {
  const d3 a = g.e2, b = -1.0 * g.e1;
  const double inv = 1.0 / fmax(dot(g.n, g.n), kP1Clamp);
  const d3 w = cross(g.e0, t0) + cross(g.e1, t1) + cross(g.e2, t2);
  const d3 dn = axpy(-2.0 * dot(g.n, w) * inv * inv, g.n, inv * w);
  const d3 d0 = inv * cross(t0, g.n), d1 = inv * cross(t1, g.n), d2 = inv * cross(t2, g.n);
  const d3 ga = coef * ((cross(b, dn) - d0) + d2);
  const d3 gb = coef * ((cross(dn, a) + d0) - d1);
  out.g1 = out.g1 + ga;
  out.g2 = out.g2 + gb;
  out.g0 = out.g0 - (ga + gb);
}

struct LfEnergies {
  double e_bt, e_tilt, e_smooth;
};

// Sweep 3 (per facet): energies; corner payloads of the shape gradient (9 doubles per facet, may be null)
// and of the tilt gradient (9 per facet, may be null).  with_bt / with_tilt select the modules.
MS_HD LfEnergies lf_facet_b(const LeafletMesh& m, int f, const double* vbuf, bool with_bt, bool with_tilt,
                            bool with_smooth, double* corner_shape, double* corner_tilt) {
  LfEnergies r = {0.0, 0.0, 0.0};
  double* os = corner_shape ? corner_shape + 9 * size_t(f) : nullptr;
  double* ot = corner_tilt ? corner_tilt + 9 * size_t(f) : nullptr;
  if (os) for (int k = 0; k < 9; ++k) os[k] = 0.0;
  if (ot) for (int k = 0; k < 9; ++k) ot[k] = 0.0;
  int idx[3];
  if (!lf_facet_ok(m, f, idx)) return r;
  const bool kept = lf_kept(m, f);
  if (!kept && !with_bt) return r;
  const FacetGeom g = facet_geom(lf_row(m.pos, idx[0]), lf_row(m.pos, idx[1]), lf_row(m.pos, idx[2]));
  const d3 t[3] = {lf_row(m.tilts, idx[0]), lf_row(m.tilts, idx[1]), lf_row(m.tilts, idx[2])};
  d3 tg[3] = {make_d3(0, 0, 0), make_d3(0, 0, 0), make_d3(0, 0, 0)};

  // tilt magnitude (tilt_leaflet.py:77-160); facets with |n| < 1e-12 are skipped (tilt_utils.py:14-25)
  double q_tilt = 0.0;
  if (with_tilt && kept && m.k_tilt != 0.0 && g.S >= kSurfaceSkip) {
    const double w[3] = {m.row_weight ? m.row_weight[idx[0]] : 1.0, m.row_weight ? m.row_weight[idx[1]] : 1.0,
                         m.row_weight ? m.row_weight[idx[2]] : 1.0};
    const d3 s0 = w[0] * t[0], s1 = w[1] * t[1], s2 = w[2] * t[2];
    const double sq = dot(s0, s0) + dot(s1, s1) + dot(s2, s2);
    const bool cons = m.consistent ? (m.consistent[f] != 0) : (m.consistent_u != 0);
    const double area = 0.5 * g.S;
    if (cons) {
      q_tilt = (m.k_tilt / 12.0) * (sq + dot(s0, s1) + dot(s1, s2) + dot(s2, s0));
      const double fa = m.k_tilt * area / 12.0;
      const d3 sum = s0 + s1 + s2;
      tg[0] = (w[0] * fa) * (s0 + sum);
      tg[1] = (w[1] * fa) * (s1 + sum);
      tg[2] = (w[2] * fa) * (s2 + sum);
    } else {
      q_tilt = 0.5 * m.k_tilt * (sq / 3.0);
      const double fa = m.k_tilt * area / 3.0;
      tg[0] = (w[0] * fa) * s0;
      tg[1] = (w[1] * fa) * s1;
      tg[2] = (w[2] * fa) * s2;
    }
    r.e_tilt = q_tilt * area;
  }

  // tilt smoothness (tilt_smoothness.py:103-187): cotangent Dirichlet energy of the ambient field; no shape gradient
  if (with_smooth && kept && m.k_smooth != 0.0) {
    const CornerA c = facet_pass_a(g, false, false, false);
    const d3 d12 = t[1] - t[2], d20 = t[2] - t[0], d01 = t[0] - t[1];
    r.e_smooth = 0.25 * m.k_smooth * (c.c0 * dot(d12, d12) + c.c1 * dot(d20, d20) + c.c2 * dot(d01, d01));
    const double hk = 0.5 * m.k_smooth;
    tg[0] = tg[0] + hk * (c.c2 * d01 - c.c1 * d20);
    tg[1] = tg[1] + hk * (c.c0 * d12 - c.c2 * d01);
    tg[2] = tg[2] + hk * (c.c1 * d20 - c.c0 * d12);
  }

  CornerG cg;
  if (with_bt) {
    BendIn b;
    b.f0 = make_d3(vbuf[kLfVertex * size_t(idx[0]) + 1], vbuf[kLfVertex * size_t(idx[0]) + 2], vbuf[kLfVertex * size_t(idx[0]) + 3]);
    b.f1 = make_d3(vbuf[kLfVertex * size_t(idx[1]) + 1], vbuf[kLfVertex * size_t(idx[1]) + 2], vbuf[kLfVertex * size_t(idx[1]) + 3]);
    b.f2 = make_d3(vbuf[kLfVertex * size_t(idx[2]) + 1], vbuf[kLfVertex * size_t(idx[2]) + 2], vbuf[kLfVertex * size_t(idx[2]) + 3]);
    b.fe0 = b.fe1 = b.fe2 = 0.0;
    b.fv0 = b.fv1 = b.fv2 = 0.0;
    b.i0 = b.i1 = b.i2 = true;
    double d_div = 0.0;
    if (kept) {
      const bool bd[3] = {lf_boundary(m, idx[0]), lf_boundary(m, idx[1]), lf_boundary(m, idx[2])};
      const CornerA c = facet_pass_a(g, bd[0], bd[1], bd[2]);
      const double ve[3] = {c.ve0, c.ve1, c.ve2};
      const P1 p = facet_p1(g, t[0], t[1], t[2]);
      const double div = m.sign * p.div;
      double fe[3], e = 0.0;
      for (int k = 0; k < 3; ++k) {
        const double kap = m.kappa ? m.kappa[idx[k]] : m.kappa_u;
        const double term = vbuf[kLfVertex * size_t(idx[k])] + div;
        fe[k] = 0.5 * kap * (term * term);
        e += kap * (term * term) * ve[k];
        d_div += kap * term * ve[k];
      }
      d_div *= m.sign;
      r.e_bt = 0.5 * e;
      b.fe0 = fe[0]; b.fe1 = fe[1]; b.fe2 = fe[2];
      b.fv0 = vbuf[kLfVertex * size_t(idx[0]) + 4];
      b.fv1 = vbuf[kLfVertex * size_t(idx[1]) + 4];
      b.fv2 = vbuf[kLfVertex * size_t(idx[2]) + 4];
      b.i0 = !bd[0]; b.i1 = !bd[1]; b.i2 = !bd[2];
      tg[0] = axpy(d_div, p.g0, tg[0]);
      tg[1] = axpy(d_div, p.g1, tg[1]);
      tg[2] = axpy(d_div, p.g2, tg[2]);
    }
    if (os) {
      cg = facet_pass_b<true>(g, 0.0, q_tilt, b, false);
      if (kept) lf_div_shape_gradient(g, t[0], t[1], t[2], d_div, cg);
    }
  } else if (os) {
    BendIn b{};
    cg = facet_pass_b<false>(g, 0.0, q_tilt, b, false);
  }
  if (os) {
    os[0] = cg.g0.x; os[1] = cg.g0.y; os[2] = cg.g0.z;
    os[3] = cg.g1.x; os[4] = cg.g1.y; os[5] = cg.g1.z;
    os[6] = cg.g2.x; os[7] = cg.g2.y; os[8] = cg.g2.z;
  }
  if (ot) {
    ot[0] = tg[0].x; ot[1] = tg[0].y; ot[2] = tg[0].z;
    ot[3] = tg[1].x; ot[4] = tg[1].y; ot[5] = tg[1].z;
    ot[6] = tg[2].x; ot[7] = tg[2].y; ot[8] = tg[2].z;
  }
  return r;
}

}  // namespace ms
