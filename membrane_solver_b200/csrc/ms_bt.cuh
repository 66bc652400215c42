// Bending-tilt coupling (Kozlov-Hamm splay term) on plain global arrays, shared by the device
// kernels and the test-only host emulator.
//
//   E = 1/2 sum_f sum_k kappa_k (base_k + s div_f)^2 va_eff,k        base = 2H - c0 (0 on the boundary)
//
// Reference: modules/energy/bending_tilt.py:151-482 (single field, s = +1), P1 divergence
// geometry/tilt_operators.py:158-175, per-vertex divergence average bending_tilt.py:254-268
// (modules/energy/scatter.py:8-58).  The shape gradient treats div(t) as constant
// (bending_tilt.py:14-19): it is the bending back-propagation with term = base + div_eff,
// i.e. pass B of the patch kernels fed with the seeds computed here.
#pragma once

#include "ms_math.cuh"

namespace ms {

struct BtMesh {
  int32_t nv, nf;
  const int32_t* tri;          // (nf,3)
  const double* pos;           // (nv,3)
  const double* tilts;         // (nv,3)
  const uint8_t* is_boundary;  // nv or nullptr
  const double* kappa;         // nv or nullptr -> kappa_u
  const double* c0;            // nv or nullptr -> c0_u
  double kappa_u, c0_u;
  const int32_t* csr_ptr;      // vertex -> corners (corner id = 3 f + k), facet-major order
  const int32_t* csr_idx;
};

MS_HD d3 bt_row(const double* p, int i) { return make_d3(p[3 * size_t(i)], p[3 * size_t(i) + 1], p[3 * size_t(i) + 2]); }

MS_HD bool bt_facet_ok(const BtMesh& m, int f, int& i0, int& i1, int& i2) {
  i0 = m.tri[3 * size_t(f)];
  i1 = m.tri[3 * size_t(f) + 1];
  i2 = m.tri[3 * size_t(f) + 2];
  return i0 >= 0 && i0 < m.nv && i1 >= 0 && i1 < m.nv && i2 >= 0 && i2 < m.nv;
}

// Effective corner areas and the P1 divergence of one facet.
MS_HD void bt_facet_core(const BtMesh& m, int i0, int i1, int i2, double sign, double ve[3], double& div, P1& p,
                         d3& normal) {
  const FacetGeom g = facet_geom(bt_row(m.pos, i0), bt_row(m.pos, i1), bt_row(m.pos, i2));
  const bool b0 = m.is_boundary && m.is_boundary[i0], b1 = m.is_boundary && m.is_boundary[i1],
             b2 = m.is_boundary && m.is_boundary[i2];
  const CornerA c = facet_pass_a(g, b0, b1, b2);
  ve[0] = c.ve0; ve[1] = c.ve1; ve[2] = c.ve2;
  p = facet_p1(g, bt_row(m.tilts, i0), bt_row(m.tilts, i1), bt_row(m.tilts, i2));
  div = sign * p.div;
  normal = g.n;  // (v1 - v0) x (v2 - v0): the area-weighted facet normal (bending_utils.py:13-34)
}

// Step 1 (per facet): corner payload [va_eff,k * div_f, n.x, n.y, n.z] for the vertex gathers.
MS_HD void bt_facet_a(const BtMesh& m, int f, double sign, double* corner4) {
  double* o = corner4 + 12 * size_t(f);
  int i0, i1, i2;
  if (!bt_facet_ok(m, f, i0, i1, i2)) {
    for (int k = 0; k < 12; ++k) o[k] = 0.0;
    return;
  }
  double ve[3], div;
  P1 p;
  d3 n;
  bt_facet_core(m, i0, i1, i2, sign, ve, div, p, n);
  for (int k = 0; k < 3; ++k) {
    o[4 * k] = ve[k] * div;
    o[4 * k + 1] = n.x; o[4 * k + 2] = n.y; o[4 * k + 3] = n.z;
  }
}

// Step 2 (per vertex): base term, averaged divergence, seeds of the back-propagation.
MS_HD void bt_vertex(const BtMesh& m, int v, const double* k_vecs, const double* a_vor, const double* a_eff,
                     const double* corner4, double* seeds /*(nv,5)*/, double* base /*nv*/) {
  double num = 0.0;
  d3 n = make_d3(0, 0, 0);
  for (int j = m.csr_ptr[v]; j < m.csr_ptr[v + 1]; ++j) {
    const double* c = corner4 + 4 * size_t(m.csr_idx[j]);
    num += c[0];
    n = n + make_d3(c[1], c[2], c[3]);
  }
  const double nm = sqrt(dot(n, n));
  if (nm > 1.0e-15) n = (1.0 / nm) * n;
  const bool boundary = m.is_boundary && m.is_boundary[v];
  const double kap = m.kappa ? m.kappa[v] : m.kappa_u;
  const double c0 = m.c0 ? m.c0[v] : m.c0_u;
  const double ae = a_eff[v];
  const double div_eff = ae > 1.0e-20 ? num / ae : 0.0;
  const VertexSeed s = vertex_stage(bt_row(k_vecs, v), a_vor[v], ae, kap, c0, boundary, false, n, div_eff);
  double* o = seeds + 5 * size_t(v);
  o[0] = s.fK.x; o[1] = s.fK.y; o[2] = s.fK.z; o[3] = s.fAe; o[4] = s.fAv;
  base[v] = boundary ? 0.0 : 2.0 * s.H - c0;
}

// Step 3 (per facet): energy of the facet; corner payload fac * g_k (3 doubles per corner) when
// tilt gradients are wanted (corner3 may be null).
MS_HD double bt_facet_b(const BtMesh& m, int f, const double* base, double sign, double* corner3) {
  int i0, i1, i2;
  if (!bt_facet_ok(m, f, i0, i1, i2)) {
    if (corner3)
      for (int k = 0; k < 9; ++k) corner3[9 * size_t(f) + k] = 0.0;
    return 0.0;
  }
  double ve[3], div;
  P1 p;
  d3 n;
  bt_facet_core(m, i0, i1, i2, sign, ve, div, p, n);
  const int idx[3] = {i0, i1, i2};
  double e = 0.0, fac = 0.0;
  for (int k = 0; k < 3; ++k) {
    const double kap = m.kappa ? m.kappa[idx[k]] : m.kappa_u;
    const double term = base[idx[k]] + div;
    e += kap * (term * term) * ve[k];
    fac += kap * term * ve[k];
  }
  if (corner3) {
    fac *= sign;
    double* o = corner3 + 9 * size_t(f);
    o[0] = fac * p.g0.x; o[1] = fac * p.g0.y; o[2] = fac * p.g0.z;
    o[3] = fac * p.g1.x; o[4] = fac * p.g1.y; o[5] = fac * p.g1.z;
    o[6] = fac * p.g2.x; o[7] = fac * p.g2.y; o[8] = fac * p.g2.z;
  }
  return 0.5 * e;
}

}  // namespace ms
