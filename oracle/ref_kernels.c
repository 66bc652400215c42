/*
 * Plain-C restatement of the reference's five f2py Fortran subroutines.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used as the checker in
 * tests/ and as the "Fortran-enabled" CPU baseline in bench.py.  gfortran is
 * not available in this image, so the Fortran sources themselves cannot be
 * built; this file follows them statement by statement instead:
 *
 *   oracle_surface_energy_and_gradient  <- fortran_kernels/surface_energy.f90:27-99
 *   oracle_grad_cotan_batch             <- fortran_kernels/bending_kernels.f90:32-74
 *   oracle_apply_beltrami_laplacian     <- fortran_kernels/bending_kernels.f90:87-131
 *   oracle_p1_triangle_divergence       <- fortran_kernels/tilt_kernels.f90:26-86
 *   oracle_compute_curvature_data       <- fortran_kernels/tilt_kernels.f90:88-190
 *
 * Arrays are C row-major (n,3), which is byte-identical to the Fortran (3,n)
 * column-major view the reference passes (surface.py:127-134).  Serial, fp64,
 * int32 indices, like the originals.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline void sub3(const double *a, const double *b, double *o) {
  o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2];
}
static inline void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
static inline double dot3(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
static inline int in_range(int32_t i, int32_t nv) { return i >= 0 && i < nv; }

/* surface_energy.f90:27-99 */
double oracle_surface_energy_and_gradient(int32_t nv, int32_t nf, const double *pos,
                                          const int32_t *tri, const double *gamma,
                                          double *grad, int32_t zero_based) {
  const int32_t shift = zero_based ? 0 : -1;
  double energy = 0.0;
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t i0 = tri[3 * f] + shift, i1 = tri[3 * f + 1] + shift, i2 = tri[3 * f + 2] + shift;
    if (!in_range(i0, nv) || !in_range(i1, nv) || !in_range(i2, nv)) continue;
    const double *v0 = pos + 3 * (size_t)i0, *v1 = pos + 3 * (size_t)i1, *v2 = pos + 3 * (size_t)i2;
    double e1[3], e2[3], n[3], nh[3], d[3], g[3];
    sub3(v1, v0, e1); sub3(v2, v0, e2);
    cross3(e1, e2, n);
    const double a2 = sqrt(dot3(n, n));
    if (a2 < 1.0e-12) continue;
    nh[0] = n[0] / a2; nh[1] = n[1] / a2; nh[2] = n[2] / a2;
    const double gm = gamma[f];
    energy += gm * (0.5 * a2);
    sub3(v1, v2, d); cross3(d, nh, g);
    for (int k = 0; k < 3; ++k) grad[3 * (size_t)i0 + k] += gm * (0.5 * g[k]);
    sub3(v2, v0, d); cross3(d, nh, g);
    for (int k = 0; k < 3; ++k) grad[3 * (size_t)i1 + k] += gm * (0.5 * g[k]);
    sub3(v0, v1, d); cross3(d, nh, g);
    for (int k = 0; k < 3; ++k) grad[3 * (size_t)i2 + k] += gm * (0.5 * g[k]);
  }
  return energy;
}

/* bending_kernels.f90:32-74 */
void oracle_grad_cotan_batch(int32_t n, const double *u, const double *v, double *gu, double *gv) {
  memset(gu, 0, sizeof(double) * 3 * (size_t)n);
  memset(gv, 0, sizeof(double) * 3 * (size_t)n);
  for (int32_t i = 0; i < n; ++i) {
    const double *ui = u + 3 * (size_t)i, *vi = v + 3 * (size_t)i;
    double w[3], vxw[3], wxu[3];
    const double c = dot3(ui, vi);
    cross3(ui, vi, w);
    const double s = sqrt(dot3(w, w));
    if (s <= 1.0e-15) continue;
    const double inv_s = 1.0 / s, inv_s3 = 1.0 / (s * s * s);
    cross3(vi, w, vxw);
    cross3(w, ui, wxu);
    for (int k = 0; k < 3; ++k) {
      gu[3 * (size_t)i + k] = vi[k] * inv_s - (c * inv_s3) * vxw[k];
      gv[3 * (size_t)i + k] = ui[k] * inv_s - (c * inv_s3) * wxu[k];
    }
  }
}

/* bending_kernels.f90:87-131 */
void oracle_apply_beltrami_laplacian(int32_t dim, int32_t nv, int32_t nf, const double *weights,
                                     const int32_t *tri, const double *field, double *out,
                                     int32_t zero_based) {
  const int32_t shift = zero_based ? 0 : -1;
  memset(out, 0, sizeof(double) * (size_t)dim * (size_t)nv);
  for (int32_t f = 0; f < nf; ++f) {
    const double c0 = weights[3 * f], c1 = weights[3 * f + 1], c2 = weights[3 * f + 2];
    const int32_t a = tri[3 * f] + shift, b = tri[3 * f + 1] + shift, c = tri[3 * f + 2] + shift;
    if (!in_range(a, nv) || !in_range(b, nv) || !in_range(c, nv)) continue;
    for (int32_t d = 0; d < dim; ++d) {
      const double f0 = field[(size_t)a * dim + d], f1 = field[(size_t)b * dim + d],
                   f2 = field[(size_t)c * dim + d];
      out[(size_t)a * dim + d] += 0.5 * (c1 * (f0 - f2) + c2 * (f0 - f1));
      out[(size_t)b * dim + d] += 0.5 * (c2 * (f1 - f0) + c0 * (f1 - f2));
      out[(size_t)c * dim + d] += 0.5 * (c0 * (f2 - f1) + c1 * (f2 - f0));
    }
  }
}

/* tilt_kernels.f90:26-86 */
void oracle_p1_triangle_divergence(int32_t nv, int32_t nf, const double *pos, const double *tilts,
                                   const int32_t *tri, double *div_tri, double *area, double *g0,
                                   double *g1, double *g2, int32_t zero_based) {
  const int32_t shift = zero_based ? 0 : -1;
  memset(div_tri, 0, sizeof(double) * (size_t)nf);
  memset(area, 0, sizeof(double) * (size_t)nf);
  memset(g0, 0, sizeof(double) * 3 * (size_t)nf);
  memset(g1, 0, sizeof(double) * 3 * (size_t)nf);
  memset(g2, 0, sizeof(double) * 3 * (size_t)nf);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t a = tri[3 * f] + shift, b = tri[3 * f + 1] + shift, c = tri[3 * f + 2] + shift;
    if (!in_range(a, nv) || !in_range(b, nv) || !in_range(c, nv)) continue;
    const double *p0 = pos + 3 * (size_t)a, *p1 = pos + 3 * (size_t)b, *p2 = pos + 3 * (size_t)c;
    double u[3], w[3], n[3], e0[3], e1[3], e2[3], x[3];
    sub3(p1, p0, u); sub3(p2, p0, w);
    cross3(u, w, n);
    const double n2 = dot3(n, n);
    const double den = n2 > 1.0e-20 ? n2 : 1.0e-20;
    sub3(p2, p1, e0); sub3(p0, p2, e1); sub3(p1, p0, e2);
    cross3(n, e0, x); for (int k = 0; k < 3; ++k) g0[3 * (size_t)f + k] = x[k] / den;
    cross3(n, e1, x); for (int k = 0; k < 3; ++k) g1[3 * (size_t)f + k] = x[k] / den;
    cross3(n, e2, x); for (int k = 0; k < 3; ++k) g2[3 * (size_t)f + k] = x[k] / den;
    div_tri[f] = dot3(tilts + 3 * (size_t)a, g0 + 3 * (size_t)f) +
                 dot3(tilts + 3 * (size_t)b, g1 + 3 * (size_t)f) +
                 dot3(tilts + 3 * (size_t)c, g2 + 3 * (size_t)f);
    area[f] = 0.5 * sqrt(n2 > 0.0 ? n2 : 0.0);
  }
}

/* tilt_kernels.f90:88-190; va0/va1/va2 may be NULL (the Fortran's optional outputs) */
void oracle_compute_curvature_data(int32_t nv, int32_t nf, const double *pos, const int32_t *tri,
                                   double *k_vecs, double *vertex_areas, double *weights,
                                   int32_t zero_based, double *va0_out, double *va1_out,
                                   double *va2_out) {
  const int32_t shift = zero_based ? 0 : -1;
  memset(k_vecs, 0, sizeof(double) * 3 * (size_t)nv);
  memset(vertex_areas, 0, sizeof(double) * (size_t)nv);
  memset(weights, 0, sizeof(double) * 3 * (size_t)nf);
  if (va0_out) memset(va0_out, 0, sizeof(double) * (size_t)nf);
  if (va1_out) memset(va1_out, 0, sizeof(double) * (size_t)nf);
  if (va2_out) memset(va2_out, 0, sizeof(double) * (size_t)nf);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t a = tri[3 * f] + shift, b = tri[3 * f + 1] + shift, c = tri[3 * f + 2] + shift;
    if (!in_range(a, nv) || !in_range(b, nv) || !in_range(c, nv)) continue;
    const double *p0 = pos + 3 * (size_t)a, *p1 = pos + 3 * (size_t)b, *p2 = pos + 3 * (size_t)c;
    double e0[3], e1[3], e2[3], cr[3];
    sub3(p2, p1, e0); sub3(p0, p2, e1); sub3(p1, p0, e2);
    const double l0 = dot3(e0, e0), l1 = dot3(e1, e1), l2 = dot3(e2, e2);
    cross3(e1, e2, cr);
    double a2 = sqrt(dot3(cr, cr));
    if (a2 < 1.0e-12) a2 = 1.0e-12;
    const double t = 0.5 * a2;
    const double c0 = -dot3(e1, e2) / a2, c1 = -dot3(e2, e0) / a2, c2 = -dot3(e0, e1) / a2;
    weights[3 * (size_t)f] = c0; weights[3 * (size_t)f + 1] = c1; weights[3 * (size_t)f + 2] = c2;
    for (int k = 0; k < 3; ++k) {
      k_vecs[3 * (size_t)a + k] += 0.5 * (c1 * (-e1[k]) + c2 * e2[k]);
      k_vecs[3 * (size_t)b + k] += 0.5 * (c2 * (-e2[k]) + c0 * e0[k]);
      k_vecs[3 * (size_t)c + k] += 0.5 * (c0 * (-e0[k]) + c1 * e1[k]);
    }
    const int o0 = c0 < 0.0, o1 = c1 < 0.0, o2 = c2 < 0.0;
    double va0, va1, va2;
    if (!(o0 || o1 || o2)) {
      va0 = (l1 * c1 + l2 * c2) / 8.0;
      va1 = (l2 * c2 + l0 * c0) / 8.0;
      va2 = (l0 * c0 + l1 * c1) / 8.0;
    } else {
      va0 = va1 = va2 = 0.0;
      if (o0) va0 = t / 2.0;
      if (o1 || o2) va0 = t / 4.0;
      if (o1) va1 = t / 2.0;
      if (o0 || o2) va1 = t / 4.0;
      if (o2) va2 = t / 2.0;
      if (o0 || o1) va2 = t / 4.0;
    }
    vertex_areas[a] += va0; vertex_areas[b] += va1; vertex_areas[c] += va2;
    if (va0_out) va0_out[f] = va0;
    if (va1_out) va1_out[f] = va1;
    if (va2_out) va2_out[f] = va2;
  }
}
