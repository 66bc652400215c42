// Per-facet fp64 math of the energy+gradient path, shared by the device kernels
// and (compiled for the host) by the test-only emulator in tests/emul/.
//
// Notation follows SURVEY.md Appendix A: facet (v0,v1,v2), e0=v2-v1, e1=v0-v2,
// e2=v1-v0 (edge k is opposite corner k), n=e1 x e2, S=|n|, q_k = e_k x n.
//
// Reference semantics restated here (file:line relative to the reference root):
//   surface tension      modules/energy/surface.py:181-221, fortran_kernels/surface_energy.f90:51-98
//   body volume          geometry/body.py:192-252
//   curvature data       geometry/curvature.py:254-332, fortran_kernels/tilt_kernels.f90:122-189
//   effective areas      modules/energy/bending_utils.py:83-171
//   vertex stage         modules/energy/bending.py:112-158
//   analytic back-prop   modules/energy/bending_gradient.py:17-175,
//                        geometry/bending_derivatives.py:48-102
//   tilt magnitude       modules/energy/tilt.py:99-172
//   P1 divergence        geometry/tilt_operators.py:158-175, fortran_kernels/tilt_kernels.f90:26-86
//   bending-tilt         modules/energy/bending_tilt.py:151-482
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MS_HD __host__ __device__ __forceinline__
#else
#define MS_HD inline
#endif

namespace ms {

struct d3 {
  double x, y, z;
};

// 16-byte pair: shared/global rows of 2k doubles move as k vectors
struct alignas(16) dd2 {
  double a, b;
};

MS_HD d3 make_d3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
MS_HD d3 operator+(d3 a, d3 b) { return make_d3(a.x + b.x, a.y + b.y, a.z + b.z); }
MS_HD d3 operator-(d3 a, d3 b) { return make_d3(a.x - b.x, a.y - b.y, a.z - b.z); }
MS_HD d3 operator-(d3 a) { return make_d3(-a.x, -a.y, -a.z); }
MS_HD d3 operator*(double s, d3 a) { return make_d3(s * a.x, s * a.y, s * a.z); }
MS_HD double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
MS_HD d3 cross(d3 a, d3 b) {
  return make_d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// r + s*a
MS_HD d3 axpy(double s, d3 a, d3 r) { return make_d3(r.x + s * a.x, r.y + s * a.y, r.z + s * a.z); }

// Thresholds of the reference (SURVEY.md §7 hard part 6).
constexpr double kSurfaceSkip = 1.0e-12;   // surface.py:190: facet skipped if |n| < 1e-12
constexpr double kAreaClamp = 1.0e-12;     // curvature.py:271: 2A = max(|e1 x e2|, 1e-12)
constexpr double kCotGradEps = 1.0e-15;    // bending_derivatives.py:62: grad cot = 0 if S <= 1e-15
constexpr double kP1Clamp = 1.0e-20;       // tilt_operators.py:164: max(|n|^2, 1e-20)

// correctly rounded reciprocal: the device intrinsic is the short sequence (MUFU.RCP64H + Newton steps) without the
// general division's slow path
MS_HD double recip(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// Geometry every per-facet routine starts from.
struct FacetGeom {
  d3 e0, e1, e2, n;
  double S;   // |n| = twice the facet area
  double rS;  // 1/|n| (0 for a zero normal): ONE reciprocal square root serves sqrt and both divisions
};

MS_HD double recip_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

MS_HD FacetGeom facet_geom(d3 v0, d3 v1, d3 v2) {
  FacetGeom g;
  g.e0 = v2 - v1;
  g.e1 = v0 - v2;
  g.e2 = v1 - v0;
  g.n = cross(g.e1, g.e2);
  const double n2 = dot(g.n, g.n);
  g.rS = n2 > 1.0e-300 ? recip_sqrt(n2) : 0.0;
  g.S = n2 * g.rS;
  return g;
}

// 1 / max(S, 1e-12) without a division (curvature.py:271 clamps twice the area)
MS_HD double inv_clamped_area2(const FacetGeom& g) { return g.S >= kAreaClamp ? g.rS : 1.0 / kAreaClamp; }

// ---------------------------------------------------------------------------
// Pass A: cotangents, curvature-vector and mixed-Voronoi corner contributions.
// ---------------------------------------------------------------------------
struct CornerA {
  d3 K0, K1, K2;        // contributions to the integrated curvature vector K_v
  double va0, va1, va2; // raw mixed-Voronoi corner areas (A_vor)
  double ve0, ve1, ve2; // effective corner areas (A_eff, boundary-redistributed)
  double c0, c1, c2;    // cotangents
};

// h_k = c_k / 2 are the HALF cotangents (they save the 1/2 of every cotan-weight formula).
MS_HD void mixed_voronoi(double l0, double l1, double l2, double h0, double h1, double h2,
                         double T, double& va0, double& va1, double& va2) {
  const bool o0 = h0 < 0.0, o1 = h1 < 0.0, o2 = h2 < 0.0;
  if (!(o0 || o1 || o2)) {
    va0 = (l1 * h1 + l2 * h2) * 0.25;
    va1 = (l2 * h2 + l0 * h0) * 0.25;
    va2 = (l0 * h0 + l1 * h1) * 0.25;
  } else {
    // own-angle T/2, then ANY other obtuse corner overrides with T/4 (curvature.py:308-315)
    va0 = (o1 || o2) ? 0.25 * T : (o0 ? 0.5 * T : 0.0);
    va1 = (o0 || o2) ? 0.25 * T : (o1 ? 0.5 * T : 0.0);
    va2 = (o0 || o1) ? 0.25 * T : (o2 ? 0.5 * T : 0.0);
  }
}

// b0..b2: corner is a boundary vertex.
MS_HD CornerA facet_pass_a(const FacetGeom& g, bool b0, bool b1, bool b2) {
  CornerA r;
  const double D = fmax(g.S, kAreaClamp);
  const double invD = inv_clamped_area2(g);
  // e0 = -(e1 + e2): every edge product follows from l1 = e1.e1, l2 = e2.e2, C0 = -e1.e2
  const double l1 = dot(g.e1, g.e1), l2 = dot(g.e2, g.e2), C0 = -dot(g.e1, g.e2);
  const double C1 = l2 - C0, C2 = l1 - C0, l0 = (l1 + l2) - 2.0 * C0;
  const double hD = 0.5 * invD;
  const double h0 = C0 * hD, h1 = C1 * hD, h2 = C2 * hD;  // half cotangents
  r.c0 = 2.0 * h0;  // cotangents: only the stateless curvature kernel stores them
  r.c1 = 2.0 * h1;
  r.c2 = 2.0 * h2;
  // K[i0] += 1/2 (c1 (-e1) + c2 e2), cyclic; the three contributions sum to zero
  r.K0 = h2 * g.e2 - h1 * g.e1;
  r.K1 = -(h0 * g.e1 + (h0 + h2) * g.e2);
  r.K2 = -(r.K0 + r.K1);
  mixed_voronoi(l0, l1, l2, h0, h1, h2, 0.5 * D, r.va0, r.va1, r.va2);
  // bending_utils.py:101-102 clamps the AREA (not twice the area) for A_eff
  const double Te = fmax(0.5 * g.S, kAreaClamp);
  if (Te == 0.5 * D) {
    r.ve0 = r.va0; r.ve1 = r.va1; r.ve2 = r.va2;
  } else {
    mixed_voronoi(l0, l1, l2, h0, h1, h2, Te, r.ve0, r.ve1, r.ve2);
  }
  const int nb = int(b0) + int(b1) + int(b2);
  if (nb == 1 || nb == 2) {
    // boundary corner areas move, in equal shares, to the interior corners (bending_utils.py:121-153)
    const double moved = (b0 ? r.ve0 : 0.0) + (b1 ? r.ve1 : 0.0) + (b2 ? r.ve2 : 0.0);
    const double extra = moved / double(3 - nb);
    r.ve0 = b0 ? 0.0 : r.ve0 + extra;
    r.ve1 = b1 ? 0.0 : r.ve1 + extra;
    r.ve2 = b2 ? 0.0 : r.ve2 + extra;
  }
  return r;
}

// ---------------------------------------------------------------------------
// Vertex stage: densities and back-propagation seeds (bending.py:112-158).
// ---------------------------------------------------------------------------
struct VertexSeed {
  d3 fK;        // K_dir * kappa*term*ratio
  double fAe;   // dE/dA_eff
  double fAv;   // dE/dA_vor
  double E;     // energy contribution of this vertex
  double H;     // |K| / (2 A_vor)
};

// tau_add: extra curvature term added to the Helfrich term at interior vertices
// (area-averaged divergence for bending_tilt, bending_tilt.py:254-258); 0 for bending.
// energy_from_vertex=false leaves E=0 (bending_tilt assembles its energy per facet).
MS_HD VertexSeed vertex_stage(d3 K, double a_vor, double a_eff, double kappa, double c0,
                              bool boundary, bool willmore, d3 normal, double tau_add) {
  VertexSeed s;
  // one reciprocal square root and one reciprocal serve |K|, 1/|K|, H and A_eff / A_vor (the vertex stage runs on
  // the epilogue warps next to the consumers: every fp64 division it saves is issue bandwidth for them)
  const double safe = fmax(a_vor, 1.0e-12);
  const double k2 = dot(K, K);
  const double rk = (k2 > 1.0e-30) ? recip_sqrt(k2) : 0.0;   // |K| > 1e-15
  const double kmag = k2 * rk;
  const double inv_safe = recip(safe);
  const double H = 0.5 * (kmag * inv_safe);
  const double ratio = a_eff * inv_safe;                // safe >= 1e-12 > 1e-15 always
  double scale;
  if (!willmore) {
    const double term = boundary ? 0.0 : (2.0 * H - c0) + tau_add;
    s.E = 0.5 * (kappa * (term * term) * a_eff);
    scale = kappa * term * ratio;
    s.fAe = 0.5 * kappa * (term * term);
    s.fAv = -2.0 * kappa * term * ratio * H;
  } else {
    const double He = boundary ? 0.0 : H;
    s.E = kappa * (He * He) * a_eff;
    scale = kappa * He * ratio;
    s.fAe = kappa * (He * He);
    s.fAv = -2.0 * kappa * (He * He) * ratio;
  }
  d3 dir = (k2 > 1.0e-30) ? rk * K : normal;
  s.fK = scale * dir;
  s.H = H;
  return s;
}

// ---------------------------------------------------------------------------
// Pass B: all shape-gradient contributions of one facet to its three corners.
// ---------------------------------------------------------------------------
struct BendIn {
  d3 f0, f1, f2;          // fK at the corners
  double fe0, fe1, fe2;   // fA_eff
  double fv0, fv1, fv2;   // fA_vor
  bool i0, i1, i2;        // corner is an interior vertex
};

struct CornerG {
  d3 g0, g1, g2;
};

// gamma_eff: surface tension of the facet (0 when the surface module is off).
// area_coeff: extra dE/dT of this facet (tilt magnitude: q_f of tilt.py:146), also
// multiplied by dT/dx = -q_m/(2S); applied under the same |n| >= 1e-12 rule.
template <bool BENDING>
MS_HD CornerG facet_pass_b(const FacetGeom& g, double gamma_eff, double area_coeff,
                           const BendIn& b, bool approx) {
  CornerG r;
  double Bq = 0.0;  // coefficient of q_m = e_m x n, identical for the three corners
  const double invS = (g.S > kCotGradEps) ? g.rS : 0.0;
  if (g.S >= kSurfaceSkip) Bq -= 0.5 * (gamma_eff + area_coeff) * invS;

  if (!BENDING) {
    r.g0 = Bq * cross(g.e0, g.n);
    r.g1 = Bq * cross(g.e1, g.n);
    r.g2 = Bq * cross(g.e2, g.n);
    return r;
  }
  // With e0 = -(e1 + e2) and n = e1 x e2 every contribution is a combination of FOUR vectors:
  // e1, e2 and the seed differences u = fK1 - fK0, w = fK2 - fK0.
  //   q_0 = -C1 e1 + C2 e2,  q_1 = -C0 e1 - l1 e2,  q_2 = l2 e1 + C0 e2   (a x (b x c) rule)
  // and, the energy being translation invariant, g2 = -(g0 + g1).
  const double hD = 0.5 * inv_clamped_area2(g);
  const double l1 = dot(g.e1, g.e1), l2 = dot(g.e2, g.e2), C0 = -dot(g.e1, g.e2);
  const double C1 = l2 - C0, C2 = l1 - C0;
  const double h0 = C0 * hD, h1 = C1 * hD, h2 = C2 * hD;  // half cotangents c_k / 2
  const d3 u = b.f1 - b.f0, w = b.f2 - b.f0;
  // term 1: g -= L fK  (bending_math.py:111-118): g0 = 1/2 (c1 d20 - c2 d01), d20 = w, d01 = -u
  const double P0 = h2, Q0 = h1;
  const double P1 = -(h0 + h2), Q1 = h0;
  double A0 = 0.0, B0 = 0.0, A1 = 0.0, B1 = 0.0;  // coefficients of e1, e2 for corners 0 and 1
  if (!approx) {
    // term 2 weights: dE/dc_k = -1/2 (fK_i - fK_j).(v_i - v_j); t_k = 2 a_k carries the factor
    const double ue1 = dot(u, g.e1), ue2 = dot(u, g.e2), we1 = dot(w, g.e1), we2 = dot(w, g.e2);
    double t0 = (we1 + we2) - (ue1 + ue2);  // (f1 - f2).e0
    double t1 = we1;                        // (f2 - f0).e1
    double t2 = -ue2;                       // (f0 - f1).e2
    // term 3: chi_k = (interior ? fA_eff : mean over interior corners) + fA_vor
    const int ni = int(b.i0) + int(b.i1) + int(b.i2);
    const double sum_i = (b.i0 ? b.fe0 : 0.0) + (b.i1 ? b.fe1 : 0.0) + (b.i2 ? b.fe2 : 0.0);
    // ni is 0..3: select the reciprocal instead of dividing (a double division is ~20 instructions)
    const double mean_i = ni == 3 ? sum_i * (1.0 / 3.0) : (ni == 2 ? sum_i * 0.5 : (ni == 1 ? sum_i : 0.0));
    const double x0 = (b.i0 ? b.fe0 : mean_i) + b.fv0;
    const double x1 = (b.i1 ? b.fe1 : mean_i) + b.fv1;
    const double x2 = (b.i2 ? b.fe2 : mean_i) + b.fv2;
    const bool o0 = h0 < 0.0, o1 = h1 < 0.0, o2 = h2 < 0.0;
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;  // n_k = 2 m_k,  m_k = 1/4 c_k (chi_a + chi_b)
    if (!(o0 || o1 || o2)) {
      const double l0 = (l1 + l2) - 2.0 * C0;
      const double s12 = x1 + x2, s02 = x0 + x2, s01 = x0 + x1;
      t0 += 0.25 * l0 * s12;   // a_k += 1/8 l_k (chi_a + chi_b)
      t1 += 0.25 * l1 * s02;
      t2 += 0.25 * l2 * s01;
      n0 = h0 * s12;
      n1 = h1 * s02;
      n2 = h2 * s01;
    } else {
      // each obtuse corner k adds (1/2 chi_k + 1/4 chi_a + 1/4 chi_b) dT/dx
      double phi = 0.0;
      if (o0) phi += 0.5 * x0 + 0.25 * x1 + 0.25 * x2;
      if (o1) phi += 0.5 * x1 + 0.25 * x0 + 0.25 * x2;
      if (o2) phi += 0.5 * x2 + 0.25 * x0 + 0.25 * x1;
      Bq -= 0.5 * phi * invS;
    }
    // grad cot_k = 0 when S <= 1e-15 (invS == 0 then)
    const double invSh = 0.5 * invS;
    Bq += (t0 * C0 + t1 * C1 + t2 * C2) * (invSh * invS * invS);
    const double aa0 = t0 * invSh, aa1 = t1 * invSh, aa2 = t2 * invSh;  // a_k / S
    // coefficients of (e0, e1, e2) per corner:  corner 0: (aa1-aa2, aa0+m1, -aa0-m2)
    //                                           corner 1: (-aa1-m0, aa2-aa0, aa1+m2)
    // folded onto (e1, e2) with e0 = -(e1 + e2)
    const double k00 = aa1 - aa2, k10 = -(aa1 + 0.5 * n0);
    A0 = (aa0 + 0.5 * n1) - k00;
    B0 = -(aa0 + 0.5 * n2) - k00;
    A1 = (aa2 - aa0) - k10;
    B1 = (aa1 + 0.5 * n2) - k10;
  }
  A0 -= Bq * C1; B0 += Bq * C2;
  A1 -= Bq * C0; B1 -= Bq * l1;
  r.g0 = axpy(A0, g.e1, axpy(B0, g.e2, axpy(P0, u, Q0 * w)));
  r.g1 = axpy(A1, g.e1, axpy(B1, g.e2, axpy(P1, u, Q1 * w)));
  r.g2 = -(r.g0 + r.g1);
  return r;
}

// Volume: V6 = (v1 x v2).v0 (divide the total by 6); dV/dv0 = (v1 x v2)/6, cyclic.
MS_HD double facet_volume6(d3 v0, d3 v1, d3 v2) { return dot(cross(v1, v2), v0); }

MS_HD CornerG facet_volume_grad(d3 v0, d3 v1, d3 v2) {
  CornerG r;
  const double s = 1.0 / 6.0;
  r.g0 = s * cross(v1, v2);
  r.g1 = s * cross(v2, v0);
  r.g2 = s * cross(v0, v1);
  return r;
}

// Patch path: six times dV/dx per corner; the 1/6 is applied once per vertex in the epilogue.
MS_HD CornerG facet_volume_grad6(d3 v0, d3 v1, d3 v2) {
  CornerG r;
  r.g0 = cross(v1, v2);
  r.g1 = cross(v2, v0);
  r.g2 = cross(v0, v1);
  return r;
}

// P1 divergence (ambient_v1): g_k = n x e_k / max(|n|^2, 1e-20), div = sum t_k . g_k.
struct P1 {
  d3 g0, g1, g2;
  double div, area;
};

MS_HD P1 facet_p1(const FacetGeom& g, d3 t0, d3 t1, d3 t2) {
  P1 r;
  const double n2 = dot(g.n, g.n);
  const double inv = 1.0 / fmax(n2, kP1Clamp);
  r.g0 = inv * cross(g.n, g.e0);
  r.g1 = inv * cross(g.n, g.e1);
  r.g2 = inv * cross(g.n, g.e2);
  r.div = dot(t0, r.g0) + dot(t1, r.g1) + dot(t2, r.g2);
  r.area = 0.5 * sqrt(fmax(n2, 0.0));
  return r;
}

// grad cot(u,v) (bending_derivatives.py:48-79), for the stateless shim.
MS_HD void grad_cotan(d3 u, d3 v, d3& gu, d3& gv) {
  const double C = dot(u, v);
  const d3 w = cross(u, v);
  const double S = sqrt(dot(w, w));
  if (S <= kCotGradEps) {
    gu = make_d3(0, 0, 0);
    gv = gu;
    return;
  }
  const double invS = 1.0 / S;
  const double k = C / (S * S * S);
  gu = axpy(-k, cross(v, w), invS * v);
  gv = axpy(-k, cross(w, u), invS * u);
}

}  // namespace ms
