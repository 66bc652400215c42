// Per-record and per-vertex bodies of the patch kernels, written against
// patch-LOCAL arrays so that the device kernels (shared memory) and the test-only
// host emulator (tests/emul/emul.cpp, heap arrays) execute the very same code.
#pragma once

#include "../../include/ms_b200.h"
#include "ms_math.cuh"
#include "ms_pack.h"

namespace ms {

MS_HD d3 ld3(const double* p, int i) { return make_d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
MS_HD void add3(double* p, int i, d3 v) {
  p[3 * i] += v.x;
  p[3 * i + 1] += v.y;
  p[3 * i + 2] += v.z;
}

// Per-patch partial sums (one row of PatchLaunch::partials).
enum PartialSlot : int {
  PS_E_SURFACE = 0,
  PS_AREA = 1,
  PS_VOLUME6 = 2,   // six times the body volume
  PS_E_BENDING = 3,
  PS_E_TILT = 4,
  PS_E_BENDING_TILT = 5,
  PS_G_G = 8,       // <g,g>, <g,gC>, <gC,gC> over the owned rows (pass B, before projection)
  PS_G_GC = 9,
  PS_GC_GC = 10,
  PS_COUNT = 12,
};

struct LocalA {
  const double* pos;   // (L,3)
  const uint8_t* bfl;  // L  boundary flags; nullptr = closed mesh (no boundary vertex)
  const double* t2;    // L  |tilt|^2 (tilt module only)
  double* accK;        // (P,3)
  double* accAv;       // P
  double* accAe;       // P
  int P;               // owned vertices; local indices >= P are halo (read-only)
};

// Per-facet scalars (primary listing only) + pass-A corner contributions, computed in
// registers; the accumulation into the owned-vertex arrays is a separate step so that
// several thread groups can compute concurrently and accumulate one after the other.
MS_HD CornerA facet_compute_a(FacetRec rec, double gam, const LocalA& s, uint32_t modules,
                              double k_tilt, double* sums) {
  const d3 v0 = ld3(s.pos, rec.a), v1 = ld3(s.pos, rec.b), v2 = ld3(s.pos, rec.c);
  const FacetGeom g = facet_geom(v0, v1, v2);
  if (rec.flags & REC_PRIMARY) {
    const double T = 0.5 * g.S;
    sums[PS_AREA] += T;
    if (g.S >= kSurfaceSkip) {
      if (modules & MS_MOD_SURFACE) sums[PS_E_SURFACE] += gam * T;
      if ((modules & MS_MOD_TILT) && s.t2)
        sums[PS_E_TILT] += 0.5 * k_tilt * ((s.t2[rec.a] + s.t2[rec.b] + s.t2[rec.c]) / 3.0) * T;
    }
    if ((modules & MS_MOD_VOLUME) && (rec.flags & REC_BODY)) sums[PS_VOLUME6] += facet_volume6(v0, v1, v2);
  }
  CornerA c;
  if (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT))
    c = s.bfl ? facet_pass_a(g, s.bfl[rec.a] != 0, s.bfl[rec.b] != 0, s.bfl[rec.c] != 0)
              : facet_pass_a(g, false, false, false);
  return c;
}

MS_HD void facet_accumulate_a(FacetRec rec, const CornerA& c, const LocalA& s, uint32_t modules) {
  if (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT)) {
    if (rec.a < s.P) { add3(s.accK, rec.a, c.K0); s.accAv[rec.a] += c.va0; s.accAe[rec.a] += c.ve0; }
    if (rec.b < s.P) { add3(s.accK, rec.b, c.K1); s.accAv[rec.b] += c.va1; s.accAe[rec.b] += c.ve1; }
    if (rec.c < s.P) { add3(s.accK, rec.c, c.K2); s.accAv[rec.c] += c.va2; s.accAe[rec.c] += c.ve2; }
  }
}

MS_HD void facet_body_a(FacetRec rec, double gam, const LocalA& s, uint32_t modules, double k_tilt,
                        double* sums) {
  const CornerA c = facet_compute_a(rec, gam, s, modules, k_tilt, sums);
  facet_accumulate_a(rec, c, s, modules);
}

// Area-weighted vertex-normal accumulation (bending_utils.py:13-34), only run for
// patches where some interior vertex has |K| <= 1e-15 (bending.py:154-158).
MS_HD void normal_body(FacetRec rec, const double* pos, double* nrm, int P) {
  const d3 v0 = ld3(pos, rec.a), v1 = ld3(pos, rec.b), v2 = ld3(pos, rec.c);
  const d3 n = cross(v1 - v0, v2 - v0);
  if (rec.a < P) add3(nrm, rec.a, n);
  if (rec.b < P) add3(nrm, rec.b, n);
  if (rec.c < P) add3(nrm, rec.c, n);
}

MS_HD bool vertex_needs_normal(const LocalA& s, int i) {
  const d3 K = ld3(s.accK, i);
  return !(sqrt(dot(K, K)) > 1.0e-15) && !(s.bfl && s.bfl[i]);
}

MS_HD VertexSeed vertex_body_a(int i, const LocalA& s, const double* nrm, bool use_normal,
                               double kappa, double c0, bool willmore) {
  d3 n = make_d3(0, 0, 0);
  if (use_normal) {
    n = ld3(nrm, i);
    const double m = sqrt(dot(n, n));
    if (m > 1.0e-15) n = (1.0 / m) * n;
  }
  return vertex_stage(ld3(s.accK, i), s.accAv[i], s.accAe[i], kappa, willmore ? 0.0 : c0,
                      s.bfl && s.bfl[i] != 0, willmore, n, 0.0);
}

struct LocalB {
  const double* pos;   // (L,3)
  const double* seed;  // (L,kSeedStride)   bending only
  const uint8_t* bfl;  // L                 bending only; nullptr = closed mesh
  const double* t2;    // L                 tilt only
  double* accG;        // (P,3) shape gradient
  double* accV;        // (P,3) dV/dx
  double* accAb;       // P     barycentric vertex area (tilt gradient)
  int P;
};

constexpr int kSeedStrideBody = 5;

// Results of one facet in pass B, held in registers between compute and accumulation.
struct FacetOutB {
  CornerG cg;     // shape gradient contributions
  CornerG vg;     // dV/dx contributions (valid when in_body)
  double third;   // T/3 for the barycentric area (tilt), 0 when the facet is skipped
  bool in_body;
};

template <bool BENDING>
MS_HD FacetOutB facet_compute_b(FacetRec rec, double gam, const LocalB& s, uint32_t modules,
                                uint32_t flags, double k_tilt, bool scalars_here, double* sums) {
  FacetOutB o;
  const d3 v0 = ld3(s.pos, rec.a), v1 = ld3(s.pos, rec.b), v2 = ld3(s.pos, rec.c);
  const FacetGeom g = facet_geom(v0, v1, v2);
  const double T = 0.5 * g.S;
  const bool primary = (rec.flags & REC_PRIMARY) != 0;
  o.in_body = (modules & MS_MOD_VOLUME) && (rec.flags & REC_BODY);
  o.third = 0.0;
  if (!(modules & MS_MOD_SURFACE)) gam = 0.0;
  double coeff = 0.0;
  if ((modules & MS_MOD_TILT) && s.t2) {
    coeff = 0.5 * k_tilt * ((s.t2[rec.a] + s.t2[rec.b] + s.t2[rec.c]) / 3.0);
    if (g.S >= kSurfaceSkip) {
      if (primary) sums[PS_E_TILT] += coeff * T;
      o.third = T / 3.0;
    }
  }
  if (scalars_here && primary) {
    sums[PS_AREA] += T;
    if ((modules & MS_MOD_SURFACE) && g.S >= kSurfaceSkip) sums[PS_E_SURFACE] += gam * T;
    if (o.in_body) sums[PS_VOLUME6] += facet_volume6(v0, v1, v2);
  }
  BendIn b;
  if (BENDING) {
    // seed rows: 5 doubles (odd stride, see ms_pack.cpp on bank conflicts)
    const double* sa = s.seed + kSeedStrideBody * rec.a;
    const double* sb = s.seed + kSeedStrideBody * rec.b;
    const double* sc = s.seed + kSeedStrideBody * rec.c;
    b.f0 = make_d3(sa[0], sa[1], sa[2]); b.fe0 = sa[3]; b.fv0 = sa[4];
    b.f1 = make_d3(sb[0], sb[1], sb[2]); b.fe1 = sb[3]; b.fv1 = sb[4];
    b.f2 = make_d3(sc[0], sc[1], sc[2]); b.fe2 = sc[3]; b.fv2 = sc[4];
    if (s.bfl) {
      b.i0 = !s.bfl[rec.a]; b.i1 = !s.bfl[rec.b]; b.i2 = !s.bfl[rec.c];
    } else {
      b.i0 = b.i1 = b.i2 = true;
    }
  } else {
    b.f0 = b.f1 = b.f2 = make_d3(0, 0, 0);
    b.fe0 = b.fe1 = b.fe2 = b.fv0 = b.fv1 = b.fv2 = 0.0;
    b.i0 = b.i1 = b.i2 = true;
  }
  o.cg = facet_pass_b<BENDING>(g, gam, coeff, b, (flags & MS_FLAG_APPROX) != 0);
  if (o.in_body) o.vg = facet_volume_grad(v0, v1, v2);
  return o;
}

MS_HD void facet_accumulate_b(FacetRec rec, const FacetOutB& o, const LocalB& s) {
  if (rec.a < s.P) add3(s.accG, rec.a, o.cg.g0);
  if (rec.b < s.P) add3(s.accG, rec.b, o.cg.g1);
  if (rec.c < s.P) add3(s.accG, rec.c, o.cg.g2);
  if (o.in_body) {
    if (rec.a < s.P) add3(s.accV, rec.a, o.vg.g0);
    if (rec.b < s.P) add3(s.accV, rec.b, o.vg.g1);
    if (rec.c < s.P) add3(s.accV, rec.c, o.vg.g2);
  }
  if (s.t2 && o.third != 0.0) {
    if (rec.a < s.P) s.accAb[rec.a] += o.third;
    if (rec.b < s.P) s.accAb[rec.b] += o.third;
    if (rec.c < s.P) s.accAb[rec.c] += o.third;
  }
}

template <bool BENDING>
MS_HD void facet_body_b(FacetRec rec, double gam, const LocalB& s, uint32_t modules, uint32_t flags,
                        double k_tilt, bool scalars_here, double* sums) {
  const FacetOutB o = facet_compute_b<BENDING>(rec, gam, s, modules, flags, k_tilt, scalars_here, sums);
  facet_accumulate_b(rec, o, s);
}

}  // namespace ms
