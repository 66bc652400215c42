// Per-step and per-vertex bodies of the patch kernels, written against patch-LOCAL arrays so that
// the device kernels (shared memory) and the test-only host emulator (tests/emul/emul.cpp, heap
// arrays) execute the very same code.
//
// A lane walks triangle strips (ms_pack.h).  It holds three SLOTS; slot k carries the inputs of
// one vertex (position, seeds, flags) and the partial sums of the contributions that vertex has
// received from the facets this lane has evaluated since the vertex was loaded.  Step s replaces
// the content of slot s % 3: the old vertex's partial sums go to its EVENT row (a plain store:
// no other lane writes that row), the new vertex is gathered from the patch-local arrays, and
// the facet formed by the three slots is evaluated.  After the last step every owned vertex
// adds up its event rows in index order.
//
// Patch-local layout (array of structures, rows of 3 / 5 doubles):
//   pos   [L][3]   positions of owned rows 0..P-1 and halo rows P..L-1
//   seed  [L][5]   pass-A results fK(3), fA_eff, fA_vor          (pass B, bending)
//   evA   [E][5]   pass A events: K(3), A_vor, A_eff
//   evV   [E][3]   events of 6 dV/dx (pass A when it produces the volume gradient, else pass B)
//   evG   [E][3]   pass B events: shape gradient
//   evT   [E]      pass B events: barycentric vertex area (tilt magnitude module)
// The facet orientation seen by the slots may be reversed (STEP_NEG): every quantity below is
// invariant under relabelling the corners except the signed volume terms, which take the sign.
#pragma once

#include "../../include/ms_b200.h"
#include "ms_math.cuh"
#include "ms_pack.h"

namespace ms {

// Per-thread running sums (reduced once per kernel; one row per CTA in PatchLaunch::partials).
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

constexpr int kSeedStrideBody = 5;  // global seed rows: fK(3), fA_eff, fA_vor

// array-of-structures row i of an (n,3) array
MS_HD d3 ld3(const double* p, int i) { return make_d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
MS_HD void st3(double* p, size_t i, d3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

MS_HD int step_index(uint32_t w) { return int(w & STEP_INDEX_MASK); }
MS_HD int step_event(uint32_t w) { return int(w >> STEP_EVENT_SHIFT) - 1; }  // -1: none

// ---------------------------------------------------------------------------
// Pass A
// ---------------------------------------------------------------------------
struct SlotA {
  d3 p;          // position
  int32_t bnd;   // boundary flag
  double t2;     // |tilt|^2 (tilt magnitude module)
  d3 K;          // partial sums: integrated curvature vector, mixed-Voronoi and effective corner areas
  double va, ve;
  d3 vg;         // partial sum of six times dV/dx
};

struct LocalA {
  const double* pos;   // [L][3]
  const int32_t* bfl;  // [L] boundary flags; nullptr = closed mesh (no boundary vertex)
  const double* t2;    // [L] |tilt|^2 (tilt module only)
  double* evA;         // [E][5]
  double* evV;         // [E][3]
};

MS_HD void slot_clear_a(SlotA& s) {
  s.K = make_d3(0, 0, 0);
  s.va = s.ve = 0.0;
  s.vg = make_d3(0, 0, 0);
}

// ZERO = false: the step that loads the slot assigns its partial sums (step_compute_a<.., K>)
template <bool ZERO = true>
MS_HD void slot_load_a(SlotA& s, const LocalA& l, int i) {
  s.p = ld3(l.pos, i);
  s.bnd = l.bfl ? l.bfl[i] : 0;
  s.t2 = l.t2 ? l.t2[i] : 0.0;
  if (ZERO) slot_clear_a(s);
}

template <bool BENDING, bool VG>
MS_HD void slot_flush_a(const SlotA& s, const LocalA& l, int e) {
  if (BENDING) {
    double* o = l.evA + 5 * e;
    o[0] = s.K.x; o[1] = s.K.y; o[2] = s.K.z; o[3] = s.va; o[4] = s.ve;
  }
  if (VG) st3(l.evV, e, s.vg);
}

// r = fresh ? c : r + c.  FRESH is a compile-time property of the slot a step has just loaded: its first
// contribution is assigned, so loading a vertex does not have to zero its partial sums.
template <bool FRESH>
MS_HD void acc1(double& r, double c) { r = FRESH ? c : r + c; }
template <bool FRESH>
MS_HD void acc3(d3& r, d3 c) { r = FRESH ? c : r + c; }

// Evaluate the facet held by the three slots: per-facet scalars (primary listing only) and the corner
// contributions of pass A, added to the slots' partial sums.  K (0..2) names the slot this step loaded (its sums
// start here); K = -1: no slot is fresh (host emulator, restart-loaded slots are zeroed when they are loaded).
template <bool BENDING, bool VG, int K = -1>
MS_HD void step_compute_a(SlotA& s0, SlotA& s1, SlotA& s2, uint32_t word, double gam, uint32_t modules, double k_tilt,
                          double* sums) {
  const FacetGeom g = facet_geom(s0.p, s1.p, s2.p);
  const bool body = (modules & MS_MOD_VOLUME) && (word & STEP_BODY);
  const bool primary = (word & STEP_PRIMARY) != 0;
  const double sgn = (word & STEP_NEG) ? -1.0 : 1.0;
  d3 c12 = make_d3(0, 0, 0);
  if (modules & MS_MOD_VOLUME) c12 = (body ? sgn : 0.0) * cross(s1.p, s2.p);
  {  // scalars of the primary listing (branch-free: other listings add zeros)
    const double T = primary ? 0.5 * g.S : 0.0;
    sums[PS_AREA] += T;
    const double Ts = g.S >= kSurfaceSkip ? T : 0.0;
    if (modules & MS_MOD_SURFACE) sums[PS_E_SURFACE] += gam * Ts;
    if (modules & MS_MOD_TILT) sums[PS_E_TILT] += 0.5 * k_tilt * ((s0.t2 + s1.t2 + s2.t2) / 3.0) * Ts;
    if (modules & MS_MOD_VOLUME) sums[PS_VOLUME6] += primary ? dot(c12, s0.p) : 0.0;
  }
  if (BENDING) {
    const CornerA c = facet_pass_a(g, s0.bnd != 0, s1.bnd != 0, s2.bnd != 0);
    acc3<K == 0>(s0.K, c.K0); acc1<K == 0>(s0.va, c.va0); acc1<K == 0>(s0.ve, c.ve0);
    acc3<K == 1>(s1.K, c.K1); acc1<K == 1>(s1.va, c.va1); acc1<K == 1>(s1.ve, c.ve1);
    acc3<K == 2>(s2.K, c.K2); acc1<K == 2>(s2.va, c.va2); acc1<K == 2>(s2.ve, c.ve2);
  }
  if (VG) {  // facets outside the body contribute zeros (c12 and sgn0 vanish)
    const double sg = body ? sgn : 0.0;
    acc3<K == 0>(s0.vg, c12);
    acc3<K == 1>(s1.vg, sg * cross(s2.p, s0.p));
    acc3<K == 2>(s2.vg, sg * cross(s0.p, s1.p));
  }
}

// Sums of the event rows of one owned vertex, in row order.
struct VertexSumsA {
  d3 K;
  double va, ve;
  d3 vg;
};

template <bool BENDING, bool VG>
MS_HD VertexSumsA vertex_sums_a(const LocalA& l, int e0, int e1) {
  VertexSumsA r;
  r.K = make_d3(0, 0, 0);
  r.va = r.ve = 0.0;
  r.vg = make_d3(0, 0, 0);
  for (int e = e0; e < e1; ++e) {
    if (BENDING) {
      const double* o = l.evA + 5 * e;
      r.K = r.K + make_d3(o[0], o[1], o[2]);
      r.va += o[3];
      r.ve += o[4];
    }
    if (VG) r.vg = r.vg + ld3(l.evV, e);
  }
  return r;
}

// Area-weighted vertex normal of owned vertex i (bending_utils.py:13-34): only needed where an
// interior vertex has |K| <= 1e-15 (bending.py:154-158), i.e. on flat regions.  Scans the
// patch's compact facet records in order (fixed summation order).
MS_HD d3 vertex_normal_scan(const FacetRec* recs, int n_fac, const double* pos, int i) {
  d3 n = make_d3(0, 0, 0);
  for (int k = 0; k < n_fac; ++k) {
    const FacetRec rec = recs[k];
    if (rec.a != i && rec.b != i && rec.c != i) continue;
    const d3 v0 = ld3(pos, rec.a), v1 = ld3(pos, rec.b), v2 = ld3(pos, rec.c);
    n = n + cross(v1 - v0, v2 - v0);
  }
  const double m = sqrt(dot(n, n));
  if (m > 1.0e-15) n = (1.0 / m) * n;
  return n;
}

// Vertex stage of an owned vertex (bending.py:112-158).  normal_of() supplies the unit
// area-weighted vertex normal; it is only called for flat interior vertices.
template <class NormalFn>
MS_HD VertexSeed vertex_body_a(const VertexSumsA& a, bool boundary, NormalFn normal_of, double kappa, double c0,
                               bool willmore) {
  d3 n = make_d3(0, 0, 0);
  if (!(sqrt(dot(a.K, a.K)) > 1.0e-15) && !boundary) n = normal_of();
  return vertex_stage(a.K, a.va, a.ve, kappa, willmore ? 0.0 : c0, boundary, willmore, n, 0.0);
}

// ---------------------------------------------------------------------------
// Pass B
// ---------------------------------------------------------------------------
struct SlotB {
  d3 p;            // position
  d3 f;            // seeds: fK, fA_eff, fA_vor
  double fe, fv;
  int32_t bnd;
  double t2;
  d3 g;            // partial sums: shape gradient, six times dV/dx, barycentric area
  d3 vg;
  double ab;
};

struct LocalB {
  const double* pos;   // [L][3]
  const double* seed;  // [L][5]  bending only
  const int32_t* bfl;  // [L]     bending only; nullptr = closed mesh
  const double* t2;    // [L]     tilt only
  double* evG;         // [E][3]
  double* evV;         // [E][3]
  double* evT;         // [E]
};

MS_HD void slot_clear_b(SlotB& s) {
  s.g = make_d3(0, 0, 0);
  s.vg = make_d3(0, 0, 0);
  s.ab = 0.0;
}

template <bool BENDING, bool ZERO = true>
MS_HD void slot_load_b(SlotB& s, const LocalB& l, int i) {
  s.p = ld3(l.pos, i);
  if (BENDING) {
    const double* q = l.seed + 5 * i;
    s.f = make_d3(q[0], q[1], q[2]);
    s.fe = q[3];
    s.fv = q[4];
    s.bnd = l.bfl ? l.bfl[i] : 0;
  } else {
    s.f = make_d3(0, 0, 0);
    s.fe = s.fv = 0.0;
    s.bnd = 0;
  }
  s.t2 = l.t2 ? l.t2[i] : 0.0;
  if (ZERO) slot_clear_b(s);
}

template <bool VG, bool TILT>
MS_HD void slot_flush_b(const SlotB& s, const LocalB& l, int e) {
  st3(l.evG, e, s.g);
  if (VG) st3(l.evV, e, s.vg);
  if (TILT) l.evT[e] = s.ab;
}

// scalars_here: pass A did not run (no bending), so the per-facet scalars are summed here.
// K: the slot this step loaded (see step_compute_a).
template <bool BENDING, bool VG, bool TILT, int K = -1>
MS_HD void step_compute_b(SlotB& s0, SlotB& s1, SlotB& s2, uint32_t word, double gam, uint32_t modules, uint32_t flags,
                          double k_tilt, bool scalars_here, double* sums) {
  const FacetGeom g = facet_geom(s0.p, s1.p, s2.p);
  const bool primary = (word & STEP_PRIMARY) != 0;
  const bool body = (modules & MS_MOD_VOLUME) && (word & STEP_BODY);
  const double sgn = (word & STEP_NEG) ? -1.0 : 1.0;
  const double T = 0.5 * g.S;
  if (!(modules & MS_MOD_SURFACE)) gam = 0.0;
  double coeff = 0.0;
  if (TILT) {
    coeff = 0.5 * k_tilt * ((s0.t2 + s1.t2 + s2.t2) / 3.0);
    const double Ts = g.S >= kSurfaceSkip ? T : 0.0;
    sums[PS_E_TILT] += primary ? coeff * Ts : 0.0;
    const double third = Ts / 3.0;
    acc1<K == 0>(s0.ab, third); acc1<K == 1>(s1.ab, third); acc1<K == 2>(s2.ab, third);
  }
  d3 c12 = make_d3(0, 0, 0);
  if ((modules & MS_MOD_VOLUME) && (VG || scalars_here)) c12 = (body ? sgn : 0.0) * cross(s1.p, s2.p);
  if (scalars_here) {
    const double Tp = primary ? T : 0.0;
    sums[PS_AREA] += Tp;
    if (modules & MS_MOD_SURFACE) sums[PS_E_SURFACE] += g.S >= kSurfaceSkip ? gam * Tp : 0.0;
    if (modules & MS_MOD_VOLUME) sums[PS_VOLUME6] += primary ? dot(c12, s0.p) : 0.0;
  }
  BendIn b;
  if (BENDING) {
    b.f0 = s0.f; b.fe0 = s0.fe; b.fv0 = s0.fv;
    b.f1 = s1.f; b.fe1 = s1.fe; b.fv1 = s1.fv;
    b.f2 = s2.f; b.fe2 = s2.fe; b.fv2 = s2.fv;
    b.i0 = !s0.bnd; b.i1 = !s1.bnd; b.i2 = !s2.bnd;
  } else {
    b.f0 = b.f1 = b.f2 = make_d3(0, 0, 0);
    b.fe0 = b.fe1 = b.fe2 = b.fv0 = b.fv1 = b.fv2 = 0.0;
    b.i0 = b.i1 = b.i2 = true;
  }
  const CornerG cg = facet_pass_b<BENDING>(g, gam, coeff, b, (flags & MS_FLAG_APPROX) != 0);
  acc3<K == 0>(s0.g, cg.g0);
  acc3<K == 1>(s1.g, cg.g1);
  acc3<K == 2>(s2.g, cg.g2);
  if (VG) {  // unscaled: the vertex sums multiply by 1/6; facets outside the body contribute zeros
    const double sg = body ? sgn : 0.0;
    acc3<K == 0>(s0.vg, c12);
    acc3<K == 1>(s1.vg, sg * cross(s2.p, s0.p));
    acc3<K == 2>(s2.vg, sg * cross(s0.p, s1.p));
  }
}

struct VertexSumsB {
  d3 g, vg;
  double ab;
};

template <bool VG, bool TILT>
MS_HD VertexSumsB vertex_sums_b(const LocalB& l, int e0, int e1) {
  VertexSumsB r;
  r.g = make_d3(0, 0, 0);
  r.vg = make_d3(0, 0, 0);
  r.ab = 0.0;
  for (int e = e0; e < e1; ++e) {
    r.g = r.g + ld3(l.evG, e);
    if (VG) r.vg = r.vg + ld3(l.evV, e);
    if (TILT) r.ab += l.evT[e];
  }
  return r;
}

// ---------------------------------------------------------------------------
// One lane's walk over its steps (host emulator and reference for the device loop).
// load(slot, index), flush(slot, event), compute(word, word position) are callables; `words`
// points at the lane's first word, consecutive steps are `lanes` words apart; the three tail
// rows and the restart rows follow (ms_pack.h).
// ---------------------------------------------------------------------------
template <class LoadFn, class FlushFn, class ComputeFn>
MS_HD void walk_lane(const uint32_t* words, int lanes, int n_steps, LoadFn load, FlushFn flush, ComputeFn compute) {
  const uint32_t* aux = words + size_t(n_steps + 3) * size_t(lanes);
  for (int s = 0; s < n_steps; ++s) {
    const uint32_t w = words[size_t(s) * size_t(lanes)];
    const int k = s % 3;
    if (w & STEP_RESTART) {
      for (int i = 1; i <= 2; ++i) {
        const uint32_t a = *aux;
        aux += lanes;
        const int kk = (k + i) % 3;
        if (step_event(a) >= 0) flush(kk, step_event(a));
        load(kk, step_index(a));
      }
    }
    if (step_event(w) >= 0) flush(k, step_event(w));
    if (w & STEP_LOAD) load(k, step_index(w));
    if (w & STEP_COMPUTE) compute(w, s);
  }
  for (int k = 0; k < 3; ++k) {
    const uint32_t w = words[size_t(n_steps + k) * size_t(lanes)];
    if (step_event(w) >= 0) flush(k, step_event(w));
  }
}

}  // namespace ms
