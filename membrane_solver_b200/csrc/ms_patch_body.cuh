// Per-record and per-vertex bodies of the patch kernels, written against
// patch-LOCAL structure-of-arrays so that the device kernels (shared memory,
// compile-time strides) and the test-only host emulator (tests/emul/emul.cpp, heap
// arrays, run-time strides) execute the very same code.
//
// Layouts.  Patch-local INPUTS (positions, seeds) are arrays of structures, exactly as they lie in global
// memory: row i of the positions at base[3 i .. 3 i + 2], of the seeds at base[5 i .. 5 i + 4] -- so the owned
// rows of a patch arrive with ONE bulk copy per array, and a vertex is reached from one address register plus
// immediate offsets.  A half-warp that gathers rows with distinct indices mod 16 is free of bank conflicts
// (64-bit accesses, 3 and 5 are coprime to 16).  ACCUMULATORS are structure-of-arrays: component k of owned vertex
// i at base[k * stride + i].
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
constexpr int kDumpRows = 16;       // accumulator rows that swallow halo-corner contributions

// Strides of the patch-local arrays.  Device: compile-time constants; emulator: run-time.
template <int LS, int AS>
struct StaticStrides {
  static constexpr int L = LS;  // per-local-vertex inputs (positions, seeds, flags)
  static constexpr int A = AS;  // owned-vertex accumulators (+ kDumpRows)
};
struct DynamicStrides {
  int L, A;
};

// array-of-structures row i of an (n,3) array (global memory, stateless kernels)
MS_HD d3 ld3(const double* p, int i) { return make_d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
MS_HD d3 ld3s(const double* p, int stride, int i) {
  return make_d3(p[i], p[stride + i], p[2 * stride + i]);
}
// row i of a patch-local (n,3) input array
MS_HD d3 ld3a(const double* p, int i) { return make_d3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
MS_HD void add3s(double* p, int stride, int i, d3 v) {
  p[i] += v.x;
  p[stride + i] += v.y;
  p[2 * stride + i] += v.z;
}

// Accumulator row of a corner: owned vertices have their own row; halo corners go to one of
// kDumpRows scratch rows chosen by the SAME residue the gather used, so a half-warp that
// gathers without bank conflicts also accumulates without them, with no branch.
// The dump rows are the LAST kDumpRows rows of the accumulator stride (a multiple of 16).
MS_HD int acc_row(int local, int P, int A) { return local < P ? local : (A - kDumpRows) + (local & (kDumpRows - 1)); }

struct LocalA {
  const double* pos;   // L x 3 (rows)
  const int32_t* bfl;  // L  boundary flags; nullptr = closed mesh (no boundary vertex)
  const double* t2;    // L  |tilt|^2 (tilt module only)
  double* acc;         // 5 x A: K.x K.y K.z A_vor A_eff
  int P;               // owned vertices; local indices >= P are halo (read-only)
};

// Per-facet scalars (primary listing only) + pass-A corner contributions, in registers.
template <class S>
MS_HD CornerA facet_compute_a(const S& st, FacetRec rec, double gam, const LocalA& s, uint32_t modules,
                              double k_tilt, double* sums) {
  const d3 v0 = ld3a(s.pos, rec.a), v1 = ld3a(s.pos, rec.b), v2 = ld3a(s.pos, rec.c);
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

// Read-modify-write of the three corners' accumulator rows.  All loads are issued before the
// first store: the three rows are distinct for owned corners (facets naming a vertex twice are
// never listed, ms_pack.cpp), and a collision can only happen between two halo corners sharing a
// dump row, whose content is never read -- so the loads of one corner need not wait for the stores
// of another (shortest possible token hold time).
template <int N>
MS_HD void rmw3(double* base, int stride, int ia, int ib, int ic, const double (&ca)[N], const double (&cb)[N],
                const double (&cc)[N]) {
  double xa[N], xb[N], xc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    xa[k] = base[k * stride + ia];
    xb[k] = base[k * stride + ib];
    xc[k] = base[k * stride + ic];
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    base[k * stride + ia] = xa[k] + ca[k];
    base[k * stride + ib] = xb[k] + cb[k];
    base[k * stride + ic] = xc[k] + cc[k];
  }
}

template <class S>
MS_HD void facet_accumulate_a(const S& st, FacetRec rec, const CornerA& c, const LocalA& s, uint32_t modules) {
  if (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT)) {
    const int ia = acc_row(rec.a, s.P, st.A), ib = acc_row(rec.b, s.P, st.A), ic = acc_row(rec.c, s.P, st.A);
    const double ca[5] = {c.K0.x, c.K0.y, c.K0.z, c.va0, c.ve0};
    const double cb[5] = {c.K1.x, c.K1.y, c.K1.z, c.va1, c.ve1};
    const double cc[5] = {c.K2.x, c.K2.y, c.K2.z, c.va2, c.ve2};
    rmw3<5>(s.acc, st.A, ia, ib, ic, ca, cb, cc);
  }
}

// Area-weighted vertex normal of owned vertex i (bending_utils.py:13-34): only needed where an
// interior vertex has |K| <= 1e-15 (bending.py:154-158), i.e. on flat regions.  Scans the
// patch's records in slot order (fixed summation order).
template <class S>
MS_HD d3 vertex_normal_scan(const S& st, const FacetRec* recs, int n_slots, const double* pos, int i) {
  d3 n = make_d3(0, 0, 0);
  for (int k = 0; k < n_slots; ++k) {
    const FacetRec rec = recs[k];
    if (!(rec.flags & REC_VALID)) continue;
    if (rec.a != i && rec.b != i && rec.c != i) continue;
    const d3 v0 = ld3a(pos, rec.a), v1 = ld3a(pos, rec.b), v2 = ld3a(pos, rec.c);
    n = n + cross(v1 - v0, v2 - v0);
  }
  const double m = sqrt(dot(n, n));
  if (m > 1.0e-15) n = (1.0 / m) * n;
  return n;
}

// Vertex stage of owned vertex i (bending.py:112-158).  normal_of(i) supplies the unit
// area-weighted vertex normal; it is only called for flat interior vertices.
template <class S, class NormalFn>
MS_HD VertexSeed vertex_body_a(const S& st, int i, const LocalA& s, bool boundary, NormalFn normal_of, double kappa,
                               double c0, bool willmore) {
  const d3 K = ld3s(s.acc, st.A, i);
  d3 n = make_d3(0, 0, 0);
  if (!(dot(K, K) > 1.0e-30) && !boundary) n = normal_of(i);
  return vertex_stage(K, s.acc[3 * st.A + i], s.acc[4 * st.A + i], kappa, willmore ? 0.0 : c0, boundary,
                      willmore, n, 0.0);
}

struct LocalB {
  const double* pos;   // L x 3 (rows)
  const double* seed;  // L x 5 (rows)   bending only: fK.x fK.y fK.z fA_eff fA_vor
  const int32_t* bfl;  // L       bending only; nullptr = closed mesh
  const double* t2;    // L       tilt only
  double* acc;         // 6 x A: shape gradient (3), 6 dV/dx (3)
  double* accAb;       // A       barycentric vertex area (tilt gradient)
  int P;
};

// Results of one facet in pass B, held in registers between compute and accumulation.
struct FacetOutB {
  CornerG cg;     // shape gradient contributions
  CornerG vg;     // 6 dV/dx contributions (valid when in_body)
  double third;   // T/3 for the barycentric area (tilt), 0 when the facet is skipped
  bool in_body;
};

template <bool BENDING, class S>
MS_HD FacetOutB facet_compute_b(const S& st, FacetRec rec, double gam, const LocalB& s, uint32_t modules,
                                uint32_t flags, double k_tilt, bool scalars_here, double* sums) {
  FacetOutB o;
  const d3 v0 = ld3a(s.pos, rec.a), v1 = ld3a(s.pos, rec.b), v2 = ld3a(s.pos, rec.c);
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
    const double* q = s.seed;
    const double* q0 = q + kSeedStrideBody * rec.a;
    const double* q1 = q + kSeedStrideBody * rec.b;
    const double* q2 = q + kSeedStrideBody * rec.c;
    b.f0 = make_d3(q0[0], q0[1], q0[2]); b.fe0 = q0[3]; b.fv0 = q0[4];
    b.f1 = make_d3(q1[0], q1[1], q1[2]); b.fe1 = q1[3]; b.fv1 = q1[4];
    b.f2 = make_d3(q2[0], q2[1], q2[2]); b.fe2 = q2[3]; b.fv2 = q2[4];
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
  if (o.in_body) o.vg = facet_volume_grad6(v0, v1, v2);  // unscaled: the epilogue multiplies by 1/6
  return o;
}

template <class S>
MS_HD void facet_accumulate_b(const S& st, FacetRec rec, const FacetOutB& o, const LocalB& s, bool do_volume,
                              bool do_tilt) {
  const int ia = acc_row(rec.a, s.P, st.A), ib = acc_row(rec.b, s.P, st.A), ic = acc_row(rec.c, s.P, st.A);
  {
    const double ca[3] = {o.cg.g0.x, o.cg.g0.y, o.cg.g0.z};
    const double cb[3] = {o.cg.g1.x, o.cg.g1.y, o.cg.g1.z};
    const double cc[3] = {o.cg.g2.x, o.cg.g2.y, o.cg.g2.z};
    rmw3<3>(s.acc, st.A, ia, ib, ic, ca, cb, cc);
  }
  if (do_volume && o.in_body) {
    const double ca[3] = {o.vg.g0.x, o.vg.g0.y, o.vg.g0.z};
    const double cb[3] = {o.vg.g1.x, o.vg.g1.y, o.vg.g1.z};
    const double cc[3] = {o.vg.g2.x, o.vg.g2.y, o.vg.g2.z};
    rmw3<3>(s.acc + 3 * st.A, st.A, ia, ib, ic, ca, cb, cc);
  }
  if (do_tilt && o.third != 0.0) {
    s.accAb[ia] += o.third;
    s.accAb[ib] += o.third;
    s.accAb[ic] += o.third;
  }
}

}  // namespace ms
