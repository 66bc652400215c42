// sm_100a kernels of the energy+gradient path.
//
// Patch kernels (pass A / pass B): PERSISTENT CTAs, one per SM, each walking its share of the
// vertex patches; the facets of a patch are walked as lane-local triangle strips (ms_pack.h,
// ms_patch_body.cuh) -- see k_patch below.
//
// Reference functions replaced: see the header of ms_math.cuh.
#include "ms_kernels.cuh"

#include <cstdlib>

#include "ms_bt.cuh"
#include "ms_patch_body.cuh"

namespace ms {

namespace {

static_assert(kSeedStride == kSeedStrideBody, "seed row layout");
static_assert(int(SC_E_BENDING_TILT) == int(PS_E_BENDING_TILT) && int(SC_G_G) == int(PS_G_G) &&
                  int(SC_GC_GC) == int(PS_GC_GC) && kPartialStride == PS_COUNT, "partial layout");

// Deterministic block sum of N values per thread over the first n_threads threads (a multiple
// of 32) synchronised through named barrier `bar`; result valid in thread 0.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* red, int n_threads, int bar) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = n_threads >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n_threads) : "memory");  // red may still be read
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[warp * N + k] = v[k];
  }
  asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n_threads) : "memory");
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double x = (lane < n_warps) ? red[lane * N + k] : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
      v[k] = x;
    }
  }
}

// ---- asynchronous copies: LDGSTS (cp.async) for scattered rows, bulk copies for contiguous ranges ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return unsigned(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// One contiguous global range -> shared memory (16-byte aligned on both sides, size a multiple of 16);
// completion is counted in bytes on the mbarrier.
__device__ __forceinline__ void bulk_copy(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- mbarriers (producer <-> consumers) and named barriers (one per team) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned addr, unsigned parity) {
  unsigned ok = 0;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  while (!mbar_try(addr, parity)) {
  }
}
// for the producer warp: back off between polls so that the spin does not take issue slots from the consumers
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  while (!mbar_try(addr, parity)) __nanosleep(200);
}
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---------------------------------------------------------------------------
// Shared-memory plan (run-time sizes: the largest patch of the packed mesh).  A CTA holds
//   n_buf staging buffers   positions (+ seeds) of the patch-local vertices, event offsets, header
//   n_teams event buffers   one per team of consumer warps
// Every section starts on a 16-byte boundary.
// ---------------------------------------------------------------------------
struct PatchHdrS {  // header of the staged patch, written by the producer
  int32_t v_lo, n_owned, n_halo, n_steps;
  int32_t n_events, n_fac, halo_off, reserved;
  int64_t step_off, fac_off;
};

struct PlanFlags {
  bool seed, bfl, t2;       // staging arrays besides the positions
  bool evA, evV, evG, evT;  // event arrays
};

struct Plan {
  // offsets inside one staging buffer
  unsigned oPos, oSeed, oPtr, oWords, oBfl, oT2, oHdr, in_bytes;
  // offsets inside one event buffer
  unsigned oEvA, oEvV, oEvG, oEvT, ev_bytes;
  // whole window
  unsigned oEv, oBars, oRed, total;
};

__host__ __device__ inline unsigned up16(unsigned x) { return (x + 15u) & ~15u; }

__host__ __device__ inline Plan make_plan(const PlanFlags& f, int max_local, int max_owned, int max_events, int max_words,
                                          int n_buf, int n_teams, int n_warps) {
  Plan p;
  const unsigned L = unsigned(max_local + 2), E = unsigned(max_events + 2);
  unsigned o = 0;
  p.oPos = o; o += up16(L * 24u);
  p.oSeed = o; if (f.seed) o += up16(L * 40u);
  p.oPtr = o; o += up16(unsigned(max_owned + 9) * 2u);
  p.oWords = o; o += up16(unsigned(max_words) * 4u);
  p.oBfl = o; if (f.bfl) o += up16(L * 4u);
  p.oT2 = o; if (f.t2) o += up16(L * 8u);
  p.oHdr = o; o += 64u;
  p.in_bytes = o;
  o = 0;
  p.oEvA = o; if (f.evA) o += up16(E * 40u);
  p.oEvV = o; if (f.evV) o += up16(E * 24u);
  p.oEvG = o; if (f.evG) o += up16(E * 24u);
  p.oEvT = o; if (f.evT) o += up16(E * 8u);
  p.ev_bytes = o;
  p.oEv = unsigned(n_buf) * p.in_bytes;
  p.oBars = p.oEv + unsigned(n_teams) * p.ev_bytes;
  p.oRed = p.oBars + 16u * unsigned(n_buf);
  p.total = p.oRed + unsigned(n_warps + 1) * PS_COUNT * 8u;
  return p;
}

constexpr int kMaxBuffers = 6;

// Area-weighted unit normal of owned vertex i of a patch, read from GLOBAL memory (rare path: flat interior
// vertices only).  Same facet order as vertex_normal_scan.
__device__ d3 vertex_normal_global(const PatchLaunch& a, const PatchHdrS& hs, int i) {
  const int32_t* ids = a.halo_ids + hs.halo_off;
  const FacetRec* recs = a.recs + hs.fac_off;
  d3 n = make_d3(0, 0, 0);
  for (int k = 0; k < hs.n_fac; ++k) {
    const FacetRec rec = recs[k];
    if (rec.a != i && rec.b != i && rec.c != i) continue;
    const int ra = rec.a < hs.n_owned ? hs.v_lo + rec.a : ids[rec.a - hs.n_owned];
    const int rb = rec.b < hs.n_owned ? hs.v_lo + rec.b : ids[rec.b - hs.n_owned];
    const int rc = rec.c < hs.n_owned ? hs.v_lo + rec.c : ids[rec.c - hs.n_owned];
    const d3 v0 = ld3(a.pos, ra), v1 = ld3(a.pos, rb), v2 = ld3(a.pos, rc);
    n = n + cross(v1 - v0, v2 - v0);
  }
  const double m = sqrt(dot(n, n));
  if (m > 1.0e-15) n = (1.0 / m) * n;
  return n;
}

template <int K, class S>
__device__ __forceinline__ S& pick(S& s0, S& s1, S& s2) {
  return K == 0 ? s0 : (K == 1 ? s1 : s2);
}

// Per-launch constants of the consumer loops.
struct StepCtx {
  uint32_t modules, flags;
  double gamma_u, k_tilt;
  const double* step_gamma;  // per step word, or nullptr
  bool scalars_here;
};

// ---- pass A: one step on slot K.  `aux` walks the lane's restart rows. ----
template <int K, bool BEND, bool VG>
__device__ __forceinline__ void step_a(uint32_t w, const uint32_t*& aux, int lanes, SlotA& s0, SlotA& s1, SlotA& s2,
                                       const LocalA& loc, const StepCtx& cx, const double* gam_p, double* sums) {
  if (w & STEP_RESTART) {
    const uint32_t ax0 = aux[0], ax1 = aux[lanes];
    aux += 2 * lanes;
    SlotA& n1 = pick<(K + 1) % 3>(s0, s1, s2);
    SlotA& n2 = pick<(K + 2) % 3>(s0, s1, s2);
    if (step_event(ax0) >= 0) slot_flush_a<BEND, VG>(n1, loc, step_event(ax0));
    slot_load_a(n1, loc, step_index(ax0));
    if (step_event(ax1) >= 0) slot_flush_a<BEND, VG>(n2, loc, step_event(ax1));
    slot_load_a(n2, loc, step_index(ax1));
  }
  SlotA& me = pick<K>(s0, s1, s2);
  if (step_event(w) >= 0) slot_flush_a<BEND, VG>(me, loc, step_event(w));
  if (w & STEP_COMPUTE) {  // every step that evaluates a facet also loads its slot (ms_pack.cpp); others are no-ops
    slot_load_a<false>(me, loc, step_index(w));
    const double gam = gam_p ? *gam_p : cx.gamma_u;
    step_compute_a<BEND, VG, K>(s0, s1, s2, w, gam, cx.modules, cx.k_tilt, sums);
  }
}

template <int K, bool BEND, bool VG, bool TILT>
__device__ __forceinline__ void step_b(uint32_t w, const uint32_t*& aux, int lanes, SlotB& s0, SlotB& s1, SlotB& s2,
                                       const LocalB& loc, const StepCtx& cx, const double* gam_p, double* sums) {
  if (w & STEP_RESTART) {
    const uint32_t ax0 = aux[0], ax1 = aux[lanes];
    aux += 2 * lanes;
    SlotB& n1 = pick<(K + 1) % 3>(s0, s1, s2);
    SlotB& n2 = pick<(K + 2) % 3>(s0, s1, s2);
    if (step_event(ax0) >= 0) slot_flush_b<VG, TILT>(n1, loc, step_event(ax0));
    slot_load_b<BEND>(n1, loc, step_index(ax0));
    if (step_event(ax1) >= 0) slot_flush_b<VG, TILT>(n2, loc, step_event(ax1));
    slot_load_b<BEND>(n2, loc, step_index(ax1));
  }
  SlotB& me = pick<K>(s0, s1, s2);
  if (step_event(w) >= 0) slot_flush_b<VG, TILT>(me, loc, step_event(w));
  if (w & STEP_COMPUTE) {  // every step that evaluates a facet also loads its slot (ms_pack.cpp); others are no-ops
    slot_load_b<BEND, false>(me, loc, step_index(w));
    const double gam = gam_p ? *gam_p : cx.gamma_u;
    step_compute_b<BEND, VG, TILT, K>(s0, s1, s2, w, gam, cx.modules, cx.flags, cx.k_tilt, cx.scalars_here, sums);
  }
}

// The strip walk of one lane over one patch (device form of walk_lane, ms_patch_body.cuh).  `words` points at
// the lane's column of the patch's step words in SHARED memory (staged by the producer with the rest of the
// patch: the consumers never wait for global memory); n_steps is a multiple of 3 (ms_pack.cpp pads with no-op
// rows).  The word of step s + 1 is read while step s is evaluated.
template <bool BEND, bool VG>
__device__ __forceinline__ void walk_a(const uint32_t* words, int lanes, int n_steps, const LocalA& loc, const StepCtx& cx,
                                       const double* gam, double* sums) {
  SlotA s0, s1, s2;
  s0.p = s1.p = s2.p = make_d3(0, 0, 0);
  s0.bnd = s1.bnd = s2.bnd = 0;
  s0.t2 = s1.t2 = s2.t2 = 0.0;
  slot_clear_a(s0); slot_clear_a(s1); slot_clear_a(s2);
  const uint32_t* aux = words + (n_steps + 3) * lanes;
  uint32_t w0 = words[0];
  for (int s = 0; s < n_steps; s += 3) {
    const uint32_t* row = words + s * lanes;
    const double* g = gam ? gam + size_t(s) * lanes : nullptr;
    const uint32_t w1 = row[lanes];
    step_a<0, BEND, VG>(w0, aux, lanes, s0, s1, s2, loc, cx, g, sums);
    const uint32_t w2 = row[2 * lanes];
    step_a<1, BEND, VG>(w1, aux, lanes, s0, s1, s2, loc, cx, g ? g + lanes : nullptr, sums);
    w0 = row[3 * lanes];
    step_a<2, BEND, VG>(w2, aux, lanes, s0, s1, s2, loc, cx, g ? g + 2 * lanes : nullptr, sums);
  }
  // tail rows: w0 holds row n_steps
  const uint32_t t1 = words[(n_steps + 1) * lanes], t2 = words[(n_steps + 2) * lanes];
  if (step_event(w0) >= 0) slot_flush_a<BEND, VG>(s0, loc, step_event(w0));
  if (step_event(t1) >= 0) slot_flush_a<BEND, VG>(s1, loc, step_event(t1));
  if (step_event(t2) >= 0) slot_flush_a<BEND, VG>(s2, loc, step_event(t2));
}

template <bool BEND, bool VG, bool TILT>
__device__ __forceinline__ void walk_b(const uint32_t* words, int lanes, int n_steps, const LocalB& loc, const StepCtx& cx,
                                       const double* gam, double* sums) {
  SlotB s0, s1, s2;
  s0.p = s1.p = s2.p = make_d3(0, 0, 0);
  s0.f = s1.f = s2.f = make_d3(0, 0, 0);
  s0.fe = s1.fe = s2.fe = s0.fv = s1.fv = s2.fv = 0.0;
  s0.bnd = s1.bnd = s2.bnd = 0;
  s0.t2 = s1.t2 = s2.t2 = 0.0;
  slot_clear_b(s0); slot_clear_b(s1); slot_clear_b(s2);
  const uint32_t* aux = words + (n_steps + 3) * lanes;
  uint32_t w0 = words[0];
  for (int s = 0; s < n_steps; s += 3) {
    const uint32_t* row = words + s * lanes;
    const double* g = gam ? gam + size_t(s) * lanes : nullptr;
    const uint32_t w1 = row[lanes];
    step_b<0, BEND, VG, TILT>(w0, aux, lanes, s0, s1, s2, loc, cx, g, sums);
    const uint32_t w2 = row[2 * lanes];
    step_b<1, BEND, VG, TILT>(w1, aux, lanes, s0, s1, s2, loc, cx, g ? g + lanes : nullptr, sums);
    w0 = row[3 * lanes];
    step_b<2, BEND, VG, TILT>(w2, aux, lanes, s0, s1, s2, loc, cx, g ? g + 2 * lanes : nullptr, sums);
  }
  const uint32_t t1 = words[(n_steps + 1) * lanes], t2 = words[(n_steps + 2) * lanes];
  if (step_event(w0) >= 0) slot_flush_b<VG, TILT>(s0, loc, step_event(w0));
  if (step_event(t1) >= 0) slot_flush_b<VG, TILT>(s1, loc, step_event(t1));
  if (step_event(t2) >= 0) slot_flush_b<VG, TILT>(s2, loc, step_event(t2));
}

// ---------------------------------------------------------------------------
// The persistent patch kernel.  PASS 0 = pass A (curvature accumulation, per-facet scalars,
// vertex stage -> seeds, and dV/dx when a gradient evaluation with bending follows; alone it is the
// energy-only evaluation of the line search); PASS 1 = pass B (shape gradient of surface + bending
// (+ tilt magnitude); dV/dx when pass A did not run).
//
// One CTA per SM walks patches blockIdx.x, +gridDim.x, ...  Warp roles: the last warp is the PRODUCER:
// it stages patch j into staging buffer j % n_buf -- one bulk copy per contiguous array (owned
// position rows, owned seed rows, event offsets), 8-byte asynchronous copies for the halo rows -- and
// hands it over through a full/empty mbarrier pair.  The other warps form n_teams TEAMS; team t takes
// the CTA's patches t, t + n_teams, ...: its lanes walk their strip pieces (registers only, event rows
// are plain stores), meet at the team's named barrier, then each lane sums the event rows of its
// owned vertices in row order and writes their output rows; a second barrier frees the team's event
// buffer.  While one team is in its epilogue or waits for a patch, the others keep the fp64 pipe busy.
//
// KIND fixes parts of the configuration at compile time: 1 = headline (closed mesh, uniform
// gamma / kappa / c0, surface + Helfrich analytic + volume), 2 = surface and/or volume only,
// 3 = bending with run-time parameters but no tilt, 0 = everything at run time.
// ---------------------------------------------------------------------------
constexpr uint32_t kFastModules = MS_MOD_SURFACE | MS_MOD_BENDING | MS_MOD_VOLUME;

template <int PASS, int KIND>
__global__ void __launch_bounds__(kPatchThreads, 1) k_patch(const PatchLaunch a, const PlanFlags pf, const Plan pl, bool bending_b,
                                                            bool scalars_here_arg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int lanes = a.threads;                 // lanes of one team
  const int n_teams = a.teams;
  const int n_buf = n_teams + 1;
  const int n_cons = lanes * n_teams;          // consumer threads; the producer warp follows
  const int n_my = a.patch_count > int(blockIdx.x) ? (a.patch_count - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x) : 0;

  constexpr bool FAST = KIND == 1 || KIND == 2;
  const uint32_t modules = KIND == 1 ? kFastModules
                           : KIND == 2 ? (a.modules & (MS_MOD_SURFACE | MS_MOD_VOLUME)) : a.modules;
  const uint32_t flags = FAST ? 0u : a.flags;
  const bool do_tilt = KIND == 0 && (modules & MS_MOD_TILT) && a.tilts != nullptr;
  const bool has_boundary = !FAST && a.is_boundary != nullptr;
  const bool do_bending = (KIND == 1 || KIND == 3) ? true
                          : KIND == 2 ? false
                          : (PASS == 0 ? (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT)) != 0 : bending_b);
  const bool do_volume = KIND == 1 ? true : (modules & MS_MOD_VOLUME) != 0;
  const bool scalars_here = (KIND == 1 || KIND == 3) ? false : scalars_here_arg;
  // dV/dx: produced by pass A when it runs ahead of a gradient evaluation, else by pass B
  const bool vg_here = PASS == 0 ? (do_bending && do_volume && a.volgrad_in_a != 0 && a.volgrad != nullptr)
                                 : (!do_bending && do_volume && a.volgrad != nullptr);
  const bool want_epi = PASS == 1 || do_bending;  // pass A without bending accumulates nothing

  // pf / pl: the staging and event arrays of this launch and their offsets (computed by the host: plan_flags, make_plan)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + pl.oBars);
  uint64_t* bar_empty = bar_full + n_buf;
  double* red = reinterpret_cast<double*>(smem + pl.oRed);
  if (tid == 0) {
    for (int b = 0; b < n_buf; ++b) {
      mbar_init(&bar_full[b], 32);
      mbar_init(&bar_empty[b], unsigned(lanes / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  double sums[PS_COUNT];
#pragma unroll
  for (int k = 0; k < PS_COUNT; ++k) sums[k] = 0.0;

  if (tid >= n_cons) {
    // =========================== producer warp ===========================
    // Software pipelined by one patch: while the copies of patch j are in flight, the header and the halo
    // ids of patch j + 1 are fetched into registers, so that a freed staging buffer is refilled after ONE
    // memory latency (no header -> ids -> rows chain in front of the copies).
    const int lane = tid - n_cons;
    constexpr int kIds = 10;  // halo ids per lane held in registers (320 per patch); more are read in place
    auto patch_of = [&](int j) {
      const int pidx = a.patch_begin + int(blockIdx.x) + j * int(gridDim.x);
      return a.patch_list ? a.patch_list[pidx] : pidx;
    };
    PatchHeader h;
    int64_t words_end = 0;
    int32_t ids[kIds];
    auto fetch = [&](int j) {
      const int pid = patch_of(j);
      h = a.patches[pid];
      words_end = a.patches[pid + 1].step_off;  // a sentinel header closes the last patch
      const int32_t* hsrc = a.halo_ids + h.halo_off;
#pragma unroll
      for (int q = 0; q < kIds; ++q) {
        const int k = lane + 32 * q;
        ids[q] = k < h.n_halo ? hsrc[k] : 0;
      }
    };
    if (n_my > 0) fetch(0);
    for (int j = 0; j < n_my; ++j) {
      const int b = j % n_buf;
      if (j >= n_buf) mbar_wait_relaxed(&bar_empty[b], unsigned((j / n_buf - 1) & 1));
      unsigned char* in = smem + size_t(b) * pl.in_bytes;
      double* pos = reinterpret_cast<double*>(in + pl.oPos);
      double* seed = reinterpret_cast<double*>(in + pl.oSeed);
      const int Pn = h.n_owned, Hn = h.n_halo, v_lo = h.v_lo;
      const int32_t* hsrc = a.halo_ids + h.halo_off;
      // rows [0, Pb) of the owned range move as bulk copies (16-byte aligned start, even row count)
      const int Pb = (v_lo & 1) ? 0 : (Pn & ~1);
      if (lane == 0) {
        PatchHdrS* hs = reinterpret_cast<PatchHdrS*>(in + pl.oHdr);
        hs->v_lo = h.v_lo; hs->n_owned = h.n_owned; hs->n_halo = h.n_halo; hs->n_steps = h.n_steps;
        hs->n_events = h.n_events; hs->n_fac = h.n_fac; hs->halo_off = h.halo_off; hs->reserved = 0;
        hs->step_off = h.step_off; hs->fac_off = h.fac_off;
        const unsigned ptr_bytes = up16(unsigned(Pn + 1) * 2u);
        const unsigned word_bytes = unsigned(words_end - h.step_off) * 4u;
        unsigned tx = ptr_bytes + word_bytes + unsigned(Pb) * 24u;
        if (pf.seed) tx += unsigned(Pb) * 40u;
        mbar_expect_tx(&bar_full[b], tx);
        bulk_copy(in + pl.oPtr, a.evt_ptr + h.evt_off, ptr_bytes, &bar_full[b]);
        if (word_bytes) bulk_copy(in + pl.oWords, a.steps + h.step_off, word_bytes, &bar_full[b]);
        if (Pb) {
          bulk_copy(pos, a.pos + size_t(v_lo) * 3, unsigned(Pb) * 24u, &bar_full[b]);
          if (pf.seed) bulk_copy(seed, a.seeds + size_t(v_lo) * kSeedStride, unsigned(Pb) * 40u, &bar_full[b]);
        }
      }
      int32_t* bf = pf.bfl ? reinterpret_cast<int32_t*>(in + pl.oBfl) : nullptr;
      double* t2 = pf.t2 ? reinterpret_cast<double*>(in + pl.oT2) : nullptr;
      auto stage_row = [&](int i, size_t row) {  // one vertex row with 8-byte asynchronous copies
        const double* prow = a.pos + row * 3;
        cp_async8(pos + 3 * i, prow);
        cp_async8(pos + 3 * i + 1, prow + 1);
        cp_async8(pos + 3 * i + 2, prow + 2);
        if (pf.seed) {
          const double* srow = a.seeds + row * kSeedStride;
#pragma unroll
          for (int c = 0; c < kSeedStride; ++c) cp_async8(seed + kSeedStride * i + c, srow + c);
        }
        if (bf) cp_async4(bf + i, a.boundary32 + row);
        if (t2) cp_async8(t2 + i, a.tilt_sq + row);
      };
#pragma unroll
      for (int q = 0; q < kIds; ++q) {  // halo rows whose ids are already in registers
        const int k = lane + 32 * q;
        if (k < Hn) stage_row(Pn + k, size_t(ids[q]));
      }
      for (int k = lane + 32 * kIds; k < Hn; k += 32) stage_row(Pn + k, size_t(hsrc[k]));
      for (int i = Pb + lane; i < Pn; i += 32) stage_row(i, size_t(v_lo) + i);  // owned rows outside the bulk copy
      if (bf || t2) {  // flags (int32) and |t|^2 of the bulk-copied rows
        for (int i = lane; i < Pb; i += 32) {
          if (bf) cp_async4(bf + i, a.boundary32 + v_lo + i);
          if (t2) cp_async8(t2 + i, a.tilt_sq + v_lo + i);
        }
      }
      if (j + 1 < n_my) fetch(j + 1);  // in flight while the copies land
      cp_async_wait_all();
      mbar_arrive(&bar_full[b]);
    }
  } else {
    // ============================= consumers =============================
    const int team = tid / lanes, lt = tid - team * lanes;
    const int bar_id = 1 + team;
    const bool willmore = (flags & MS_FLAG_WILLMORE) != 0;
    StepCtx cx;
    cx.modules = modules;
    cx.flags = flags;
    cx.gamma_u = a.gamma_u;
    cx.k_tilt = a.k_tilt;
    cx.step_gamma = (!FAST && a.step_gamma) ? a.step_gamma : nullptr;
    cx.scalars_here = scalars_here;
    unsigned char* evb = smem + pl.oEv + size_t(team) * pl.ev_bytes;
    for (int j = team; j < n_my; j += n_teams) {
      const int b = j % n_buf;
      mbar_wait(&bar_full[b], unsigned((j / n_buf) & 1));
      unsigned char* in = smem + size_t(b) * pl.in_bytes;
      const PatchHdrS& hs = *reinterpret_cast<const PatchHdrS*>(in + pl.oHdr);  // fields are read where they are needed
      const uint16_t* ptr = reinterpret_cast<const uint16_t*>(in + pl.oPtr);
      const uint32_t* words = reinterpret_cast<const uint32_t*>(in + pl.oWords) + lt;
      const double* gam = cx.step_gamma ? cx.step_gamma + hs.step_off + lt : nullptr;
      const int Pn = hs.n_owned;
      if (PASS == 0) {
        LocalA la;
        la.pos = reinterpret_cast<const double*>(in + pl.oPos);
        la.bfl = pf.bfl ? reinterpret_cast<const int32_t*>(in + pl.oBfl) : nullptr;
        la.t2 = pf.t2 ? reinterpret_cast<const double*>(in + pl.oT2) : nullptr;
        la.evA = reinterpret_cast<double*>(evb + pl.oEvA);
        la.evV = reinterpret_cast<double*>(evb + pl.oEvV);
        if (do_bending) {
          if (vg_here) walk_a<true, true>(words, lanes, hs.n_steps, la, cx, gam, sums);
          else walk_a<true, false>(words, lanes, hs.n_steps, la, cx, gam, sums);
        } else {
          walk_a<false, false>(words, lanes, hs.n_steps, la, cx, gam, sums);
        }
        if (want_epi) {
          named_sync(bar_id, lanes);
          // vertex stage + seeds (+ dV/dx rows) of the owned vertices
          for (int i = lt; i < Pn; i += lanes) {
            const size_t row = size_t(hs.v_lo) + i;
            const VertexSumsA vs = vg_here ? vertex_sums_a<true, true>(la, ptr[i], ptr[i + 1])
                                           : vertex_sums_a<true, false>(la, ptr[i], ptr[i + 1]);
            const double kap = (!FAST && a.kappa) ? a.kappa[row] : a.kappa_u;
            const double c0 = (!FAST && a.c0) ? a.c0[row] : a.c0_u;
            const bool on_boundary = has_boundary && a.is_boundary[row] != 0;
            auto normal_fn = [&]() { return vertex_normal_global(a, hs, i); };
            const VertexSeed sd = vertex_body_a(vs, on_boundary, normal_fn, kap, c0, willmore);
            sums[PS_E_BENDING] += sd.E;
            if (a.seeds) {
              double* o = a.seeds + row * kSeedStride;
              o[0] = sd.fK.x; o[1] = sd.fK.y; o[2] = sd.fK.z; o[3] = sd.fAe; o[4] = sd.fAv;
            }
            if (vg_here) st3(a.volgrad, row, (1.0 / 6.0) * vs.vg);
            if (!FAST) {
              if (a.k_vecs) st3(a.k_vecs, row, vs.K);
              if (a.a_vor) a.a_vor[row] = vs.va;
              if (a.a_eff) a.a_eff[row] = vs.ve;
              if (a.e_vertex) a.e_vertex[row] = sd.E;
            }
          }
          named_sync(bar_id, lanes);  // every event row has been read: the next patch may overwrite them
        }
      } else {
        LocalB lb;
        lb.pos = reinterpret_cast<const double*>(in + pl.oPos);
        lb.seed = reinterpret_cast<const double*>(in + pl.oSeed);
        lb.bfl = pf.bfl ? reinterpret_cast<const int32_t*>(in + pl.oBfl) : nullptr;
        lb.t2 = pf.t2 ? reinterpret_cast<const double*>(in + pl.oT2) : nullptr;
        lb.evG = reinterpret_cast<double*>(evb + pl.oEvG);
        lb.evV = reinterpret_cast<double*>(evb + pl.oEvV);
        lb.evT = reinterpret_cast<double*>(evb + pl.oEvT);
        if (KIND == 0) {
          if (do_bending) {
            if (do_tilt) walk_b<true, false, true>(words, lanes, hs.n_steps, lb, cx, gam, sums);
            else walk_b<true, false, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
          } else if (vg_here) {
            if (do_tilt) walk_b<false, true, true>(words, lanes, hs.n_steps, lb, cx, gam, sums);
            else walk_b<false, true, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
          } else {
            if (do_tilt) walk_b<false, false, true>(words, lanes, hs.n_steps, lb, cx, gam, sums);
            else walk_b<false, false, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
          }
        } else if (KIND == 2) {
          if (vg_here) walk_b<false, true, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
          else walk_b<false, false, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
        } else {
          walk_b<true, false, false>(words, lanes, hs.n_steps, lb, cx, gam, sums);
        }
        named_sync(bar_id, lanes);
        // dV/dx rows written by pass A (bending evaluations): issued ahead of the event sums that hide their latency
        // (not ahead of the barrier: a register reload in front of it would wait for these loads)
        constexpr int kPre = 3;
        d3 vg_pre[kPre];
        const bool vg_from_a = !vg_here && do_volume && a.volgrad != nullptr;
        if (vg_from_a) {
#pragma unroll
          for (int q = 0; q < kPre; ++q) {
            const int i = lt + q * lanes;
            vg_pre[q] = i < Pn ? ld3(a.volgrad + 3 * size_t(hs.v_lo), i) : make_d3(0, 0, 0);
          }
        }
        // gradient (+ dV/dx) rows of the owned vertices and the three KKT dot products
        auto vertex_b = [&](int i, d3& g_out) {
          const size_t row = size_t(hs.v_lo) + i;
          VertexSumsB vs;
          if (vg_here) vs = do_tilt ? vertex_sums_b<true, true>(lb, ptr[i], ptr[i + 1]) : vertex_sums_b<true, false>(lb, ptr[i], ptr[i + 1]);
          else vs = do_tilt ? vertex_sums_b<false, true>(lb, ptr[i], ptr[i + 1]) : vertex_sums_b<false, false>(lb, ptr[i], ptr[i + 1]);
          st3(a.grad, row, vs.g);
          sums[PS_G_G] += dot(vs.g, vs.g);
          g_out = vs.g;
          if (vg_here) {
            const d3 vg = (1.0 / 6.0) * vs.vg;
            st3(a.volgrad, row, vg);
            sums[PS_G_GC] += dot(vs.g, vg);
            sums[PS_GC_GC] += dot(vg, vg);
          }
          if (do_tilt && a.tilt_grad) {  // tilt.py:163-170: dE/dt_v = k_t t_v A_bary(v)
            a.tilt_grad[3 * row] = a.k_tilt * a.tilts[3 * row] * vs.ab;
            a.tilt_grad[3 * row + 1] = a.k_tilt * a.tilts[3 * row + 1] * vs.ab;
            a.tilt_grad[3 * row + 2] = a.k_tilt * a.tilts[3 * row + 2] * vs.ab;
          }
        };
        d3 g_pre[kPre];
#pragma unroll
        for (int q = 0; q < kPre; ++q) {
          const int i = lt + q * lanes;
          g_pre[q] = make_d3(0, 0, 0);
          if (i < Pn) vertex_b(i, g_pre[q]);
        }
        if (vg_from_a) {
#pragma unroll
          for (int q = 0; q < kPre; ++q) {  // rows beyond the patch carry zeros on both sides
            sums[PS_G_GC] += dot(g_pre[q], vg_pre[q]);
            sums[PS_GC_GC] += dot(vg_pre[q], vg_pre[q]);
          }
        }
        for (int i = lt + kPre * lanes; i < Pn; i += lanes) {  // patches with more than kPre vertices per lane
          d3 g;
          vertex_b(i, g);
          if (vg_from_a) {
            const d3 vg = ld3(a.volgrad + 3 * size_t(hs.v_lo), i);
            sums[PS_G_GC] += dot(g, vg);
            sums[PS_GC_GC] += dot(vg, vg);
          }
        }
        named_sync(bar_id, lanes);
      }
      // leaving the patch: its staging buffer may be refilled
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(&bar_empty[b]);
    }
  }

  block_sum<PS_COUNT>(sums, red, n_cons + 32, 15);
  if (tid == 0) {
    double* p = a.partials + (size_t(a.partial_row0) + blockIdx.x) * kPartialStride;
#pragma unroll
    for (int k = 0; k < PS_COUNT; ++k) p[k] = sums[k];
  }
}

// Fixed-order reduction of the per-CTA partial sums of pass A and pass B (one CTA of
// 64 x 12 threads: thread (row, k) sums slot k of rows row, row+64, ... and 12 threads then
// add the 64 row sums in index order).  Slot k comes from pass B's rows when bit k of
// b_mask is set, else from pass A's.
constexpr int kReduceRows = 64;
__global__ void __launch_bounds__(kReduceRows* kPartialStride)
    k_reduce_partials(const double* __restrict__ pa, int rows_a, const double* __restrict__ pb, int rows_b,
                      unsigned b_mask, double* scalars) {
  __shared__ double part[kReduceRows][kPartialStride];
  const int k = threadIdx.x % kPartialStride, row = threadIdx.x / kPartialStride;
  const bool from_b = (b_mask >> k) & 1u;
  const double* src = from_b ? pb : pa;
  const int count = from_b ? rows_b : rows_a;
  double v = 0.0;
  for (int p = row; p < count; p += kReduceRows) v += src[size_t(p) * kPartialStride + k];
  part[row][k] = v;
  __syncthreads();
  if (threadIdx.x < kPartialStride) {
    double t = 0.0;
    for (int r = 0; r < kReduceRows; ++r) t += part[r][threadIdx.x];
    scalars[threadIdx.x] = (threadIdx.x == PS_VOLUME6) ? t / 6.0 : t;
  }
}

// ---------------------------------------------------------------------------
// KKT projection helpers (runtime/constraint_manager.py:294-301).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dots(const double* __restrict__ g,
                                              const double* __restrict__ gc, int64_t n,
                                              double* block_partials) {
  __shared__ double red[32 * 3];
  double v[3] = {0.0, 0.0, 0.0};
  // fixed chunk per block -> fixed summation order
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = per * blockIdx.x, hi = (lo + per < n) ? lo + per : n;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double a = g[i], b = gc ? gc[i] : 0.0;
    v[0] += a * a;
    v[1] += a * b;
    v[2] += b * b;
  }
  block_sum<3>(v, red, 256, 0);
  if (threadIdx.x == 0) {
    block_partials[3 * blockIdx.x] = v[0];
    block_partials[3 * blockIdx.x + 1] = v[1];
    block_partials[3 * blockIdx.x + 2] = v[2];
  }
}

__global__ void __launch_bounds__(256) k_dots_final(const double* __restrict__ block_partials,
                                                    int n_blocks, double* scalars) {
  __shared__ double red[32 * 3];
  double v[3] = {0.0, 0.0, 0.0};
  for (int p = threadIdx.x; p < n_blocks; p += blockDim.x) {
    v[0] += block_partials[3 * p];
    v[1] += block_partials[3 * p + 1];
    v[2] += block_partials[3 * p + 2];
  }
  block_sum<3>(v, red, 256, 0);
  if (threadIdx.x == 0) {
    scalars[SC_G_G] = v[0];
    scalars[SC_G_GC] = v[1];
    scalars[SC_GC_GC] = v[2];
  }
}

__global__ void __launch_bounds__(256) k_project(double* g, const double* __restrict__ gc,
                                                 const uint8_t* __restrict__ fixed, int64_t nv,
                                                 double* scalars, int mode, double k_vol,
                                                 double v_target) {
  double coef = 0.0;
  if (gc) {
    if (mode == 0) {
      const double den = scalars[SC_GC_GC];
      coef = den > 1.0e-18 ? -(scalars[SC_G_GC] / den) : 0.0;
    } else {
      coef = k_vol * (scalars[SC_VOLUME] - v_target);
    }
  }
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) scalars[SC_LAMBDA] = (mode == 0) ? -coef : coef;
  if (i >= 3 * nv) return;
  double x = g[i];
  if (gc) x += coef * gc[i];
  if (fixed && fixed[i / 3]) x = 0.0;
  g[i] = x;
}

__global__ void __launch_bounds__(256) k_gather_rows(const double* __restrict__ src, int width,
                                                     const int32_t* __restrict__ rows, int64_t n,
                                                     double* out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t r = i / width;
  const int c = int(i - r * width);
  out[i] = src[size_t(rows[r]) * width + c];
}

// out[rows[r], :] = src[r, :]  (inverse of k_gather_rows: internal row order -> caller's order)
__global__ void __launch_bounds__(256) k_scatter_rows(const double* __restrict__ src, int width,
                                                      const int32_t* __restrict__ rows, int64_t n,
                                                      double* out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t r = i / width;
  const int c = int(i - r * width);
  out[size_t(rows[r]) * width + c] = src[i];
}

__global__ void __launch_bounds__(256) k_axpy(const double* __restrict__ x,
                                              const double* __restrict__ d, double alpha,
                                              double* out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[i] + alpha * d[i];
}

// ---------------------------------------------------------------------------
// Generic triangle-soup kernels (stateless shims): per-facet pass writes corner
// contributions, per-vertex pass gathers them through the corner CSR in fixed order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool soup_facet(const SoupArgs& s, int f, int& i0, int& i1, int& i2) {
  i0 = s.tri[3 * size_t(f)] + s.shift;
  i1 = s.tri[3 * size_t(f) + 1] + s.shift;
  i2 = s.tri[3 * size_t(f) + 2] + s.shift;
  return i0 >= 0 && i0 < s.nv && i1 >= 0 && i1 < s.nv && i2 >= 0 && i2 < s.nv;
}


__global__ void k_soup_surface(SoupArgs s, const double* __restrict__ gamma, double* corner,
                               double* facet_e) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.nf) return;
  int i0, i1, i2;
  d3 z = make_d3(0, 0, 0);
  CornerG cg; cg.g0 = z; cg.g1 = z; cg.g2 = z;
  double e = 0.0;
  if (soup_facet(s, f, i0, i1, i2)) {
    const FacetGeom g = facet_geom(ld3(s.pos, i0), ld3(s.pos, i1), ld3(s.pos, i2));
    if (g.S >= kSurfaceSkip) {
      BendIn b;
      cg = facet_pass_b<false>(g, gamma[f], 0.0, b, false);
      e = gamma[f] * (0.5 * g.S);
    }
  }
  st3(corner, 3 * size_t(f), cg.g0);
  st3(corner, 3 * size_t(f) + 1, cg.g1);
  st3(corner, 3 * size_t(f) + 2, cg.g2);
  facet_e[f] = e;
}

__global__ void k_soup_volume(SoupArgs s, double factor, double* corner, double* facet_v) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.nf) return;
  int i0, i1, i2;
  d3 z = make_d3(0, 0, 0);
  CornerG cg; cg.g0 = z; cg.g1 = z; cg.g2 = z;
  double v6 = 0.0;
  if (soup_facet(s, f, i0, i1, i2)) {
    const d3 v0 = ld3(s.pos, i0), v1 = ld3(s.pos, i1), v2 = ld3(s.pos, i2);
    cg = facet_volume_grad(v0, v1, v2);
    v6 = facet_volume6(v0, v1, v2);
  }
  st3(corner, 3 * size_t(f), factor * cg.g0);
  st3(corner, 3 * size_t(f) + 1, factor * cg.g1);
  st3(corner, 3 * size_t(f) + 2, factor * cg.g2);
  facet_v[f] = v6;
}

// corner payload: 4 doubles per corner (K.x,K.y,K.z,va)
__global__ void k_soup_curvature(SoupArgs s, double* corner, double* weights, double* va0,
                                 double* va1, double* va2) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.nf) return;
  int i0, i1, i2;
  CornerA c;
  d3 z = make_d3(0, 0, 0);
  c.K0 = z; c.K1 = z; c.K2 = z;
  c.va0 = c.va1 = c.va2 = 0.0;
  c.c0 = c.c1 = c.c2 = 0.0;
  if (soup_facet(s, f, i0, i1, i2)) {
    const FacetGeom g = facet_geom(ld3(s.pos, i0), ld3(s.pos, i1), ld3(s.pos, i2));
    c = facet_pass_a(g, false, false, false);
  }
  double* o = corner + 12 * size_t(f);
  o[0] = c.K0.x; o[1] = c.K0.y; o[2] = c.K0.z; o[3] = c.va0;
  o[4] = c.K1.x; o[5] = c.K1.y; o[6] = c.K1.z; o[7] = c.va1;
  o[8] = c.K2.x; o[9] = c.K2.y; o[10] = c.K2.z; o[11] = c.va2;
  weights[3 * size_t(f)] = c.c0; weights[3 * size_t(f) + 1] = c.c1; weights[3 * size_t(f) + 2] = c.c2;
  if (va0) va0[f] = c.va0;
  if (va1) va1[f] = c.va1;
  if (va2) va2[f] = c.va2;
}

__global__ void k_soup_laplacian(SoupArgs s, int dim, const double* __restrict__ weights,
                                 const double* __restrict__ field, double* corner) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.nf) return;
  int i0, i1, i2;
  const bool ok = soup_facet(s, f, i0, i1, i2);
  const double c0 = weights[3 * size_t(f)], c1 = weights[3 * size_t(f) + 1],
               c2 = weights[3 * size_t(f) + 2];
  double* o = corner + size_t(f) * 3 * dim;
  for (int d = 0; d < dim; ++d) {
    double a = 0.0, b = 0.0, c = 0.0;
    if (ok) {
      const double f0 = field[size_t(i0) * dim + d], f1 = field[size_t(i1) * dim + d],
                   f2 = field[size_t(i2) * dim + d];
      a = 0.5 * (c1 * (f0 - f2) + c2 * (f0 - f1));
      b = 0.5 * (c2 * (f1 - f0) + c0 * (f1 - f2));
      c = 0.5 * (c0 * (f2 - f1) + c1 * (f2 - f0));
    }
    o[d] = a; o[dim + d] = b; o[2 * dim + d] = c;
  }
}

// out[v*out_stride + d] (+)= sum over corners of v of corner[c*in_stride + d], d < dim
__global__ void k_gather(int nv, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                         const double* __restrict__ corner, int in_stride, int in_off, int dim,
                         double* out, int out_stride, int accumulate) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (dim == 3) {  // the common case: one pass over the corner list, three running sums (same order per component)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = ptr[v]; j < ptr[v + 1]; ++j) {
      const double* q = corner + size_t(idx[j]) * in_stride + in_off;
      a0 += q[0];
      a1 += q[1];
      a2 += q[2];
    }
    double* o = out + size_t(v) * out_stride;
    o[0] = accumulate ? o[0] + a0 : a0;
    o[1] = accumulate ? o[1] + a1 : a1;
    o[2] = accumulate ? o[2] + a2 : a2;
    return;
  }
  for (int d = 0; d < dim; ++d) {
    double acc = 0.0;
    for (int j = ptr[v]; j < ptr[v + 1]; ++j) acc += corner[size_t(idx[j]) * in_stride + in_off + d];
    double* o = out + size_t(v) * out_stride + d;
    *o = accumulate ? *o + acc : acc;
  }
}

__global__ void k_grad_cotan(int n, const double* __restrict__ u, const double* __restrict__ v,
                             double* gu, double* gv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  d3 a, b;
  grad_cotan(ld3(u, i), ld3(v, i), a, b);
  st3(gu, i, a);
  st3(gv, i, b);
}

__global__ void k_p1_divergence(SoupArgs s, const double* __restrict__ tilts, double* div,
                                double* area, double* g0, double* g1, double* g2) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= s.nf) return;
  int i0, i1, i2;
  P1 p;
  d3 z = make_d3(0, 0, 0);
  p.g0 = z; p.g1 = z; p.g2 = z; p.div = 0.0; p.area = 0.0;
  if (soup_facet(s, f, i0, i1, i2)) {
    const FacetGeom g = facet_geom(ld3(s.pos, i0), ld3(s.pos, i1), ld3(s.pos, i2));
    p = facet_p1(g, ld3(tilts, i0), ld3(tilts, i1), ld3(tilts, i2));
  }
  div[f] = p.div;
  area[f] = p.area;
  st3(g0, f, p.g0);
  st3(g1, f, p.g1);
  st3(g2, f, p.g2);
}

// barycentric-area-averaged vertex divergence (geometry/tilt_operators.py:414-465): fixed-order gather
__global__ void k_p1_vertex_divergence(SoupArgs s, const double* __restrict__ div, const double* __restrict__ area,
                                       double* div_v, double* area_v) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= s.nv) return;
  double num = 0.0, den = 0.0;
  for (int j = s.csr_ptr[v]; j < s.csr_ptr[v + 1]; ++j) {
    const int f = s.csr_idx[j] / 3;
    const double w = area[f] / 3.0;
    num += w * div[f];
    den += w;
  }
  div_v[v] = den > 1.0e-20 ? num / den : 0.0;
  area_v[v] = den;
}

// single-CTA fixed-order sum (stateless shims only; sizes are modest there)
__global__ void __launch_bounds__(256) k_sum(const double* __restrict__ x, int64_t n, double scale,
                                             double* out) {
  __shared__ double red[32];
  double v[1] = {0.0};
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v[0] += x[i];
  block_sum<1>(v, red, 256, 0);
  if (threadIdx.x == 0) *out = v[0] * scale;
}

// Two-stage fixed-order sum for long arrays: block b sums the contiguous chunk b (threads strided inside
// it) into partial[b]; k_sum then adds the partials in index order.  Same result on every run.
__global__ void __launch_bounds__(256) k_sum_partial(const double* __restrict__ x, int64_t n, double* partial) {
  __shared__ double red[32];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = chunk * blockIdx.x, hi = lo + chunk < n ? lo + chunk : n;
  double v[1] = {0.0};
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) v[0] += x[i];
  block_sum<1>(v, red, 256, 0);
  if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

// ---- bending-tilt coupling (ms_bt.cuh): per-facet / per-vertex passes on global arrays ----
__global__ void __launch_bounds__(128) k_bt_facet_a(BtMesh m, double sign, double* corner4) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < m.nf) bt_facet_a(m, f, sign, corner4);
}

__global__ void __launch_bounds__(128) k_bt_vertex(BtMesh m, const double* __restrict__ k_vecs,
                                                   const double* __restrict__ a_vor,
                                                   const double* __restrict__ a_eff,
                                                   const double* __restrict__ corner4, double* seeds, double* base) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < m.nv) bt_vertex(m, v, k_vecs, a_vor, a_eff, corner4, seeds, base);
}

__global__ void __launch_bounds__(128) k_bt_facet_b(BtMesh m, const double* __restrict__ base, double sign,
                                                    double* corner3, double* facet_e) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < m.nf) facet_e[f] = bt_facet_b(m, f, base, sign, corner3);
}

// scalars[E_BENDING_TILT] <- *e_bt; the pass-A bending energy slot was only a by-product here
__global__ void k_bt_finalize(const double* e_bt, double* scalars) {
  scalars[SC_E_BENDING_TILT] = *e_bt;
  scalars[SC_E_BENDING] = 0.0;
}

// ---- leaflet tilt modules (ms_leaflet.cuh): three sweeps on global arrays + fixed-order gathers ----
__global__ void __launch_bounds__(128) k_lf_facet_a(LeafletMesh m, double* corner) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f < m.nf) lf_facet_a(m, f, corner);
}

__global__ void __launch_bounds__(128) k_lf_vertex(LeafletMesh m, const double* __restrict__ corner, double* vbuf) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < m.nv) lf_vertex(m, v, corner, vbuf);
}

__global__ void __launch_bounds__(128) k_lf_facet_b(LeafletMesh m, const double* __restrict__ vbuf, int with_bt,
                                                    int with_tilt, int with_smooth, double* corner_shape,
                                                    double* corner_tilt,
                                                    double* facet_e /* [e_bt (nf) | e_tilt (nf) | e_smooth (nf)] */) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= m.nf) return;
  const LfEnergies e = lf_facet_b(m, f, vbuf, with_bt != 0, with_tilt != 0, with_smooth != 0, corner_shape, corner_tilt);
  facet_e[f] = e.e_bt;
  facet_e[size_t(m.nf) + f] = e.e_tilt;
  facet_e[2 * size_t(m.nf) + f] = e.e_smooth;
}

// ---- small meshes: the whole leaflet evaluation in ONE cooperative launch ----
// The three sweeps, the energy sums and the gathers of a mesh of a few thousand facets take a few microseconds
// each; as separate launches they cost ~50 us of launch latency and pipeline drain.  Here they are phases of one
// grid (all CTAs resident: cudaLaunchCooperativeKernel) separated by a grid barrier on a 64-bit ticket counter
// that only ever grows (no reset between launches: barrier k of this launch waits for base + k * gridDim.x).
__device__ __forceinline__ void lf_grid_barrier(unsigned long long* ticket, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ticket, 1ull);
    while (*reinterpret_cast<volatile unsigned long long*>(ticket) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) k_lf_fused(LeafletMesh m, int with_bt, int with_tilt, int with_smooth, double* corner,
                                                  double* vbuf, double* corner_shape, double* corner_tilt,
                                                  double* block_e /* 3 * gridDim.x */, double* e_out3, double* grad,
                                                  int accumulate_grad, double* tilt_grad, int accumulate_tilt_grad,
                                                  unsigned long long* ticket, unsigned long long base) {
  __shared__ double red[32 * 3];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  unsigned long long target = base;
  if (with_bt) {
    for (int f = tid; f < m.nf; f += stride) lf_facet_a(m, f, corner);
    lf_grid_barrier(ticket, target += gridDim.x);
    for (int v = tid; v < m.nv; v += stride) lf_vertex(m, v, corner, vbuf);
    lf_grid_barrier(ticket, target += gridDim.x);
  }
  double e[3] = {0.0, 0.0, 0.0};
  for (int f = tid; f < m.nf; f += stride) {
    const LfEnergies r = lf_facet_b(m, f, vbuf, with_bt != 0, with_tilt != 0, with_smooth != 0,
                                    grad ? corner_shape : nullptr, tilt_grad ? corner_tilt : nullptr);
    e[0] += r.e_bt;
    e[1] += r.e_tilt;
    e[2] += r.e_smooth;
  }
  block_sum<3>(e, red, 128, 0);
  if (threadIdx.x == 0) {
    block_e[3 * blockIdx.x] = e[0];
    block_e[3 * blockIdx.x + 1] = e[1];
    block_e[3 * blockIdx.x + 2] = e[2];
  }
  lf_grid_barrier(ticket, target += gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x < 3) {  // fixed order: block 0, 1, 2, ...
    double acc = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) acc += block_e[3 * b + threadIdx.x];
    e_out3[threadIdx.x] = acc;
  }
  for (int v = tid; v < m.nv; v += stride) {
    for (int which = 0; which < 2; ++which) {
      double* out = which == 0 ? grad : tilt_grad;
      if (!out) continue;
      const double* src = which == 0 ? corner_shape : corner_tilt;
      const bool acc = which == 0 ? accumulate_grad != 0 : accumulate_tilt_grad != 0;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int j = m.csr_ptr[v]; j < m.csr_ptr[v + 1]; ++j) {
        const double* q = src + 3 * size_t(m.csr_idx[j]);
        a0 += q[0];
        a1 += q[1];
        a2 += q[2];
      }
      double* o = out + 3 * size_t(v);
      o[0] = acc ? o[0] + a0 : a0;
      o[1] = acc ? o[1] + a1 : a1;
      o[2] = acc ? o[2] + a2 : a2;
    }
  }
}

// Both leaflets of a small mesh in ONE cooperative launch: the work items of the two leaflets share every phase
// (item i < n belongs to the inner leaflet, i >= n to the outer one), so a tilt relaxation iteration costs one
// launch instead of two.  Same per-item code and the same gather order as k_lf_fused: the gradients are bitwise
// those of two separate evaluations (inner first, outer accumulated on top).
struct LfPairBuffers {
  double* corner[2];
  double* vbuf[2];
  double* shape[2];
  double* tilt[2];
  double* e_out3[2];
  double* tilt_grad[2];
};

__global__ void __launch_bounds__(128) k_lf_fused_pair(LeafletMesh m0, LeafletMesh m1, LfPairBuffers b, int with_bt,
                                                       int with_tilt, int with_smooth, double* block_e /* 6 * gridDim.x */,
                                                       double* grad, int accumulate_grad, int accumulate_tilt_grad,
                                                       unsigned long long* ticket, unsigned long long base) {
  __shared__ double red[32 * 6];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  const int nf = m0.nf, nv = m0.nv;
  unsigned long long target = base;
  if (with_bt) {
    for (int i = tid; i < 2 * nf; i += stride) {
      const int l = i >= nf;
      lf_facet_a(l ? m1 : m0, i - l * nf, b.corner[l]);
    }
    lf_grid_barrier(ticket, target += gridDim.x);
    for (int i = tid; i < 2 * nv; i += stride) {
      const int l = i >= nv;
      lf_vertex(l ? m1 : m0, i - l * nv, b.corner[l], b.vbuf[l]);
    }
    lf_grid_barrier(ticket, target += gridDim.x);
  }
  double e[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int i = tid; i < 2 * nf; i += stride) {
    const int l = i >= nf;
    const LfEnergies r = lf_facet_b(l ? m1 : m0, i - l * nf, b.vbuf[l], with_bt != 0, with_tilt != 0, with_smooth != 0,
                                    grad ? b.shape[l] : nullptr, b.tilt_grad[l] ? b.tilt[l] : nullptr);
    e[3 * l] += r.e_bt;
    e[3 * l + 1] += r.e_tilt;
    e[3 * l + 2] += r.e_smooth;
  }
  block_sum<6>(e, red, 128, 0);
  if (threadIdx.x == 0)
    for (int k = 0; k < 6; ++k) block_e[6 * blockIdx.x + k] = e[k];
  lf_grid_barrier(ticket, target += gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x < 6) {  // fixed order: block 0, 1, 2, ...
    double acc = 0.0;
    for (unsigned blk = 0; blk < gridDim.x; ++blk) acc += block_e[6 * blk + threadIdx.x];
    b.e_out3[threadIdx.x / 3][threadIdx.x % 3] = acc;
  }
  for (int i = tid; i < 2 * nv; i += stride) {  // tilt gradients: one array per leaflet
    const int l = i >= nv, v = i - l * nv;
    double* out = b.tilt_grad[l];
    if (!out) continue;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int j = m0.csr_ptr[v]; j < m0.csr_ptr[v + 1]; ++j) {
      const double* q = b.tilt[l] + 3 * size_t(m0.csr_idx[j]);
      a0 += q[0];
      a1 += q[1];
      a2 += q[2];
    }
    double* o = out + 3 * size_t(v);
    o[0] = accumulate_tilt_grad ? o[0] + a0 : a0;
    o[1] = accumulate_tilt_grad ? o[1] + a1 : a1;
    o[2] = accumulate_tilt_grad ? o[2] + a2 : a2;
  }
  if (grad)
    for (int v = tid; v < nv; v += stride) {  // shape gradient: one array, inner leaflet first, outer on top
      double* o = grad + 3 * size_t(v);
      double g0 = accumulate_grad ? o[0] : 0.0, g1 = accumulate_grad ? o[1] : 0.0, g2 = accumulate_grad ? o[2] : 0.0;
      for (int l = 0; l < 2; ++l) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int j = m0.csr_ptr[v]; j < m0.csr_ptr[v + 1]; ++j) {
          const double* q = b.shape[l] + 3 * size_t(m0.csr_idx[j]);
          a0 += q[0];
          a1 += q[1];
          a2 += q[2];
        }
        g0 = (l == 0 && !accumulate_grad) ? a0 : g0 + a0;
        g1 = (l == 0 && !accumulate_grad) ? a1 : g1 + a1;
        g2 = (l == 0 && !accumulate_grad) ? a2 : g2 + a2;
      }
      o[0] = g0;
      o[1] = g1;
      o[2] = g2;
    }
}

// ---- leaflet tilt relaxation helpers (runtime/steppers/tilt_relaxation.py:630-668,894-955) ----
// unit area-weighted vertex normals (Mesh.vertex_normals, geometry/triangle_ops.py:55-72): fixed-order gather
__global__ void __launch_bounds__(128) k_vertex_normals(int32_t nv, const int32_t* __restrict__ tri,
                                                        const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                        const double* __restrict__ pos, double* normals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  d3 n = make_d3(0, 0, 0);
  for (int j = ptr[v]; j < ptr[v + 1]; ++j) {
    const int f = idx[j] / 3;
    const d3 a = ld3(pos, tri[3 * size_t(f)]), b = ld3(pos, tri[3 * size_t(f) + 1]), c = ld3(pos, tri[3 * size_t(f) + 2]);
    n = n + cross(b - a, c - a);
  }
  const double len = sqrt(dot(n, n));
  if (len >= 1.0e-12) n = (1.0 / len) * n;
  st3(normals, v, n);
}

// t -= (t.n) n  (runtime/projections/tilt.py:8-14)
__global__ void __launch_bounds__(256) k_project_tangent(int64_t nv, const double* __restrict__ normals, double* t) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const d3 n = ld3(normals, v), x = ld3(t, v);
  st3(t, v, axpy(-dot(x, n), n, x));
}

// trial = P(t - step g); rows with fixed tilts keep their value (projections/tilt.py:99-138)
__global__ void __launch_bounds__(256) k_tilt_trial(int64_t nv, const double* __restrict__ t, const double* __restrict__ g,
                                                    const double* __restrict__ normals, const uint8_t* __restrict__ fixed,
                                                    double step, double* trial) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const d3 x = ld3(t, v);
  if (fixed && fixed[v]) {
    st3(trial, v, x);
    return;
  }
  const d3 n = ld3(normals, v);
  const d3 y = axpy(-step, ld3(g, v), x);
  st3(trial, v, axpy(-dot(y, n), n, y));
}

// zero the gradient rows of fixed tilts; rowsq[v] = |g_v|^2 of the free rows (tilt_relaxation.py:856-871)
__global__ void __launch_bounds__(256) k_masked_row_norm2(int64_t nv, double* g, const uint8_t* __restrict__ fixed,
                                                          double* rowsq) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (fixed && fixed[v]) {
    st3(g, v, make_d3(0, 0, 0));
    rowsq[v] = 0.0;
    return;
  }
  const d3 x = ld3(g, v);
  rowsq[v] = dot(x, x);
}

// Jacobi preconditioner of the leaflet tilt CG (runtime/preconditioners.py:64-146): diag_v = k_tilt * barycentric
// area (facets of the leaflet when use_keep, else every facet) + 1/2 k_smooth * sum of the two opposite cotangents
// over every facet; diag <= 1e-12 and fixed rows -> 1; the inverse is stored.
__global__ void __launch_bounds__(128) k_leaflet_jacobi(LeafletMesh m, int use_keep, double k_smooth,
                                                        const uint8_t* __restrict__ fixed, double* minv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= m.nv) return;
  double area = 0.0, cot = 0.0;
  for (int j = m.csr_ptr[v]; j < m.csr_ptr[v + 1]; ++j) {
    const int f = m.csr_idx[j] / 3, k = m.csr_idx[j] - 3 * f;
    int idx[3];
    if (!lf_facet_ok(m, f, idx)) continue;
    const FacetGeom g = facet_geom(lf_row(m.pos, idx[0]), lf_row(m.pos, idx[1]), lf_row(m.pos, idx[2]));
    if ((!use_keep || lf_kept(m, f)) && g.S >= kSurfaceSkip) area += 0.5 * g.S / 3.0;
    const CornerA c = facet_pass_a(g, false, false, false);
    cot += k == 0 ? c.c1 + c.c2 : (k == 1 ? c.c2 + c.c0 : c.c0 + c.c1);
  }
  double diag = m.k_tilt * area + 0.5 * k_smooth * cot;
  if (!(diag > 1.0e-12)) diag = 1.0;
  if (fixed && fixed[v]) diag = 1.0;
  minv[v] = 1.0 / diag;
}

// rows[v] = g_v . (minv_v g_v)  (the r.z of the preconditioned CG; minv == nullptr -> identity)
__global__ void __launch_bounds__(256) k_rz_rows(int64_t nv, const double* __restrict__ g, const double* __restrict__ minv,
                                                 double* rows) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const d3 x = ld3(g, v);
  rows[v] = dot(x, x) * (minv ? minv[v] : 1.0);
}

// dir = -minv g + beta dir   (restart: beta == 0 and the old direction is not read)
__global__ void __launch_bounds__(256) k_tilt_cg_direction(int64_t nv, const double* __restrict__ g,
                                                           const double* __restrict__ minv, double beta, int restart,
                                                           double* dir) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const d3 z = (minv ? -minv[v] : -1.0) * ld3(g, v);
  st3(dir, v, restart ? z : axpy(beta, ld3(dir, v), z));
}

// ---- halo exchange over NVLink peer memory ------------------------------------------------------------
// Every rank publishes "my owned rows of this array are written" by storing an epoch number into its flag
// word (k_halo_signal, stream-ordered after the producing kernel).  A consumer waits until every owner's flag
// has reached the epoch it expects and then copies its ghost rows straight out of the owners' arrays with
// peer loads (k_halo_pull): one kernel, no staging buffer, no host round trip.  The wait is bounded: after
// ~2 s without the flag the kernel records an error instead of spinning forever.
__global__ void k_halo_signal(unsigned long long* flag, unsigned long long epoch) {
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(flag) = epoch;
  __threadfence_system();
}

__global__ void __launch_bounds__(256) k_halo_pull(int n_ghost, int width, const double* const* __restrict__ peer_base,
                                                   unsigned long long* const* __restrict__ peer_flag, int n_slots,
                                                   int flag_index, unsigned long long epoch,
                                                   const int32_t* __restrict__ owner, const int32_t* __restrict__ row,
                                                   double* dst, int* error) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    if (!peer_flag[s] || !peer_base[s]) continue;  // only the owners of this rank's ghosts
    const volatile unsigned long long* f = peer_flag[s] + flag_index;
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > 4000000000LL) {
        atomicExch(error, 1);
        ok = 0;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!ok) return;
  __threadfence_system();
  const int total = n_ghost * width;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int g = i / width, k = i - g * width;
    dst[i] = __ldcv(peer_base[owner[g]] + size_t(row[g]) * width + k);
  }
}

// All-reduce (sum) of n <= 16 scalars over peer memory.  Word layout of every rank's exported block:
// [0..3] epoch flags, [8 + 16 p .. 8 + 16 p + 15] scalar slot of parity p.  Publish: copy this rank's scalars into
// the slot of the epoch's parity, then raise flag 2.  Gather: wait for every rank's flag, add the slots in RANK
// ORDER (the same order on every rank, so all ranks hold bitwise the same sums) and store them.  Two slots: a
// rank can publish epoch e+1 only after it has seen every flag of epoch e, and a peer can only lag behind reading
// epoch e, never e-1.
__global__ void k_allreduce_publish(const double* __restrict__ scalars, int n, unsigned long long* words,
                                    unsigned long long epoch) {
  double* slot = reinterpret_cast<double*>(words + 8 + 16 * (epoch & 1ull));
  if (int(threadIdx.x) < n) slot[threadIdx.x] = scalars[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(words + 2) = epoch;
    __threadfence_system();
  }
}

__global__ void k_allreduce_gather(unsigned long long* const* __restrict__ peer_words, int n_slots, int n,
                                   unsigned long long epoch, double* scalars, int* error) {
  __shared__ int ok;
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  for (int s = threadIdx.x; s < n_slots; s += blockDim.x) {
    const volatile unsigned long long* f = peer_words[s] + 2;
    const long long t0 = clock64();
    while (*f < epoch) {
      if (clock64() - t0 > 4000000000LL) {
        atomicExch(error, 1);
        ok = 0;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!ok) return;
  __threadfence_system();
  if (int(threadIdx.x) < n) {
    double acc = 0.0;
    for (int s = 0; s < n_slots; ++s)
      acc += __ldcv(reinterpret_cast<const double*>(peer_words[s] + 8 + 16 * (epoch & 1ull)) + threadIdx.x);
    scalars[threadIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(256) k_row_norm2(const double* __restrict__ rows, int64_t n, double* out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = rows[3 * i], y = rows[3 * i + 1], z = rows[3 * i + 2];
  out[i] = x * x + y * y + z * z;
}

// ---- line-search helpers of the device-resident loop (runtime/steppers/line_search.py:267-541) ----
// minimum edge length squared over the facets (runtime/topology.py:174-199) and maximum row norm
// squared of the search direction; min / max are order independent, so atomics keep determinism.
__global__ void __launch_bounds__(256) k_min_edge2(const int32_t* __restrict__ tri, int32_t nf, int32_t nv,
                                                   const double* __restrict__ pos, unsigned long long* out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 1.0e300;
  if (f < nf) {
    const int i0 = tri[3 * size_t(f)], i1 = tri[3 * size_t(f) + 1], i2 = tri[3 * size_t(f) + 2];
    if (i0 >= 0 && i0 < nv && i1 >= 0 && i1 < nv && i2 >= 0 && i2 < nv) {
      const d3 a = ld3(pos, i0), b = ld3(pos, i1), c = ld3(pos, i2);
      const d3 e0 = c - b, e1 = a - c, e2 = b - a;
      m = fmin(dot(e0, e0), fmin(dot(e1, e1), dot(e2, e2)));
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmin(m, __shfl_down_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMin(out, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(256) k_max_row_norm2(const double* __restrict__ rows, int64_t n,
                                                       unsigned long long* out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (i < n) {
    const double x = rows[3 * i], y = rows[3 * i + 1], z = rows[3 * i + 2];
    m = x * x + y * y + z * z;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// flag |= 1 when a facet normal turns by more than acos(cos_limit) between `old_pos` and `new_pos`, or
// collapses (runtime/topology.py:13-48)
__global__ void __launch_bounds__(256) k_normal_change(const int32_t* __restrict__ tri, int32_t nf, int32_t nv,
                                                       const double* __restrict__ old_pos,
                                                       const double* __restrict__ new_pos, double cos_limit,
                                                       int* flag) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nf) return;
  const int i0 = tri[3 * size_t(f)], i1 = tri[3 * size_t(f) + 1], i2 = tri[3 * size_t(f) + 2];
  if (!(i0 >= 0 && i0 < nv && i1 >= 0 && i1 < nv && i2 >= 0 && i2 < nv)) return;
  const d3 a = ld3(old_pos, i0), b = ld3(old_pos, i1), c = ld3(old_pos, i2);
  const d3 n0 = cross(b - a, c - a);
  const double m0 = sqrt(dot(n0, n0));
  if (!(m0 > 1.0e-12)) return;  // facets that were degenerate before the step are not judged
  const d3 a1 = ld3(new_pos, i0), b1 = ld3(new_pos, i1), c1 = ld3(new_pos, i2);
  const d3 n1 = cross(b1 - a1, c1 - a1);
  const double m1 = sqrt(dot(n1, n1));
  bool bad = m1 < 1.0e-12;
  if (!bad) {
    double d = dot(n0, n1) / (m0 * m1);
    d = fmin(1.0, fmax(-1.0, d));
    bad = !(d >= cos_limit);
  }
  if (bad) atomicOr(flag, 1);
}

// Per-vertex Polak-Ribiere direction (runtime/steppers/conjugate_gradient.py:84-104):
//   beta_v = g_v.(g_v - g'_v) / (g'_v.g'_v + 1e-20);  d_v = -g_v + beta_v d'_v, or -g_v where beta_v < 0;
// fixed rows get a zero direction.
__global__ void __launch_bounds__(256) k_cg_direction(const double* __restrict__ g, const double* __restrict__ pg,
                                                      const double* __restrict__ pd,
                                                      const uint8_t* __restrict__ fixed, int64_t nv, double* d) {
  const int64_t v = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const double gx = g[3 * v], gy = g[3 * v + 1], gz = g[3 * v + 2];
  const double px = pg[3 * v], py = pg[3 * v + 1], pz = pg[3 * v + 2];
  const double numer = gx * (gx - px) + gy * (gy - py) + gz * (gz - pz);
  const double denom = (px * px + py * py + pz * pz) + 1.0e-20;
  const double beta = numer / denom;
  double dx = -gx, dy = -gy, dz = -gz;
  if (!(beta < 0.0)) {
    dx += beta * pd[3 * v]; dy += beta * pd[3 * v + 1]; dz += beta * pd[3 * v + 2];
  }
  if (fixed && fixed[v]) dx = dy = dz = 0.0;
  d[3 * v] = dx; d[3 * v + 1] = dy; d[3 * v + 2] = dz;
}

// y[v,:] += alpha * x[v,:] on the rows that are not fixed (constraint projection of the positions)
__global__ void __launch_bounds__(256) k_axpy_rows(const double* __restrict__ x, double alpha,
                                                   const uint8_t* __restrict__ fixed, int64_t nv, double* y) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 3 * nv) return;
  if (fixed && fixed[i / 3]) return;
  y[i] += alpha * x[i];
}

__global__ void __launch_bounds__(256) k_scale(const double* __restrict__ x, double scale, double* out, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = scale * x[i];
}

inline int blocks_for(int64_t n, int t) { return int((n + t - 1) / t); }

}  // namespace

namespace {
int g_num_sms = 0;

// compile-time kernel kind of a launch (see k_patch)
int kernel_kind(const PatchLaunch& a) {
  const bool diag = a.k_vecs || a.a_vor || a.a_eff || a.e_vertex;
  if (a.modules == kFastModules && a.flags == 0 && !a.is_boundary && !a.step_gamma && !a.kappa && !a.c0 && !diag &&
      a.seeds && a.volgrad)
    return 1;
  if ((a.modules & ~uint32_t(MS_MOD_SURFACE | MS_MOD_VOLUME)) == 0 && a.modules != 0 && !a.step_gamma && !diag &&
      a.volgrad)
    return 2;
  if (!(a.modules & MS_MOD_TILT) && (a.modules & MS_MOD_BENDING) && a.seeds) return 3;  // run-time parameters, no tilt
  return 0;
}

// the staging / event arrays a launch needs: the same decisions as inside k_patch
// template KIND a launch runs with
int launched_kind(int pass, const PatchLaunch& a, bool bending, bool scalars_here) {
  const int kind = kernel_kind(a);
  if (pass == 0) return kind;
  if (kind == 1 && bending && !scalars_here) return 1;
  if (kind == 2 && !bending) return 2;
  if (kind == 3 && bending && !scalars_here) return 3;
  return 0;
}

PlanFlags plan_flags(int pass, int kind, const PatchLaunch& a, bool bending_b) {
  const bool fast = kind == 1 || kind == 2;
  const uint32_t modules = kind == 1 ? kFastModules : kind == 2 ? (a.modules & (MS_MOD_SURFACE | MS_MOD_VOLUME)) : a.modules;
  const bool do_tilt = kind == 0 && (modules & MS_MOD_TILT) && a.tilts != nullptr;
  const bool has_boundary = !fast && a.is_boundary != nullptr;
  const bool do_bending = (kind == 1 || kind == 3) ? true
                          : kind == 2 ? false
                          : (pass == 0 ? (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT)) != 0 : bending_b);
  const bool do_volume = kind == 1 ? true : (modules & MS_MOD_VOLUME) != 0;
  const bool vg_here = pass == 0 ? (do_bending && do_volume && a.volgrad_in_a != 0 && a.volgrad != nullptr)
                                 : (!do_bending && do_volume && a.volgrad != nullptr);
  PlanFlags f;
  f.seed = pass == 1 && do_bending;
  f.bfl = has_boundary;
  f.t2 = do_tilt;
  f.evA = pass == 0 && do_bending;
  f.evV = vg_here;
  f.evG = pass == 1;
  f.evT = pass == 1 && do_tilt;
  return f;
}

size_t patch_smem_bytes(int pass, const PatchLaunch& a, bool bending_b, int teams) {
  const int warps = teams * (a.threads / 32);
  const int kind = launched_kind(pass, a, bending_b, !bending_b);  // the callers sum the scalars in pass B iff pass A does not run
  return make_plan(plan_flags(pass, kind, a, bending_b), a.max_local, a.max_owned, a.max_events, a.max_words, teams + 1, teams, warps + 1).total;
}

// Largest number of teams (of a.threads lanes) a pass can run with: bounded by the warps of a CTA and by the
// shared memory its staging and event buffers need.  0: the patches do not fit at all.
int fit_teams(int pass, const PatchLaunch& a, bool bending) {
  const int w = a.threads / 32;
  if (w <= 0) return 0;
  int teams = (kPatchThreads / 32 - 1) / w;
  if (teams > kMaxBuffers - 1) teams = kMaxBuffers - 1;
  static const int forced = [] { const char* e = std::getenv("MS_TEAMS"); return e ? std::atoi(e) : 0; }();
  if (forced > 0 && forced < teams) teams = forced;
  while (teams > 0 && patch_smem_bytes(pass, a, bending, teams) > size_t(kPatchSmemBytes)) --teams;
  return teams;
}

}  // namespace

int patch_teams(int pass, const PatchLaunch& a, bool bending) { return fit_teams(pass, a, bending); }
size_t patch_smem(int pass, const PatchLaunch& a, bool bending) {
  const int t = fit_teams(pass, a, bending);
  return t > 0 ? patch_smem_bytes(pass, a, bending, t) : 0;
}

cudaError_t configure_kernels() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const void* fns[] = {(const void*)k_patch<0, 0>, (const void*)k_patch<0, 1>, (const void*)k_patch<0, 2>,
                       (const void*)k_patch<0, 3>, (const void*)k_patch<1, 0>, (const void*)k_patch<1, 1>,
                       (const void*)k_patch<1, 2>, (const void*)k_patch<1, 3>};
  for (const void* f : fns) {
    e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, kPatchSmemBytes);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int patch_grid(const PatchLaunch& a) {
  int sms = g_num_sms > 0 ? g_num_sms : 148;
  if (a.max_ctas > 0 && a.max_ctas < sms) sms = a.max_ctas;
  return a.patch_count < sms ? a.patch_count : sms;
}

cudaError_t launch_pass_a(const PatchLaunch& a_in, cudaStream_t st) {
  if (a_in.patch_count <= 0) return cudaSuccess;
  PatchLaunch a = a_in;
  a.teams = fit_teams(0, a, false);
  if (a.teams <= 0) return cudaErrorInvalidConfiguration;
  const size_t smem = patch_smem_bytes(0, a, false, a.teams);
  const int grid = patch_grid(a);
  const int block = a.teams * a.threads + 32;
  const int kind = kernel_kind(a);
  const PlanFlags pf = plan_flags(0, kind, a, false);
  const Plan pl = make_plan(pf, a.max_local, a.max_owned, a.max_events, a.max_words, a.teams + 1, a.teams, a.teams * (a.threads / 32) + 1);
  switch (kind) {
    case 1: k_patch<0, 1><<<grid, block, smem, st>>>(a, pf, pl, true, false); break;
    case 2: k_patch<0, 2><<<grid, block, smem, st>>>(a, pf, pl, false, false); break;
    case 3: k_patch<0, 3><<<grid, block, smem, st>>>(a, pf, pl, false, false); break;
    default: k_patch<0, 0><<<grid, block, smem, st>>>(a, pf, pl, false, false);
  }
  return cudaGetLastError();
}

cudaError_t launch_pass_b(const PatchLaunch& a_in, bool bending, bool scalars_here, cudaStream_t st) {
  if (a_in.patch_count <= 0) return cudaSuccess;
  PatchLaunch a = a_in;
  a.teams = fit_teams(1, a, bending);
  if (a.teams <= 0) return cudaErrorInvalidConfiguration;
  const size_t smem = patch_smem_bytes(1, a, bending, a.teams);
  const int grid = patch_grid(a);
  const int block = a.teams * a.threads + 32;
  const int kind = launched_kind(1, a, bending, scalars_here);
  const PlanFlags pf = plan_flags(1, kind, a, bending);
  const Plan pl = make_plan(pf, a.max_local, a.max_owned, a.max_events, a.max_words, a.teams + 1, a.teams, a.teams * (a.threads / 32) + 1);
  switch (kind) {
    case 1: k_patch<1, 1><<<grid, block, smem, st>>>(a, pf, pl, true, false); break;
    case 2: k_patch<1, 2><<<grid, block, smem, st>>>(a, pf, pl, false, scalars_here); break;
    case 3: k_patch<1, 3><<<grid, block, smem, st>>>(a, pf, pl, bending, scalars_here); break;
    default: k_patch<1, 0><<<grid, block, smem, st>>>(a, pf, pl, bending, scalars_here);
  }
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* pa, int rows_a, const double* pb, int rows_b, unsigned b_mask,
                                   double* scalars, cudaStream_t st) {
  k_reduce_partials<<<1, kReduceRows * kPartialStride, 0, st>>>(pa, rows_a, pb, rows_b, b_mask, scalars);
  return cudaGetLastError();
}

cudaError_t launch_dots(const double* g, const double* gc, int64_t n, double* block_partials,
                        int n_blocks, double* scalars, cudaStream_t st) {
  k_dots<<<n_blocks, 256, 0, st>>>(g, gc, n, block_partials);
  k_dots_final<<<1, 256, 0, st>>>(block_partials, n_blocks, scalars);
  return cudaGetLastError();
}

cudaError_t launch_project(double* g, const double* gc, const uint8_t* fixed, int64_t nv,
                           double* scalars, int mode, double k_vol, double v_target,
                           cudaStream_t st) {
  if (nv <= 0) return cudaSuccess;
  k_project<<<blocks_for(3 * nv, 256), 256, 0, st>>>(g, gc, fixed, nv, scalars, mode, k_vol, v_target);
  return cudaGetLastError();
}

cudaError_t launch_gather_rows(const double* src, int width, const int32_t* rows, int64_t n,
                               double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_gather_rows<<<blocks_for(n * width, 256), 256, 0, st>>>(src, width, rows, n, out);
  return cudaGetLastError();
}

cudaError_t launch_scatter_rows(const double* src, int width, const int32_t* rows, int64_t n,
                                double* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_scatter_rows<<<blocks_for(n * width, 256), 256, 0, st>>>(src, width, rows, n, out);
  return cudaGetLastError();
}

cudaError_t launch_axpy(const double* x, const double* d, double alpha, double* out, int64_t n,
                        cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_axpy<<<blocks_for(n, 256), 256, 0, st>>>(x, d, alpha, out, n);
  return cudaGetLastError();
}

cudaError_t launch_soup_surface(const SoupArgs& s, const double* gamma, double* corner,
                                double* facet_e, double* grad, double* energy_out,
                                cudaStream_t st) {
  if (s.nf > 0) k_soup_surface<<<blocks_for(s.nf, 128), 128, 0, st>>>(s, gamma, corner, facet_e);
  if (s.nv > 0 && s.nf > 0)
    k_gather<<<blocks_for(s.nv, 128), 128, 0, st>>>(s.nv, s.csr_ptr, s.csr_idx, corner, 3, 0, 3, grad, 3, 1);
  k_sum<<<1, 256, 0, st>>>(facet_e, s.nf, 1.0, energy_out);
  return cudaGetLastError();
}

cudaError_t launch_soup_volume(const SoupArgs& s, double factor, double* corner, double* facet_v,
                               double* grad, double* volume_out, cudaStream_t st) {
  if (s.nf > 0) k_soup_volume<<<blocks_for(s.nf, 128), 128, 0, st>>>(s, factor, corner, facet_v);
  if (s.nv > 0 && s.nf > 0 && grad)
    k_gather<<<blocks_for(s.nv, 128), 128, 0, st>>>(s.nv, s.csr_ptr, s.csr_idx, corner, 3, 0, 3, grad, 3, 1);
  k_sum<<<1, 256, 0, st>>>(facet_v, s.nf, 1.0 / 6.0, volume_out);
  return cudaGetLastError();
}

cudaError_t launch_soup_curvature(const SoupArgs& s, double* corner, double* k_vecs,
                                  double* vertex_areas, double* weights, double* va0, double* va1,
                                  double* va2, cudaStream_t st) {
  if (s.nf > 0) k_soup_curvature<<<blocks_for(s.nf, 128), 128, 0, st>>>(s, corner, weights, va0, va1, va2);
  if (s.nv > 0) {
    // with nf == 0 every CSR range is empty and the gathers write zeros
    k_gather<<<blocks_for(s.nv, 128), 128, 0, st>>>(s.nv, s.csr_ptr, s.csr_idx, corner, 4, 0, 3, k_vecs, 3, 0);
    k_gather<<<blocks_for(s.nv, 128), 128, 0, st>>>(s.nv, s.csr_ptr, s.csr_idx, corner, 4, 3, 1, vertex_areas, 1, 0);
  }
  return cudaGetLastError();
}

cudaError_t launch_soup_laplacian(const SoupArgs& s, int32_t dim, const double* weights,
                                  const double* field, double* corner, double* out,
                                  cudaStream_t st) {
  if (s.nf > 0) k_soup_laplacian<<<blocks_for(s.nf, 128), 128, 0, st>>>(s, dim, weights, field, corner);
  if (s.nv > 0)
    k_gather<<<blocks_for(s.nv, 128), 128, 0, st>>>(s.nv, s.csr_ptr, s.csr_idx, corner, dim, 0, dim, out, dim, 0);
  return cudaGetLastError();
}

cudaError_t launch_grad_cotan(int32_t n, const double* u, const double* v, double* gu, double* gv,
                              cudaStream_t st) {
  if (n > 0) k_grad_cotan<<<blocks_for(n, 128), 128, 0, st>>>(n, u, v, gu, gv);
  return cudaGetLastError();
}

cudaError_t launch_p1_divergence(const SoupArgs& s, const double* tilts, double* div, double* area,
                                 double* g0, double* g1, double* g2, cudaStream_t st) {
  if (s.nf > 0) k_p1_divergence<<<blocks_for(s.nf, 128), 128, 0, st>>>(s, tilts, div, area, g0, g1, g2);
  return cudaGetLastError();
}

cudaError_t launch_p1_vertex_divergence(const SoupArgs& s, const double* div, const double* area, double* div_v,
                                       double* area_v, cudaStream_t st) {
  if (s.nv > 0) k_p1_vertex_divergence<<<blocks_for(s.nv, 128), 128, 0, st>>>(s, div, area, div_v, area_v);
  return cudaGetLastError();
}

cudaError_t launch_row_norm2(const double* rows, int64_t n, double* out, cudaStream_t st) {
  if (n > 0) k_row_norm2<<<blocks_for(n, 256), 256, 0, st>>>(rows, n, out);
  return cudaGetLastError();
}

cudaError_t launch_sum(const double* x, int64_t n, double scale, double* out, cudaStream_t st) {
  k_sum<<<1, 256, 0, st>>>(x, n, scale, out);
  return cudaGetLastError();
}

// sum of x[0..n) * scale -> *out; scratch holds kSumBlocks doubles (used for long arrays only)
static void sum_fixed_order(const double* x, int64_t n, double scale, double* out, double* scratch, cudaStream_t st) {
  if (n < (int64_t(1) << 16)) {
    k_sum<<<1, 256, 0, st>>>(x, n, scale, out);
  } else {
    k_sum_partial<<<kSumBlocks, 256, 0, st>>>(x, n, scratch);
    k_sum<<<1, 256, 0, st>>>(scratch, kSumBlocks, scale, out);
  }
}

cudaError_t launch_bt_stage(const BtMesh& m, double sign, const double* k_vecs, const double* a_vor,
                            const double* a_eff, double* corner /* 12*nf doubles */, double* seeds, double* base,
                            double* facet_e, double* e_out, bool tilt_grads, cudaStream_t st) {
  if (m.nf > 0) k_bt_facet_a<<<blocks_for(m.nf, 128), 128, 0, st>>>(m, sign, corner);
  if (m.nv > 0) k_bt_vertex<<<blocks_for(m.nv, 128), 128, 0, st>>>(m, k_vecs, a_vor, a_eff, corner, seeds, base);
  // the corner buffer is free again: it now receives the tilt-gradient contributions (9 per facet)
  if (m.nf > 0) k_bt_facet_b<<<blocks_for(m.nf, 128), 128, 0, st>>>(m, base, sign, tilt_grads ? corner : nullptr, facet_e);
  sum_fixed_order(facet_e, m.nf, 1.0, e_out, e_out + 1, st);
  return cudaGetLastError();
}

cudaError_t launch_bt_tilt_gather(const BtMesh& m, const double* corner3, double* tilt_grad, bool accumulate,
                                  cudaStream_t st) {
  if (m.nv > 0)
    k_gather<<<blocks_for(m.nv, 128), 128, 0, st>>>(m.nv, m.csr_ptr, m.csr_idx, corner3, 3, 0, 3, tilt_grad, 3,
                                                     accumulate ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_bt_finalize(const double* e_bt, double* scalars, cudaStream_t st) {
  k_bt_finalize<<<1, 1, 0, st>>>(e_bt, scalars);
  return cudaGetLastError();
}

cudaError_t launch_leaflet(const LeafletMesh& m, bool with_bt, bool with_tilt, bool with_smooth, double* corner,
                           double* vbuf, double* corner_shape, double* corner_tilt, double* facet_e, double* e_out3,
                           double* sum_scratch, double* grad,
                           bool accumulate_grad, double* tilt_grad, bool accumulate_tilt_grad, cudaStream_t st) {
  if (with_bt) {
    if (m.nf > 0) k_lf_facet_a<<<blocks_for(m.nf, 128), 128, 0, st>>>(m, corner);
    if (m.nv > 0) k_lf_vertex<<<blocks_for(m.nv, 128), 128, 0, st>>>(m, corner, vbuf);
  }
  if (m.nf > 0)
    k_lf_facet_b<<<blocks_for(m.nf, 128), 128, 0, st>>>(m, vbuf, with_bt ? 1 : 0, with_tilt ? 1 : 0, with_smooth ? 1 : 0,
                                                        grad ? corner_shape : nullptr, tilt_grad ? corner_tilt : nullptr,
                                                        facet_e);
  for (int k = 0; k < 3; ++k)
    sum_fixed_order(facet_e + k * size_t(m.nf), m.nf, 1.0, e_out3 + k, sum_scratch + k * kSumBlocks, st);
  if (m.nv > 0 && grad)
    k_gather<<<blocks_for(m.nv, 128), 128, 0, st>>>(m.nv, m.csr_ptr, m.csr_idx, corner_shape, 3, 0, 3, grad, 3,
                                                     accumulate_grad ? 1 : 0);
  if (m.nv > 0 && tilt_grad)
    k_gather<<<blocks_for(m.nv, 128), 128, 0, st>>>(m.nv, m.csr_ptr, m.csr_idx, corner_tilt, 3, 0, 3, tilt_grad, 3,
                                                     accumulate_tilt_grad ? 1 : 0);
  return cudaGetLastError();
}

// Cooperative single-launch variant for small meshes.  Returns cudaErrorNotSupported when the device cannot launch
// cooperatively or the grid would not be resident; the caller then uses launch_leaflet.  *ticket_base is the
// host-side mirror of the ticket counter (advanced by the barriers this launch performs).
cudaError_t launch_leaflet_fused(const LeafletMesh& m, bool with_bt, bool with_tilt, bool with_smooth, double* corner,
                                 double* vbuf, double* corner_shape, double* corner_tilt, double* block_e, int max_blocks,
                                 double* e_out3, double* grad, bool accumulate_grad, double* tilt_grad,
                                 bool accumulate_tilt_grad, unsigned long long* ticket, unsigned long long* ticket_base,
                                 cudaStream_t st) {
  static int resident = -1;
  if (resident < 0) {
    int dev = 0, coop = 0, per_sm = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lf_fused, 128, 0);
    resident = coop ? per_sm * sms : 0;
  }
  int blocks = blocks_for(m.nf > m.nv ? m.nf : m.nv, 128);
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks > resident) blocks = resident;
  if (blocks <= 0) return cudaErrorNotSupported;
  LeafletMesh mm = m;
  int bt = with_bt ? 1 : 0, tl = with_tilt ? 1 : 0, sm = with_smooth ? 1 : 0, ag = accumulate_grad ? 1 : 0,
      at = accumulate_tilt_grad ? 1 : 0;
  unsigned long long base = *ticket_base;
  void* args[] = {&mm, &bt, &tl, &sm, &corner, &vbuf, &corner_shape, &corner_tilt, &block_e, &e_out3, &grad, &ag,
                  &tilt_grad, &at, &ticket, &base};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_lf_fused), dim3(blocks), dim3(128), args, 0, st);
  if (e == cudaSuccess) *ticket_base = base + (unsigned long long)(blocks) * (with_bt ? 3ull : 1ull);
  return e;
}

// Inner and outer leaflet of a small mesh in one cooperative launch (see k_lf_fused_pair); the arrays of index 0
// belong to m0, of index 1 to m1.  cudaErrorNotSupported -> evaluate the leaflets one after the other.
cudaError_t launch_leaflet_fused_pair(const LeafletMesh& m0, const LeafletMesh& m1, bool with_bt, bool with_tilt,
                                      bool with_smooth, double* const corner[2], double* const vbuf[2],
                                      double* const shape[2], double* const tilt[2], double* const e_out3[2],
                                      double* const tilt_grad[2], double* block_e, int max_blocks, double* grad,
                                      bool accumulate_grad, bool accumulate_tilt_grad, unsigned long long* ticket,
                                      unsigned long long* ticket_base, cudaStream_t st) {
  static int resident = -1;
  if (resident < 0) {
    int dev = 0, coop = 0, per_sm = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_lf_fused_pair, 128, 0);
    resident = coop ? per_sm * sms : 0;
  }
  int blocks = blocks_for(2 * int64_t(m0.nf > m0.nv ? m0.nf : m0.nv), 128);
  if (blocks > max_blocks) blocks = max_blocks;
  if (blocks > resident) blocks = resident;
  if (blocks <= 0) return cudaErrorNotSupported;
  LeafletMesh a = m0, b = m1;
  LfPairBuffers buf;
  for (int l = 0; l < 2; ++l) {
    buf.corner[l] = corner[l];
    buf.vbuf[l] = vbuf[l];
    buf.shape[l] = shape[l];
    buf.tilt[l] = tilt[l];
    buf.e_out3[l] = e_out3[l];
    buf.tilt_grad[l] = tilt_grad[l];
  }
  int bt = with_bt ? 1 : 0, tl = with_tilt ? 1 : 0, sm = with_smooth ? 1 : 0, ag = accumulate_grad ? 1 : 0,
      at = accumulate_tilt_grad ? 1 : 0;
  unsigned long long base = *ticket_base;
  void* args[] = {&a, &b, &buf, &bt, &tl, &sm, &block_e, &grad, &ag, &at, &ticket, &base};
  cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_lf_fused_pair), dim3(blocks), dim3(128), args, 0, st);
  if (e == cudaSuccess) *ticket_base = base + (unsigned long long)(blocks) * (with_bt ? 3ull : 1ull);
  return e;
}

cudaError_t launch_vertex_normals(int32_t nv, const int32_t* tri, const int32_t* csr_ptr, const int32_t* csr_idx,
                                  const double* pos, double* normals, cudaStream_t st) {
  if (nv > 0) k_vertex_normals<<<blocks_for(nv, 128), 128, 0, st>>>(nv, tri, csr_ptr, csr_idx, pos, normals);
  return cudaGetLastError();
}

cudaError_t launch_project_tangent(int64_t nv, const double* normals, double* t, cudaStream_t st) {
  if (nv > 0) k_project_tangent<<<blocks_for(nv, 256), 256, 0, st>>>(nv, normals, t);
  return cudaGetLastError();
}

cudaError_t launch_tilt_trial(int64_t nv, const double* t, const double* g, const double* normals, const uint8_t* fixed,
                              double step, double* trial, cudaStream_t st) {
  if (nv > 0) k_tilt_trial<<<blocks_for(nv, 256), 256, 0, st>>>(nv, t, g, normals, fixed, step, trial);
  return cudaGetLastError();
}

cudaError_t launch_masked_norm2(int64_t nv, double* g, const uint8_t* fixed, double* rowsq, double* out,
                                double* sum_scratch /* kSumBlocks */, cudaStream_t st) {
  if (nv > 0) k_masked_row_norm2<<<blocks_for(nv, 256), 256, 0, st>>>(nv, g, fixed, rowsq);
  sum_fixed_order(rowsq, nv, 1.0, out, sum_scratch, st);
  return cudaGetLastError();
}

cudaError_t launch_leaflet_jacobi(const LeafletMesh& m, bool use_keep, double k_smooth, const uint8_t* fixed, double* minv,
                                  cudaStream_t st) {
  if (m.nv > 0) k_leaflet_jacobi<<<blocks_for(m.nv, 128), 128, 0, st>>>(m, use_keep ? 1 : 0, k_smooth, fixed, minv);
  return cudaGetLastError();
}

cudaError_t launch_rz(int64_t nv, const double* g, const double* minv, double* rows, double* out,
                      double* sum_scratch /* kSumBlocks */, cudaStream_t st) {
  if (nv > 0) k_rz_rows<<<blocks_for(nv, 256), 256, 0, st>>>(nv, g, minv, rows);
  sum_fixed_order(rows, nv, 1.0, out, sum_scratch, st);
  return cudaGetLastError();
}

cudaError_t launch_tilt_cg_direction(int64_t nv, const double* g, const double* minv, double beta, bool restart, double* dir,
                                     cudaStream_t st) {
  if (nv > 0) k_tilt_cg_direction<<<blocks_for(nv, 256), 256, 0, st>>>(nv, g, minv, beta, restart ? 1 : 0, dir);
  return cudaGetLastError();
}

cudaError_t launch_halo_signal(unsigned long long* flag, unsigned long long epoch, cudaStream_t st) {
  k_halo_signal<<<1, 1, 0, st>>>(flag, epoch);
  return cudaGetLastError();
}

cudaError_t launch_halo_pull(int n_ghost, int width, const double* const* peer_base, unsigned long long* const* peer_flag,
                             int n_slots, int flag_index, unsigned long long epoch, const int32_t* owner,
                             const int32_t* row, double* dst, int* error, cudaStream_t st) {
  if (n_ghost <= 0) return cudaSuccess;
  const int blocks = (n_ghost * width + 255) / 256;
  k_halo_pull<<<blocks < 64 ? blocks : 64, 256, 0, st>>>(n_ghost, width, peer_base, peer_flag, n_slots, flag_index, epoch,
                                                          owner, row, dst, error);
  return cudaGetLastError();
}

cudaError_t launch_allreduce_peer(double* scalars, int n, unsigned long long* own_words, unsigned long long* const* peer_words,
                                  int n_slots, unsigned long long epoch, int* error, cudaStream_t st) {
  k_allreduce_publish<<<1, 32, 0, st>>>(scalars, n, own_words, epoch);
  k_allreduce_gather<<<1, 64, 0, st>>>(peer_words, n_slots, n, epoch, scalars, error);
  return cudaGetLastError();
}

// First launches load the kernels' code (lazy module loading), which may synchronise the device: do them once,
// with no work, before contexts start waiting for each other.
cudaError_t launch_halo_warmup(unsigned long long* words, double* scalars, int* error, cudaStream_t st) {
  k_halo_signal<<<1, 1, 0, st>>>(words + 3, 0ull);
  k_halo_pull<<<1, 256, 0, st>>>(0, 3, nullptr, nullptr, 0, 0, 0ull, nullptr, nullptr, nullptr, error);
  k_allreduce_publish<<<1, 32, 0, st>>>(scalars, 0, words, 0ull);
  k_allreduce_gather<<<1, 64, 0, st>>>(nullptr, 0, 0, 0ull, scalars, error);
  return cudaGetLastError();
}

cudaError_t launch_min_edge2(const int32_t* tri, int32_t nf, int32_t nv, const double* pos,
                             unsigned long long* out, cudaStream_t st) {
  if (nf > 0) k_min_edge2<<<blocks_for(nf, 256), 256, 0, st>>>(tri, nf, nv, pos, out);
  return cudaGetLastError();
}

cudaError_t launch_max_row_norm2(const double* rows, int64_t n, unsigned long long* out, cudaStream_t st) {
  if (n > 0) k_max_row_norm2<<<blocks_for(n, 256), 256, 0, st>>>(rows, n, out);
  return cudaGetLastError();
}

cudaError_t launch_normal_change(const int32_t* tri, int32_t nf, int32_t nv, const double* old_pos,
                                 const double* new_pos, double cos_limit, int* flag, cudaStream_t st) {
  if (nf > 0) k_normal_change<<<blocks_for(nf, 256), 256, 0, st>>>(tri, nf, nv, old_pos, new_pos, cos_limit, flag);
  return cudaGetLastError();
}

cudaError_t launch_cg_direction(const double* g, const double* pg, const double* pd, const uint8_t* fixed,
                                int64_t nv, double* d, cudaStream_t st) {
  if (nv > 0) k_cg_direction<<<blocks_for(nv, 256), 256, 0, st>>>(g, pg, pd, fixed, nv, d);
  return cudaGetLastError();
}

cudaError_t launch_axpy_rows(const double* x, double alpha, const uint8_t* fixed, int64_t nv, double* y,
                             cudaStream_t st) {
  if (nv > 0) k_axpy_rows<<<blocks_for(3 * nv, 256), 256, 0, st>>>(x, alpha, fixed, nv, y);
  return cudaGetLastError();
}

cudaError_t launch_scale(const double* x, double scale, double* out, int64_t n, cudaStream_t st) {
  if (n > 0) k_scale<<<blocks_for(n, 256), 256, 0, st>>>(x, scale, out, n);
  return cudaGetLastError();
}

}  // namespace ms
