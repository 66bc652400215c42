// sm_100a kernels of the energy+gradient path.
//
// Patch kernels (pass A / pass B): one CTA per vertex patch.  The CTA stages the
// positions (and pass-A vertex results) of its owned + halo vertices in shared
// memory, walks the patch's facet records round by round -- each thread computes
// one facet per round, once, entirely in registers -- and adds the corner
// contributions to shared-memory accumulators of the OWNED vertices with plain
// read-modify-writes.  Rounds are conflict-free by construction (ms_pack.cpp), so
// no atomics are needed and the summation order is fixed at pack time.
//
// Reference functions replaced: see the header of ms_math.cuh.
#include "ms_kernels.cuh"

#include "ms_patch_body.cuh"

namespace ms {

namespace {

constexpr int kMaxThreads = 256;
constexpr int kMaxWarps = kMaxThreads / 32;
static_assert(kSeedStride == kSeedStrideBody, "seed row layout");
static_assert(int(SC_E_BENDING_TILT) == int(PS_E_BENDING_TILT) && int(SC_G_G) == int(PS_G_G) &&
                  int(SC_GC_GC) == int(PS_GC_GC) && kPartialStride == PS_COUNT, "partial layout");

// Deterministic block sum of N values per thread; result valid in thread 0.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
  }
  __syncthreads();  // red may still be read by a previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[warp * N + k] = v[k];
  }
  __syncthreads();
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

// ---------------------------------------------------------------------------
// Staging.  Everything a patch needs is brought into shared memory up front with
// independent, coalesced loads (three dependent latency levels in total: headers ->
// {round table, records, halo ids, owned rows} -> halo rows), so that the round loop
// itself touches shared memory only.
// ---------------------------------------------------------------------------
// Bump allocator over the dynamic shared memory window (16-byte granularity).
struct Carver {
  unsigned char* p;
  __device__ explicit Carver(void* base) : p(static_cast<unsigned char*>(base)) {}
  template <typename T>
  __device__ T* take(size_t count) {
    T* r = reinterpret_cast<T*>(p);
    p += (count * sizeof(T) + 15) / 16 * 16;
    return r;
  }
};
inline size_t carve_bytes(size_t count, size_t elem) { return (count * elem + 15) / 16 * 16; }

// Ampere-style asynchronous copies (LDGSTS): global -> shared without a register
// round trip, so every staging load of a patch is in flight at once.
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(unsigned(__cvta_generic_to_shared(dst))), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void stage_flags(uint8_t* dst, const uint8_t* __restrict__ src,
                                            const PatchHeader& h, const int32_t* halo_local) {
  const int L = h.n_owned + h.n_halo;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int row = j < h.n_owned ? h.v_lo + j : halo_local[j - h.n_owned];
    dst[j] = src ? src[row] : uint8_t(0);
  }
}

__device__ __forceinline__ void stage_tilt_sq(double* dst, const double* __restrict__ tilts,
                                              const PatchHeader& h, const int32_t* halo_local) {
  const int L = h.n_owned + h.n_halo;
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const size_t row = size_t(j < h.n_owned ? h.v_lo + j : halo_local[j - h.n_owned]);
    const double x = tilts[3 * row], y = tilts[3 * row + 1], z = tilts[3 * row + 2];
    dst[j] = x * x + y * y + z * z;
  }
}

// First latency level: records, round table, halo ids (all contiguous per patch).
__device__ __forceinline__ void stage_topology(const PatchLaunch& a, const PatchHeader& h, int n_slots,
                                               FacetRec* recs, int32_t* halo_local) {
  const FacetRec* src = a.recs + h.slot_off;
  for (int j = threadIdx.x; j < n_slots; j += blockDim.x) cp_async8(recs + j, src + j);
  for (int j = threadIdx.x; j < h.n_halo; j += blockDim.x) cp_async4(halo_local + j, a.halo_ids + h.halo_off + j);
}

constexpr int kRedDoubles = kMaxWarps * PS_COUNT;  // block_sum scratch: warps x values

struct SmemA {
  double *pos, *t2, *accK, *accAv, *accAe, *nrm, *red;
  FacetRec* recs;
  int32_t* halo;
  uint8_t* bfl;
  __device__ SmemA(void* base, const PatchLaunch& a, bool tilt) {
    Carver c(base);
    pos = c.take<double>(3 * size_t(a.max_local));
    t2 = c.take<double>(tilt ? a.max_local : 0);
    accK = c.take<double>(5 * size_t(a.max_owned));  // K(3P) | A_vor(P) | A_eff(P), zeroed together
    accAv = accK + 3 * a.max_owned;
    accAe = accAv + a.max_owned;
    nrm = c.take<double>(3 * size_t(a.max_owned));
    red = c.take<double>(kRedDoubles);
    recs = c.take<FacetRec>(a.max_slots);
    halo = c.take<int32_t>(size_t(a.max_local));
    bfl = c.take<uint8_t>(a.max_local);
  }
};

// ---------------------------------------------------------------------------
// Pass A: facet -> owned-vertex accumulation of K, A_vor, A_eff; per-facet scalars
// (surface energy, area, volume, tilt energy); vertex stage -> seeds for pass B
// + bending energy.  Alone, it is the energy-only evaluation of the line search.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads) k_pass_a(PatchLaunch a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pid = a.patch_begin + blockIdx.x;
  const PatchHeader h = a.patches[pid];
  const int n_slots = int(a.patches[pid + 1].slot_off - h.slot_off);  // sentinel header at the end
  const int P = h.n_owned;
  const bool do_tilt = (a.modules & MS_MOD_TILT) && a.tilts != nullptr;
  const bool do_bending = a.modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT);
  SmemA s(smem_raw, a, do_tilt);

  stage_topology(a, h, n_slots, s.recs, s.halo);
  {  // owned rows do not depend on the halo ids: issue them in the same latency level
    const double* owned = a.pos + size_t(h.v_lo) * 3;
    for (int j = threadIdx.x; j < 3 * P; j += blockDim.x) cp_async8(s.pos + j, owned + j);
  }
  for (int j = threadIdx.x; j < 5 * a.max_owned; j += blockDim.x) s.accK[j] = 0.0;
  cp_async_wait_all();
  __syncthreads();
  for (int j = threadIdx.x; j < 3 * h.n_halo; j += blockDim.x) {
    const int v = j / 3, c = j - v * 3;
    cp_async8(s.pos + 3 * P + j, a.pos + size_t(s.halo[v]) * 3 + c);
  }
  if (a.is_boundary) stage_flags(s.bfl, a.is_boundary, h, s.halo);
  if (do_tilt) stage_tilt_sq(s.t2, a.tilts, h, s.halo);
  cp_async_wait_all();
  __syncthreads();

  LocalA loc;
  loc.pos = s.pos; loc.bfl = a.is_boundary ? s.bfl : nullptr; loc.t2 = do_tilt ? s.t2 : nullptr;
  loc.accK = s.accK; loc.accAv = s.accAv; loc.accAe = s.accAe; loc.P = P;
  double sums[PS_COUNT];
#pragma unroll
  for (int k = 0; k < PS_COUNT; ++k) sums[k] = 0.0;

  const double* slot_gamma = a.slot_gamma ? a.slot_gamma + h.slot_off : nullptr;
  const int grp = threadIdx.x / a.threads, lane = threadIdx.x - grp * a.threads;
  for (int r0 = 0; r0 < h.n_rounds; r0 += a.groups) {
    // every group computes one round (reads only), then the groups accumulate in turn
    const int r = r0 + grp;
    bool act = false;
    FacetRec rec;
    CornerA ca;
    if (r < h.n_rounds) {
      const int slot = r * a.threads + lane;
      rec = s.recs[slot];
      if (rec.flags & REC_VALID) {
        act = true;
        const double gam = slot_gamma ? slot_gamma[slot] : a.gamma_u;
        ca = facet_compute_a(rec, gam, loc, a.modules, a.k_tilt, sums);
      }
    }
    for (int g = 0; g < a.groups; ++g) {
      __syncthreads();
      if (act && g == grp) facet_accumulate_a(rec, ca, loc, a.modules);
    }
  }
  __syncthreads();

  if (do_bending) {
    const bool willmore = a.flags & MS_FLAG_WILLMORE;
    int need = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) need |= vertex_needs_normal(loc, i) ? 1 : 0;
    const int any_need = __syncthreads_or(need);
    if (any_need) {
      for (int j = threadIdx.x; j < 3 * P; j += blockDim.x) s.nrm[j] = 0.0;
      __syncthreads();
      for (int r = 0; r < h.n_rounds; ++r) {  // rare path (flat patches): group 0 only
        if (grp == 0) {
          const FacetRec rec = s.recs[r * a.threads + lane];
          if (rec.flags & REC_VALID) normal_body(rec, s.pos, s.nrm, P);
        }
        __syncthreads();
      }
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const size_t row = size_t(h.v_lo) + i;
      const double kap = a.kappa ? a.kappa[row] : a.kappa_u;
      const double c0 = a.c0 ? a.c0[row] : a.c0_u;
      const VertexSeed sd = vertex_body_a(i, loc, s.nrm, any_need != 0, kap, c0, willmore);
      sums[PS_E_BENDING] += sd.E;
      if (a.seeds) {
        double* o = a.seeds + row * kSeedStride;
        o[0] = sd.fK.x; o[1] = sd.fK.y; o[2] = sd.fK.z; o[3] = sd.fAe; o[4] = sd.fAv;
      }
      if (a.k_vecs) {
        a.k_vecs[3 * row] = s.accK[3 * i];
        a.k_vecs[3 * row + 1] = s.accK[3 * i + 1];
        a.k_vecs[3 * row + 2] = s.accK[3 * i + 2];
      }
      if (a.a_vor) a.a_vor[row] = s.accAv[i];
      if (a.a_eff) a.a_eff[row] = s.accAe[i];
      if (a.e_vertex) a.e_vertex[row] = sd.E;
    }
  }

  block_sum<PS_COUNT>(sums, s.red);
  if (threadIdx.x == 0) {
    double* p = a.partials + size_t(pid) * kPartialStride;
#pragma unroll
    for (int k = 0; k < PS_COUNT; ++k) p[k] = sums[k];
  }
}

struct SmemB {
  double *pos, *seed, *t2, *accG, *accV, *accAb, *red;
  FacetRec* recs;
  int32_t* halo;
  uint8_t* bfl;
  __device__ SmemB(void* base, const PatchLaunch& a, bool bending, bool tilt) {
    Carver c(base);
    seed = c.take<double>(bending ? size_t(kSeedStride) * a.max_local : 0);
    pos = c.take<double>(3 * size_t(a.max_local));
    t2 = c.take<double>(tilt ? a.max_local : 0);
    accG = c.take<double>(6 * size_t(a.max_owned));  // grad(3P) | dV/dx(3P), zeroed together
    accV = accG + 3 * a.max_owned;
    accAb = c.take<double>(tilt ? a.max_owned : 0);
    red = c.take<double>(kRedDoubles);
    recs = c.take<FacetRec>(a.max_slots);
    halo = c.take<int32_t>(size_t(a.max_local));
    bfl = c.take<uint8_t>(a.max_local);
  }
};

// ---------------------------------------------------------------------------
// Pass B: shape gradient of surface + bending (+ tilt magnitude) and dV/dx.
// ---------------------------------------------------------------------------
template <bool BENDING>
__global__ void __launch_bounds__(kMaxThreads, 2) k_pass_b(PatchLaunch a, bool scalars_here) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pid = a.patch_begin + blockIdx.x;
  const PatchHeader h = a.patches[pid];
  const int n_slots = int(a.patches[pid + 1].slot_off - h.slot_off);
  const int P = h.n_owned;
  const bool do_volume = a.modules & MS_MOD_VOLUME;
  const bool do_tilt = (a.modules & MS_MOD_TILT) && a.tilts != nullptr;
  SmemB s(smem_raw, a, BENDING, do_tilt);

  stage_topology(a, h, n_slots, s.recs, s.halo);
  {
    const double* owned = a.pos + size_t(h.v_lo) * 3;
    for (int j = threadIdx.x; j < 3 * P; j += blockDim.x) cp_async8(s.pos + j, owned + j);
    if (BENDING) {
      const double* so = a.seeds + size_t(h.v_lo) * kSeedStride;
      for (int j = threadIdx.x; j < kSeedStride * P; j += blockDim.x) cp_async8(s.seed + j, so + j);
    }
  }
  for (int j = threadIdx.x; j < 6 * a.max_owned; j += blockDim.x) s.accG[j] = 0.0;
  if (do_tilt)
    for (int j = threadIdx.x; j < P; j += blockDim.x) s.accAb[j] = 0.0;
  cp_async_wait_all();
  __syncthreads();
  for (int j = threadIdx.x; j < 3 * h.n_halo; j += blockDim.x) {
    const int v = j / 3, c = j - v * 3;
    cp_async8(s.pos + 3 * P + j, a.pos + size_t(s.halo[v]) * 3 + c);
  }
  if (BENDING) {
    double* d2 = s.seed + kSeedStride * P;
    for (int j = threadIdx.x; j < kSeedStride * h.n_halo; j += blockDim.x) {
      const int v = j / kSeedStride, c = j - v * kSeedStride;
      cp_async8(d2 + j, a.seeds + size_t(s.halo[v]) * kSeedStride + c);
    }
    if (a.is_boundary) stage_flags(s.bfl, a.is_boundary, h, s.halo);
  }
  if (do_tilt) stage_tilt_sq(s.t2, a.tilts, h, s.halo);
  cp_async_wait_all();
  __syncthreads();

  LocalB loc;
  loc.pos = s.pos; loc.seed = s.seed; loc.bfl = a.is_boundary ? s.bfl : nullptr; loc.t2 = do_tilt ? s.t2 : nullptr;
  loc.accG = s.accG; loc.accV = s.accV; loc.accAb = s.accAb; loc.P = P;
  double sums[PS_COUNT];
#pragma unroll
  for (int k = 0; k < PS_COUNT; ++k) sums[k] = 0.0;

  const double* slot_gamma = a.slot_gamma ? a.slot_gamma + h.slot_off : nullptr;
  const int grp = threadIdx.x / a.threads, lane = threadIdx.x - grp * a.threads;
  for (int r0 = 0; r0 < h.n_rounds; r0 += a.groups) {
    const int r = r0 + grp;
    bool act = false;
    FacetRec rec;
    FacetOutB out;
    if (r < h.n_rounds) {
      const int slot = r * a.threads + lane;
      rec = s.recs[slot];
      if (rec.flags & REC_VALID) {
        act = true;
        const double gam = slot_gamma ? slot_gamma[slot] : a.gamma_u;
        out = facet_compute_b<BENDING>(rec, gam, loc, a.modules, a.flags, a.k_tilt, scalars_here, sums);
      }
    }
    for (int g = 0; g < a.groups; ++g) {
      __syncthreads();
      if (act && g == grp) facet_accumulate_b(rec, out, loc);
    }
  }
  __syncthreads();

  // owned-vertex results leave as flat, coalesced copies; the dot products of the KKT
  // projection (constraint_manager.py:294-301) are summed on the way out
  double* gout = a.grad + size_t(h.v_lo) * 3;
  const bool vol_out = do_volume && a.volgrad;
  double* vout = vol_out ? a.volgrad + size_t(h.v_lo) * 3 : nullptr;
  for (int j = threadIdx.x; j < 3 * P; j += blockDim.x) {
    const double gj = s.accG[j];
    gout[j] = gj;
    sums[PS_G_G] += gj * gj;
    if (vol_out) {
      const double vj = s.accV[j];
      vout[j] = vj;
      sums[PS_G_GC] += gj * vj;
      sums[PS_GC_GC] += vj * vj;
    }
  }
  if (do_tilt && a.tilt_grad) {
    // tilt.py:163-170: dE/dt_v = k_t t_v A_bary(v)
    const size_t base = size_t(h.v_lo) * 3;
    for (int j = threadIdx.x; j < 3 * P; j += blockDim.x)
      a.tilt_grad[base + j] = a.k_tilt * a.tilts[base + j] * s.accAb[j / 3];
  }
  block_sum<PS_COUNT>(sums, s.red);
  if (threadIdx.x == 0) {
    double* p = a.partials + size_t(pid) * kPartialStride;
    if (scalars_here) {
#pragma unroll
      for (int k = 0; k < PS_COUNT; ++k) p[k] = sums[k];
    } else {
      p[PS_E_TILT] = sums[PS_E_TILT];
      p[PS_G_G] = sums[PS_G_G];
      p[PS_G_GC] = sums[PS_G_GC];
      p[PS_GC_GC] = sums[PS_GC_GC];
    }
  }
}

// Fixed-order reduction of the per-patch partial sums (one CTA of 64 x 12 threads:
// thread (row, k) sums slot k of patches row, row+64, ... -- coalesced -- and 12 threads
// then add the 64 row sums in index order).
constexpr int kReduceRows = 64;
__global__ void __launch_bounds__(kReduceRows* kPartialStride)
    k_reduce_partials(const double* __restrict__ partials, int begin, int count, double* scalars) {
  __shared__ double part[kReduceRows][kPartialStride];
  const int k = threadIdx.x % kPartialStride, row = threadIdx.x / kPartialStride;
  double v = 0.0;
  for (int p = row; p < count; p += kReduceRows) v += partials[size_t(begin + p) * kPartialStride + k];
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
  block_sum<3>(v, red);
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
  block_sum<3>(v, red);
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

__device__ __forceinline__ void st3(double* p, size_t i, d3 v) {
  p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z;
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

// single-CTA fixed-order sum (stateless shims only; sizes are modest there)
__global__ void __launch_bounds__(256) k_sum(const double* __restrict__ x, int64_t n, double scale,
                                             double* out) {
  __shared__ double red[32];
  double v[1] = {0.0};
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) v[0] += x[i];
  block_sum<1>(v, red);
  if (threadIdx.x == 0) *out = v[0] * scale;
}

inline int blocks_for(int64_t n, int t) { return int((n + t - 1) / t); }

}  // namespace

static size_t topo_smem_bytes(const PatchLaunch& a) {
  return carve_bytes(kRedDoubles, 8) + carve_bytes(size_t(a.max_slots), sizeof(FacetRec)) +
         carve_bytes(size_t(a.max_local), 4) +
         carve_bytes(size_t(a.max_local), 1);
}

size_t pass_a_smem_bytes(const PatchLaunch& a, bool tilt) {
  return carve_bytes(3 * size_t(a.max_local), 8) + carve_bytes(tilt ? a.max_local : 0, 8) +
         carve_bytes(5 * size_t(a.max_owned), 8) + carve_bytes(3 * size_t(a.max_owned), 8) +
         topo_smem_bytes(a);
}

size_t pass_b_smem_bytes(const PatchLaunch& a, bool bending, bool tilt) {
  return carve_bytes(bending ? size_t(kSeedStride) * a.max_local : 0, 8) +
         carve_bytes(3 * size_t(a.max_local), 8) + carve_bytes(tilt ? a.max_local : 0, 8) +
         carve_bytes(6 * size_t(a.max_owned), 8) + carve_bytes(tilt ? a.max_owned : 0, 8) +
         topo_smem_bytes(a);
}

cudaError_t configure_kernels() {
  const int max_dyn = 227 * 1024;
  cudaError_t e = cudaFuncSetAttribute(k_pass_a, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_pass_b<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_pass_b<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn);
}

cudaError_t launch_pass_a(const PatchLaunch& a, cudaStream_t st) {
  if (a.patch_count <= 0) return cudaSuccess;
  const size_t smem = pass_a_smem_bytes(a, (a.modules & MS_MOD_TILT) && a.tilts);
  k_pass_a<<<a.patch_count, a.threads * a.groups, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_pass_b(const PatchLaunch& a, bool bending, bool scalars_here, cudaStream_t st) {
  if (a.patch_count <= 0) return cudaSuccess;
  const bool tilt = (a.modules & MS_MOD_TILT) && a.tilts;
  const size_t smem = pass_b_smem_bytes(a, bending, tilt);
  if (bending)
    k_pass_b<true><<<a.patch_count, a.threads * a.groups, smem, st>>>(a, scalars_here);
  else
    k_pass_b<false><<<a.patch_count, a.threads * a.groups, smem, st>>>(a, scalars_here);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* partials, int begin, int count, double* scalars,
                                   cudaStream_t st) {
  k_reduce_partials<<<1, kReduceRows * kPartialStride, 0, st>>>(partials, begin, count, scalars);
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

cudaError_t launch_sum(const double* x, int64_t n, double scale, double* out, cudaStream_t st) {
  k_sum<<<1, 256, 0, st>>>(x, n, scale, out);
  return cudaGetLastError();
}

}  // namespace ms
