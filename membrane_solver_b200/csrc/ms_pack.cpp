// Patch packer: see ms_pack.h.
#include "ms_pack.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace ms {

namespace {

inline bool facet_valid(const int32_t* t, int32_t nv) {
  return t[0] >= 0 && t[0] < nv && t[1] >= 0 && t[1] < nv && t[2] >= 0 && t[2] < nv;
}

// listed by the patches: valid and naming three different vertices (see ms_pack.h)
inline bool facet_listed(const int32_t* t, int32_t nv) {
  return facet_valid(t, nv) && t[0] != t[1] && t[1] != t[2] && t[0] != t[2];
}

// vertex -> incident valid facets (a facet appears once per corner it has at the vertex)
void build_vertex_facets(int32_t nv, int32_t nf, const int32_t* tri, std::vector<int64_t>& ptr,
                         std::vector<int32_t>& fac) {
  ptr.assign(size_t(nv) + 1, 0);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_listed(t, nv)) continue;
    for (int k = 0; k < 3; ++k) ++ptr[size_t(t[k]) + 1];
  }
  for (int32_t v = 0; v < nv; ++v) ptr[size_t(v) + 1] += ptr[v];
  fac.resize(size_t(ptr[nv]));
  std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_listed(t, nv)) continue;
    for (int k = 0; k < 3; ++k) fac[size_t(cur[t[k]]++)] = f;
  }
}

struct Scratch {
  std::vector<int32_t> facet_stamp;   // per facet: last patch that listed it
  std::vector<int32_t> vert_stamp;    // per vertex: last patch that saw it as halo
  std::vector<int32_t> vert_local;    // per vertex: local halo slot in that patch
  std::vector<int32_t> facets;        // facets of the patch being built
  std::vector<int32_t> halo;          // halo vertex rows of the patch being built
};

// Collect facets and halo of the patch owning [v_lo, v_lo+n). Returns local vertex count.
int64_t collect(int32_t v_lo, int32_t n, int32_t stamp, const int32_t* tri,
                const std::vector<int64_t>& vptr, const std::vector<int32_t>& vfac, Scratch& s) {
  s.facets.clear();
  s.halo.clear();
  const int32_t v_hi = v_lo + n;
  for (int32_t v = v_lo; v < v_hi; ++v) {
    for (int64_t j = vptr[v]; j < vptr[size_t(v) + 1]; ++j) {
      const int32_t f = vfac[size_t(j)];
      if (s.facet_stamp[f] == stamp) continue;
      s.facet_stamp[f] = stamp;
      s.facets.push_back(f);
      const int32_t* t = tri + 3 * size_t(f);
      for (int k = 0; k < 3; ++k) {
        const int32_t u = t[k];
        if (u >= v_lo && u < v_hi) continue;
        if (s.vert_stamp[u] != stamp) {
          s.vert_stamp[u] = stamp;
          s.vert_local[u] = int32_t(s.halo.size());
          s.halo.push_back(u);
        }
      }
    }
  }
  return int64_t(n) + int64_t(s.halo.size());
}

}  // namespace

// Packs the patches owning the vertex rows [v_begin, v_end) into `part` (offsets relative to `part`).
static int pack_range(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* body_mask, const PackParams& prm,
                      const std::vector<int64_t>& vptr, const std::vector<int32_t>& vfac, int32_t v_begin,
                      int32_t v_end, PackedMesh& part, Scratch& s, int32_t& stamp) {
  (void)nv;
  (void)nf;

  const int32_t T = prm.threads;
  std::vector<uint64_t> used;      // per owned vertex: bitmask words of occupied rounds
  std::vector<int32_t> round_fill; // facets already placed in each round
  std::vector<int32_t> round_of;   // per listed facet
  bool too_many_slots = false;
  int32_t forced_n = 0;  // retry size after a patch exceeded the slot capacity
  for (int32_t v_lo = v_begin; v_lo < v_end;) {
    int32_t n = std::min(forced_n > 0 ? forced_n : prm.max_owned, v_end - v_lo);
    int64_t n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
    while (n_local > prm.max_local && n > 1) {
      n = std::max(1, n / 2);
      n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
    }
    if (n_local > prm.max_local) return -2;
    // Trim the patch so that its facets fill a whole number of rounds to the target: a patch
    // with 9.2 rounds' worth of facets would otherwise run 10 rounds at 92 % of the target fill.
    if (forced_n == 0 && n == prm.max_owned && n > 64) {
      const double per_round = double(T) * double(prm.fill_pct) / 100.0;
      const int64_t whole = int64_t(double(s.facets.size()) / per_round);
      int64_t target = int64_t(double(whole) * per_round);
      for (int it = 0; it < 4 && whole >= 4 && int64_t(s.facets.size()) > target; ++it) {
        n = std::max<int32_t>(64, int32_t(double(n) * double(target) / double(s.facets.size())) - (it > 0 ? 2 : 0));
        n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
      }
    }

    // --- schedule the facets into conflict-free rounds AND bank-conflict-free lanes ---
    // A facet may join a round if none of its OWNED corners is written in that round and the
    // round has a free slot.  Round r owns the T record slots [slot_off + r*T, slot_off + (r+1)*T),
    // i.e. T/16 half-warps.  The patch-local arrays in shared memory hold 8-byte elements: the
    // accumulators as structure-of-arrays (element i of a component), the staged positions / seeds
    // as rows of 3 / 5 doubles (element 3 i + k, 5 i + k; 3 and 5 are coprime to 16).  Either way a
    // 64-bit gather / read-modify-write by a half-warp is conflict free exactly when the local
    // vertex indices it uses at one corner position are pairwise distinct modulo 16 (or identical
    // -> broadcast).  ONE clash costs the whole half-warp an extra
    // wavefront, so the placement looks for a (round, half-warp, rotation) with NO clash at any of
    // the three corner positions -- the per-facet math is invariant under cyclic relabelling --
    // preferring the least-filled round; only when none exists does it take the cheapest one.
    const size_t nfac = s.facets.size();
    const int32_t v_hi = v_lo + n;
    const int n_hw = (T + 15) / 16;
    int32_t n_rounds = std::max<int32_t>(int32_t((nfac * 100 + size_t(T) * size_t(prm.fill_pct) - 1) / (size_t(T) * size_t(prm.fill_pct))),
                                         std::min<int32_t>(7, int32_t(nfac)));
    int32_t words = (n_rounds + 63) / 64 + 1;
    used.assign(size_t(n) * size_t(words), 0);
    round_fill.assign(size_t(n_rounds), 0);
    // cnt[((r*n_hw + hw)*3 + k)*16 + residue] = facets of that half-warp whose corner k has that residue
    std::vector<uint8_t> cnt(size_t(n_rounds) * size_t(n_hw) * 48, 0);
    std::vector<int32_t> members(size_t(n_rounds) * size_t(n_hw) * 16, -1);  // facet slots of a half-warp
    std::vector<int32_t> hw_fill(size_t(n_rounds) * size_t(n_hw), 0);
    struct Place { int32_t r, hw, rot, lane; };
    std::vector<Place> place(nfac);
    std::vector<int32_t> floc(3 * nfac);     // local vertex indices of the facets
    std::vector<int32_t> fown(nfac);         // number of owned corners
    for (size_t i = 0; i < nfac; ++i) {
      const int32_t* t = tri + 3 * size_t(s.facets[i]);
      int n_own = 0;
      for (int k = 0; k < 3; ++k) {
        const int32_t u = t[k];
        const bool own = u >= v_lo && u < v_hi;
        n_own += own ? 1 : 0;
        floc[3 * i + k] = own ? u - v_lo : n + s.vert_local[u];  // halo slots come from the last collect()
      }
      fown[i] = n_own;
    }
    auto grow_round = [&]() {
      ++n_rounds;
      round_fill.push_back(0);
      cnt.resize(size_t(n_rounds) * size_t(n_hw) * 48, 0);
      members.resize(size_t(n_rounds) * size_t(n_hw) * 16, -1);
      hw_fill.resize(size_t(n_rounds) * size_t(n_hw), 0);
      if (n_rounds > words * 64) {  // grow the per-vertex masks by one word
        std::vector<uint64_t> grown(size_t(n) * size_t(words + 1), 0);
        for (int32_t v = 0; v < n; ++v)
          for (int32_t w = 0; w < words; ++w)
            grown[size_t(v) * size_t(words + 1) + w] = used[size_t(v) * size_t(words) + w];
        used.swap(grown);
        ++words;
      }
    };
    auto round_free = [&](size_t i, int32_t r) {  // no owned corner of facet i is written in round r
      for (int k = 0; k < 3; ++k) {
        const int32_t l = floc[3 * i + k];
        if (l < n && ((used[size_t(l) * size_t(words) + size_t(r >> 6)] >> (r & 63)) & 1u)) return false;
      }
      return true;
    };
    auto mark_round = [&](size_t i, int32_t r, bool on) {
      for (int k = 0; k < 3; ++k) {
        const int32_t l = floc[3 * i + k];
        if (l >= n) continue;
        uint64_t& w = used[size_t(l) * size_t(words) + size_t(r >> 6)];
        if (on) w |= uint64_t(1) << (r & 63); else w &= ~(uint64_t(1) << (r & 63));
      }
    };
    auto cost_at = [&](size_t i, int32_t r, int hw, int rot) {  // clashing corner positions
      const uint8_t* c = cnt.data() + (size_t(r) * n_hw + hw) * 48;
      int cost = 0;
      for (int k = 0; k < 3; ++k) cost += c[k * 16 + (floc[3 * i + (k + rot) % 3] & 15)] ? 1 : 0;
      return cost;
    };
    auto insert = [&](size_t i, int32_t r, int hw, int rot) {
      uint8_t* c = cnt.data() + (size_t(r) * n_hw + hw) * 48;
      for (int k = 0; k < 3; ++k) ++c[k * 16 + (floc[3 * i + (k + rot) % 3] & 15)];
      int32_t* m = members.data() + (size_t(r) * n_hw + hw) * 16;
      for (int l = 0; l < 16; ++l)
        if (m[l] < 0) { m[l] = int32_t(i); break; }
      ++hw_fill[size_t(r) * n_hw + hw];
      ++round_fill[r];
      place[i].r = r; place[i].hw = hw; place[i].rot = rot;
      mark_round(i, r, true);
    };
    auto remove = [&](size_t i) {
      const Place pl = place[i];
      uint8_t* c = cnt.data() + (size_t(pl.r) * n_hw + pl.hw) * 48;
      for (int k = 0; k < 3; ++k) --c[k * 16 + (floc[3 * i + (k + pl.rot) % 3] & 15)];
      int32_t* m = members.data() + (size_t(pl.r) * n_hw + pl.hw) * 16;
      for (int l = 0; l < 16; ++l)
        if (m[l] == int32_t(i)) { m[l] = -1; break; }
      --hw_fill[size_t(pl.r) * n_hw + pl.hw];
      --round_fill[pl.r];
      mark_round(i, pl.r, false);
    };
    auto is_bad = [&](size_t i) {  // shares a residue with another facet of its half-warp
      const Place pl = place[i];
      const uint8_t* c = cnt.data() + (size_t(pl.r) * n_hw + pl.hw) * 48;
      for (int k = 0; k < 3; ++k)
        if (c[k * 16 + (floc[3 * i + (k + pl.rot) % 3] & 15)] > 1) return true;
      return false;
    };
    // best (round, half-warp, rotation) with a free lane; returns the cost or -1 if no round is feasible
    auto find_slot = [&](size_t i, int32_t& br, int& bhw, int& brot, int stop_cost) {
      int best_cost = 1 << 30, best_fill = 1 << 30;
      br = -1;
      for (int32_t r = 0; r < n_rounds; ++r) {
        if (round_fill[r] >= T) continue;
        if (best_cost <= stop_cost && round_fill[r] >= best_fill) continue;
        if (!round_free(i, r)) continue;
        for (int hw = 0; hw < n_hw; ++hw) {
          if (hw_fill[size_t(r) * n_hw + hw] >= std::min(16, T - hw * 16)) continue;
          bool done = false;
          for (int rot = 0; rot < 3; ++rot) {
            const int cost = cost_at(i, r, hw, rot);
            if (cost < best_cost || (cost == best_cost && round_fill[r] < best_fill)) {
              best_cost = cost; best_fill = round_fill[r];
              br = r; bhw = hw; brot = rot;
            }
            if (cost == 0) { done = true; break; }
          }
          if (done) break;  // this round cannot do better
        }
      }
      return br < 0 ? -1 : best_cost;
    };
    for (size_t i = 0; i < nfac; ++i) {
      int32_t r; int hw, rot;
      if (find_slot(i, r, hw, rot, 0) < 0) {
        grow_round();
        find_slot(i, r, hw, rot, 0);
      }
      insert(i, r, hw, rot);
    }
    // repair: relocate facets that still clash -- to a clash-free free lane anywhere, or by
    // exchanging places with a facet of another half-warp of the same round
    for (int sweep = 0; sweep < prm.repair_sweeps; ++sweep) {
      int64_t fixed = 0;
      for (size_t i = 0; i < nfac; ++i) {
        if (!is_bad(i)) continue;
        const Place old = place[i];
        remove(i);
        int32_t r; int hw, rot;
        if (find_slot(i, r, hw, rot, 0) == 0) {
          insert(i, r, hw, rot);
          ++fixed;
          continue;
        }
        bool swapped = false;
        for (int hw2 = 0; hw2 < n_hw && !swapped; ++hw2) {
          if (hw2 == old.hw) continue;
          for (int l = 0; l < 16 && !swapped; ++l) {
            const int32_t j = members[(size_t(old.r) * n_hw + hw2) * 16 + l];
            if (j < 0) continue;
            const Place pj = place[size_t(j)];
            remove(size_t(j));
            int rot_i = -1, rot_j = -1;
            for (int q = 0; q < 3 && rot_i < 0; ++q)
              if (cost_at(i, old.r, hw2, q) == 0) rot_i = q;
            for (int q = 0; q < 3 && rot_i >= 0 && rot_j < 0; ++q)
              if (cost_at(size_t(j), old.r, old.hw, q) == 0) rot_j = q;
            if (rot_i >= 0 && rot_j >= 0) {
              insert(i, old.r, hw2, rot_i);
              insert(size_t(j), old.r, old.hw, rot_j);
              swapped = true;
              ++fixed;
            } else {
              insert(size_t(j), pj.r, pj.hw, pj.rot);
            }
          }
        }
        if (!swapped) insert(i, old.r, old.hw, old.rot);
      }
      if (fixed == 0) break;
    }
    // lanes: position inside the half-warp's member list
    for (int32_t r = 0; r < n_rounds; ++r)
      for (int hw = 0; hw < n_hw; ++hw) {
        int lane = 0;
        const int32_t* m = members.data() + (size_t(r) * n_hw + hw) * 16;
        for (int l = 0; l < 16; ++l)
          if (m[l] >= 0) place[size_t(m[l])].lane = lane++;
      }
    int64_t patch_clashes = 0;
    for (size_t i = 0; i < nfac; ++i) patch_clashes += is_bad(i) ? 1 : 0;
    // drop rounds that stayed empty (tiny patches)
    std::vector<int32_t> remap(size_t(n_rounds), -1);
    {
      int32_t kept = 0;
      for (int32_t r = 0; r < n_rounds; ++r)
        if (round_fill[r] > 0) remap[r] = kept++;
      n_rounds = kept;
    }
    if (size_t(n_rounds) * size_t(T) > size_t(prm.max_slots) && n > 1) {  // over the slot capacity: halve
      too_many_slots = true;
    } else {
      if (size_t(n_rounds) * size_t(T) > size_t(prm.max_slots)) return -3;
      // --- emit header, halo list and the records ---
      PatchHeader h;
      h.v_lo = v_lo;
      h.n_owned = n;
      h.halo_off = int32_t(part.halo_ids.size());
      h.n_halo = int32_t(s.halo.size());
      h.slot_off = int64_t(part.recs.size());
      h.reserved = 0;
      h.n_rounds = n_rounds;
      part.patches.push_back(h);
      part.halo_ids.insert(part.halo_ids.end(), s.halo.begin(), s.halo.end());

      const size_t base = part.recs.size();
      const size_t n_slots = size_t(n_rounds) * size_t(T);
      FacetRec empty;
      empty.a = empty.b = empty.c = 0;
      empty.flags = 0;
      part.recs.resize(base + n_slots, empty);
      part.slot_facet.resize(base + n_slots, -1);
      for (size_t i = 0; i < nfac; ++i) {
        const int32_t f = s.facets[i];
        const int32_t* t = tri + 3 * size_t(f);
        int32_t loc[3];
        for (int k = 0; k < 3; ++k) {
          const int32_t u = t[k];
          loc[k] = (u >= v_lo && u < v_hi) ? (u - v_lo) : (n + s.vert_local[u]);
        }
        uint16_t flags = REC_VALID;
        if (t[0] >= v_lo && t[0] < v_hi) flags |= REC_PRIMARY;
        if (body_mask && body_mask[f]) flags |= REC_BODY;
        const Place& pl = place[i];
        FacetRec rec;
        rec.a = uint16_t(loc[pl.rot % 3]);
        rec.b = uint16_t(loc[(1 + pl.rot) % 3]);
        rec.c = uint16_t(loc[(2 + pl.rot) % 3]);
        rec.flags = flags;
        const size_t slot = base + size_t(remap[pl.r]) * size_t(T) + size_t(pl.hw) * 16 + size_t(pl.lane);
        part.recs[slot] = rec;
        part.slot_facet[slot] = f;
      }
    }
    if (too_many_slots) {
      too_many_slots = false;
      forced_n = std::max(1, n / 2);
      continue;
    }
    forced_n = 0;
    part.n_lane_conflicts += patch_clashes;
    const size_t n_slots = size_t(n_rounds) * size_t(T);
    part.max_owned = std::max(part.max_owned, n);
    part.max_local = std::max(part.max_local, int32_t(n_local));
    part.max_rounds = std::max(part.max_rounds, n_rounds);
    part.max_slots = std::max(part.max_slots, int32_t(n_slots));
    part.n_round_slots += int64_t(n_rounds) * int64_t(T);
    part.n_listed += int64_t(nfac);
    v_lo += n;
  }
  return 0;
}

int pack_patches(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* body_mask,
                 const PackParams& prm, PackedMesh& out, int32_t n_owned_vertices) {
  if (n_owned_vertices < 0 || n_owned_vertices > nv) n_owned_vertices = nv;
  if (nv < 0 || nf < 0 || (nf > 0 && !tri) || prm.threads <= 0 || prm.max_owned <= 0 ||
      prm.max_local <= 0 || prm.max_local > 65535 || prm.max_slots < prm.threads || prm.fill_pct < 10 || prm.fill_pct > 100)
    return -1;
  out = PackedMesh();
  out.nv = nv;
  out.nf = nf;
  out.params = prm;

  std::vector<int64_t> vptr;
  std::vector<int32_t> vfac;
  build_vertex_facets(nv, nf, tri, vptr, vfac);
  for (int32_t f = 0; f < nf; ++f) out.n_valid += facet_valid(tri + 3 * size_t(f), nv) ? 1 : 0;

  out.n_owned_vertices = n_owned_vertices;
  // Chunks of 32 * max_owned vertex rows are packed independently (the chunk size, not the thread count,
  // decides the patch boundaries, so the layout is the same on every host) and concatenated in order.
  const int64_t chunk = int64_t(32) * int64_t(prm.max_owned);
  const int64_t n_chunks = (int64_t(n_owned_vertices) + chunk - 1) / chunk;
  std::vector<PackedMesh> parts(size_t(std::max<int64_t>(n_chunks, 0)));
  std::vector<int> rcs(parts.size(), 0);
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  if (const char* env = std::getenv("MS_PACK_THREADS")) {  // tuning / debugging knob
    const int v = std::atoi(env);
    if (v > 0) hw = unsigned(v);
  }
  // every worker owns stamp arrays of nf + 2 nv int32: keep their sum below ~6 GB
  const double per_worker = 4.0 * (double(nf) + 2.0 * double(nv)) + 1.0;
  const unsigned cap = unsigned(std::max(1.0, 6.0e9 / per_worker));
  const unsigned n_workers = unsigned(std::max<int64_t>(1, std::min<int64_t>(std::min(hw, cap), n_chunks)));
  std::atomic<int64_t> next(0);
  auto work = [&]() {
    Scratch s;  // stamp arrays live for the whole worker: the stamp counter keeps growing across chunks
    s.facet_stamp.assign(size_t(nf), -1);
    s.vert_stamp.assign(size_t(nv), -1);
    s.vert_local.assign(size_t(nv), 0);
    int32_t stamp = 0;
    for (;;) {
      const int64_t c = next.fetch_add(1);
      if (c >= n_chunks) break;
      const int32_t vb = int32_t(c * chunk), ve = int32_t(std::min<int64_t>((c + 1) * chunk, n_owned_vertices));
      rcs[size_t(c)] = pack_range(nv, nf, tri, body_mask, prm, vptr, vfac, vb, ve, parts[size_t(c)], s, stamp);
    }
  };
  if (n_workers <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_workers; ++t) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  for (int rc : rcs)
    if (rc) return rc;
  for (PackedMesh& part : parts) {
    const int32_t halo_base = int32_t(out.halo_ids.size());
    const int64_t slot_base = int64_t(out.recs.size());
    for (PatchHeader h : part.patches) {
      h.halo_off += halo_base;
      h.slot_off += slot_base;
      out.patches.push_back(h);
    }
    out.halo_ids.insert(out.halo_ids.end(), part.halo_ids.begin(), part.halo_ids.end());
    out.recs.insert(out.recs.end(), part.recs.begin(), part.recs.end());
    out.slot_facet.insert(out.slot_facet.end(), part.slot_facet.begin(), part.slot_facet.end());
    out.max_owned = std::max(out.max_owned, part.max_owned);
    out.max_local = std::max(out.max_local, part.max_local);
    out.max_rounds = std::max(out.max_rounds, part.max_rounds);
    out.max_slots = std::max(out.max_slots, part.max_slots);
    out.n_round_slots += part.n_round_slots;
    out.n_listed += part.n_listed;
    out.n_lane_conflicts += part.n_lane_conflicts;
    part = PackedMesh();  // release
  }
  const int32_t T = prm.threads;
  // exact statistic: extra shared-memory wavefronts per (round, half-warp, corner position) =
  // (largest number of DISTINCT local indices sharing one residue mod 16) - 1
  out.n_hw_groups = 0;
  out.n_hw_excess = 0;
  for (const PatchHeader& h : out.patches) {
    const int n_hwp = (T + 15) / 16;
    for (int32_t r = 0; r < h.n_rounds; ++r)
      for (int hw = 0; hw < n_hwp; ++hw) {
        const FacetRec* rr = out.recs.data() + size_t(h.slot_off) + size_t(r) * size_t(T) + size_t(hw) * 16;
        const int cap = std::min(16, T - hw * 16);
        bool any = false;
        for (int l = 0; l < cap; ++l) any |= (rr[l].flags & REC_VALID) != 0;
        if (!any) continue;
        for (int k = 0; k < 3; ++k) {
          int32_t seen[16][16];
          int cnt[16] = {0};
          for (int l = 0; l < cap; ++l) {
            if (!(rr[l].flags & REC_VALID)) continue;
            const int32_t idx = k == 0 ? rr[l].a : (k == 1 ? rr[l].b : rr[l].c);
            const int res = idx & 15;
            bool dup = false;
            for (int j = 0; j < cnt[res]; ++j) dup |= seen[res][j] == idx;
            if (!dup) seen[res][cnt[res]++] = idx;
          }
          int mx = 1;
          for (int q = 0; q < 16; ++q) mx = std::max(mx, cnt[q]);
          out.n_hw_groups += 1;
          out.n_hw_excess += mx - 1;
        }
      }
  }
  return 0;
}

void build_corner_csr(int32_t nv, int32_t nf, const int32_t* tri, std::vector<int32_t>& ptr,
                      std::vector<int32_t>& idx) {
  ptr.assign(size_t(nv) + 1, 0);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_valid(t, nv)) continue;
    for (int k = 0; k < 3; ++k) ++ptr[size_t(t[k]) + 1];
  }
  for (int32_t v = 0; v < nv; ++v) ptr[size_t(v) + 1] += ptr[v];
  idx.resize(size_t(ptr[nv]));
  std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_valid(t, nv)) continue;
    for (int k = 0; k < 3; ++k) idx[size_t(cur[t[k]]++)] = 3 * f + k;
  }
}

}  // namespace ms
