// Patch packer: see ms_pack.h.
#include "ms_pack.h"

#include <algorithm>
#include <cstring>

namespace ms {

namespace {

inline bool facet_valid(const int32_t* t, int32_t nv) {
  return t[0] >= 0 && t[0] < nv && t[1] >= 0 && t[1] < nv && t[2] >= 0 && t[2] < nv;
}

// vertex -> incident valid facets (a facet appears once per corner it has at the vertex)
void build_vertex_facets(int32_t nv, int32_t nf, const int32_t* tri, std::vector<int64_t>& ptr,
                         std::vector<int32_t>& fac) {
  ptr.assign(size_t(nv) + 1, 0);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_valid(t, nv)) continue;
    for (int k = 0; k < 3; ++k) ++ptr[size_t(t[k]) + 1];
  }
  for (int32_t v = 0; v < nv; ++v) ptr[size_t(v) + 1] += ptr[v];
  fac.resize(size_t(ptr[nv]));
  std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
  for (int32_t f = 0; f < nf; ++f) {
    const int32_t* t = tri + 3 * size_t(f);
    if (!facet_valid(t, nv)) continue;
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

int pack_patches(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* body_mask,
                 const PackParams& prm, PackedMesh& out, int32_t n_owned_vertices) {
  if (n_owned_vertices < 0 || n_owned_vertices > nv) n_owned_vertices = nv;
  if (nv < 0 || nf < 0 || (nf > 0 && !tri) || prm.threads <= 0 || prm.max_owned <= 0 ||
      prm.max_local <= 0 || prm.max_local > 65535)
    return -1;
  out = PackedMesh();
  out.nv = nv;
  out.nf = nf;
  out.params = prm;

  std::vector<int64_t> vptr;
  std::vector<int32_t> vfac;
  build_vertex_facets(nv, nf, tri, vptr, vfac);
  for (int32_t f = 0; f < nf; ++f) out.n_valid += facet_valid(tri + 3 * size_t(f), nv) ? 1 : 0;

  Scratch s;
  s.facet_stamp.assign(size_t(nf), -1);
  s.vert_stamp.assign(size_t(nv), -1);
  s.vert_local.assign(size_t(nv), 0);

  const int32_t T = prm.threads;
  std::vector<uint64_t> used;      // per owned vertex: bitmask words of occupied rounds
  std::vector<int32_t> round_fill; // facets already placed in each round
  std::vector<int32_t> round_of;   // per listed facet
  int32_t stamp = 0;

  out.n_owned_vertices = n_owned_vertices;
  for (int32_t v_lo = 0; v_lo < n_owned_vertices;) {
    int32_t n = std::min(prm.max_owned, n_owned_vertices - v_lo);
    int64_t n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
    while (n_local > prm.max_local && n > 1) {
      n = std::max(1, n / 2);
      n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
    }
    if (n_local > prm.max_local) return -2;

    // --- schedule the facets into conflict-free rounds ---
    // A facet may join a round if none of its OWNED corners is written in that round
    // and the round holds fewer than T facets.  Among the feasible rounds the least
    // filled one is taken, so rounds stay balanced; a new round opens only when needed.
    const size_t nfac = s.facets.size();
    const int32_t v_hi = v_lo + n;
    int32_t n_rounds = std::max<int32_t>(int32_t((nfac + size_t(T) - 1) / size_t(T)), std::min<int32_t>(7, int32_t(nfac)));
    int32_t words = (n_rounds + 63) / 64 + 1;
    used.assign(size_t(n) * size_t(words), 0);
    round_fill.assign(size_t(n_rounds), 0);
    round_of.assign(nfac, 0);
    for (size_t i = 0; i < nfac; ++i) {
      const int32_t* t = tri + 3 * size_t(s.facets[i]);
      int32_t own[3];
      int n_own = 0;
      for (int k = 0; k < 3; ++k)
        if (t[k] >= v_lo && t[k] < v_hi) own[n_own++] = t[k] - v_lo;
      int32_t best = -1;
      for (int32_t r = 0; r < n_rounds; ++r) {
        if (round_fill[r] >= T) continue;
        if (best >= 0 && round_fill[r] >= round_fill[best]) continue;
        bool clash = false;
        for (int k = 0; k < n_own; ++k)
          clash |= (used[size_t(own[k]) * size_t(words) + size_t(r >> 6)] >> (r & 63)) & 1u;
        if (!clash) best = r;
      }
      if (best < 0) {
        best = n_rounds++;
        round_fill.push_back(0);
        if (n_rounds > words * 64) {  // grow the per-vertex masks by one word
          std::vector<uint64_t> grown(size_t(n) * size_t(words + 1), 0);
          for (int32_t v = 0; v < n; ++v)
            for (int32_t w = 0; w < words; ++w)
              grown[size_t(v) * size_t(words + 1) + w] = used[size_t(v) * size_t(words) + w];
          used.swap(grown);
          ++words;
        }
      }
      ++round_fill[best];
      round_of[i] = best;
      for (int k = 0; k < n_own; ++k)
        used[size_t(own[k]) * size_t(words) + size_t(best >> 6)] |= uint64_t(1) << (best & 63);
    }
    // drop rounds that stayed empty (tiny patches)
    {
      std::vector<int32_t> remap(size_t(n_rounds), -1);
      int32_t kept = 0;
      for (int32_t r = 0; r < n_rounds; ++r)
        if (round_fill[r] > 0) remap[r] = kept++;
      for (size_t i = 0; i < nfac; ++i) round_of[i] = remap[round_of[i]];
      std::vector<int32_t> fill2(size_t(kept), 0);
      for (int32_t r = 0; r < n_rounds; ++r)
        if (remap[r] >= 0) fill2[remap[r]] = round_fill[r];
      round_fill.swap(fill2);
      n_rounds = kept;
    }

    // --- emit header, halo list and the records ---
    // Round r owns the T record slots [slot_off + r*T, slot_off + (r+1)*T).  Inside a round the
    // records are placed so that, within every half-warp (16 consecutive lanes), the local
    // vertex indices seen at corner 0, at corner 1 and at corner 2 are pairwise distinct
    // modulo 16 (or identical).  All per-vertex rows in shared memory have an odd stride in
    // doubles (3 or 5), so such a half-warp performs its 64-bit gathers and its
    // read-modify-writes without bank conflicts.  A facet may be rotated cyclically to fit
    // (the per-facet math is invariant under cyclic relabelling); unused slots stay invalid.
    PatchHeader h;
    h.v_lo = v_lo;
    h.n_owned = n;
    h.halo_off = int32_t(out.halo_ids.size());
    h.n_halo = int32_t(s.halo.size());
    h.slot_off = int64_t(out.recs.size());
    h.reserved = 0;
    h.n_rounds = n_rounds;
    out.patches.push_back(h);
    out.halo_ids.insert(out.halo_ids.end(), s.halo.begin(), s.halo.end());

    const size_t base = out.recs.size();
    const size_t n_slots = size_t(n_rounds) * size_t(T);
    FacetRec empty;
    empty.a = empty.b = empty.c = 0;
    empty.flags = 0;
    out.recs.resize(base + n_slots, empty);
    out.slot_facet.resize(base + n_slots, -1);

    const int n_hw = (T + 15) / 16;
    // occupant[(round*n_hw + hw)*48 + corner*16 + residue] = local index or -1
    std::vector<int32_t> occupant(size_t(n_rounds) * size_t(n_hw) * 48, -1);
    std::vector<int32_t> hw_fill(size_t(n_rounds) * size_t(n_hw), 0);
    for (size_t i = 0; i < nfac; ++i) {
      const int32_t f = s.facets[i];
      const int32_t* t = tri + 3 * size_t(f);
      int32_t loc[3];
      for (int k = 0; k < 3; ++k) {
        const int32_t u = t[k];
        // halo slots were assigned by the last collect() call of this patch
        loc[k] = (u >= v_lo && u < v_hi) ? (u - v_lo) : (n + s.vert_local[u]);
      }
      uint16_t flags = REC_VALID;
      if (t[0] >= v_lo && t[0] < v_hi) flags |= REC_PRIMARY;
      if (body_mask && body_mask[f]) flags |= REC_BODY;
      const int32_t r = round_of[i];
      int best_hw = -1, best_rot = 0, best_cost = 1 << 30;
      for (int hw = 0; hw < n_hw && best_cost > 0; ++hw) {
        const int cap = std::min(16, T - hw * 16);
        if (hw_fill[size_t(r) * n_hw + hw] >= cap) continue;
        const int32_t* occ = occupant.data() + (size_t(r) * n_hw + hw) * 48;
        for (int rot = 0; rot < 3; ++rot) {
          int cost = 0;
          for (int k = 0; k < 3; ++k) {
            const int32_t idx = loc[(k + rot) % 3];
            const int32_t o = occ[k * 16 + (idx & 15)];
            cost += (o >= 0 && o != idx) ? 1 : 0;
          }
          if (cost < best_cost) {
            best_cost = cost;
            best_hw = hw;
            best_rot = rot;
            if (cost == 0) break;
          }
        }
      }
      // a round never holds more than T facets, so some half-warp has a free lane
      int32_t* occ = occupant.data() + (size_t(r) * n_hw + best_hw) * 48;
      for (int k = 0; k < 3; ++k) {
        const int32_t idx = loc[(k + best_rot) % 3];
        if (occ[k * 16 + (idx & 15)] < 0) occ[k * 16 + (idx & 15)] = idx;
      }
      const int lane = hw_fill[size_t(r) * n_hw + best_hw]++;
      out.n_lane_conflicts += best_cost;
      FacetRec rec;
      rec.a = uint16_t(loc[best_rot % 3]);
      rec.b = uint16_t(loc[(1 + best_rot) % 3]);
      rec.c = uint16_t(loc[(2 + best_rot) % 3]);
      rec.flags = flags;
      const size_t slot = base + size_t(r) * size_t(T) + size_t(best_hw) * 16 + size_t(lane);
      out.recs[slot] = rec;
      out.slot_facet[slot] = f;
    }
    out.max_owned = std::max(out.max_owned, n);
    out.max_local = std::max(out.max_local, int32_t(n_local));
    out.max_rounds = std::max(out.max_rounds, n_rounds);
    out.max_slots = std::max(out.max_slots, int32_t(n_slots));
    out.n_round_slots += int64_t(n_rounds) * int64_t(T);
    out.n_listed += int64_t(nfac);
    v_lo += n;
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
