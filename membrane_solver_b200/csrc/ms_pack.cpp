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

// ---------------------------------------------------------------------------
// Strips of one patch.
//
// A strip is a vertex sequence x0 x1 x2 ...; its facet j is {x_j, x_j+1, x_j+2}: going from facet j
// to j+1 crosses the edge {x_j+1, x_j+2} and drops x_j (zigzag).  Through a facet run three such
// strips (one per choice of the vertex shared by the entry and the exit edge); the builder takes
// the longest one not yet used, extending it in both directions.
// ---------------------------------------------------------------------------
struct StripWork {
  std::vector<int32_t> floc;     // 3 * nfac patch-local vertex indices, facet orientation
  std::vector<int32_t> adj;      // 3 * nfac: facet across the edge opposite corner k, -1 if none
  std::vector<uint64_t> edges;   // sort keys of the edge matching
  std::vector<uint8_t> visited;
  std::vector<int32_t> stamp;
  int32_t stamp_id = 0;
  std::vector<int32_t> seq_v, seq_f;    // vertex / facet sequences of all strips, concatenated
  std::vector<int32_t> strip_begin;     // strip i: facets seq_f[strip_begin[i] .. strip_begin[i+1]),
                                        // vertices seq_v[strip_begin[i] + 2 i ...] (two more than facets)
  std::vector<int32_t> tmp_v, tmp_f, back_v, back_f;
  // emission
  std::vector<uint32_t> words;
  std::vector<int32_t> word_facet;
  struct Event { int32_t vertex, word; };
  std::vector<Event> events;
  std::vector<int32_t> evt_count;
  std::vector<uint16_t> evt_ptr;
  int32_t n_steps = 0, n_steps_used = 0, n_events = 0, n_pieces = 0;
  int64_t warp_compute = 0, gather_groups = 0, gather_excess = 0;
};

void build_adjacency(StripWork& w, int32_t nfac) {
  w.edges.clear();
  w.edges.reserve(3 * size_t(nfac));
  for (int32_t f = 0; f < nfac; ++f)
    for (int k = 0; k < 3; ++k) {
      const int32_t a = w.floc[3 * size_t(f) + (k + 1) % 3], b = w.floc[3 * size_t(f) + (k + 2) % 3];
      const uint64_t lo = uint64_t(std::min(a, b)), hi = uint64_t(std::max(a, b));
      w.edges.push_back((((lo << 16) | hi) << 32) | uint64_t(3 * f + k));
    }
  std::sort(w.edges.begin(), w.edges.end());
  w.adj.assign(3 * size_t(nfac), -1);
  for (size_t i = 0; i < w.edges.size();) {
    size_t j = i;
    while (j < w.edges.size() && (w.edges[j] >> 32) == (w.edges[i] >> 32)) ++j;
    // manifold edge: exactly two facets.  More (non-manifold): pair them up in index order.
    for (size_t q = i; q + 1 < j; q += 2) {
      const int32_t ca = int32_t(w.edges[q] & 0xffffffffu), cb = int32_t(w.edges[q + 1] & 0xffffffffu);
      if (ca / 3 == cb / 3) continue;
      w.adj[size_t(ca)] = cb / 3;
      w.adj[size_t(cb)] = ca / 3;
    }
    i = j;
  }
}

// Directed zigzag walk leaving facet f across the edge opposite its corner kd, with the strip's last two
// vertices (p, q) at corners (kp, kq).  Visits facets that are neither used nor stamped in this search.
int walk(StripWork& w, int32_t f, int kd, int kp, int kq, std::vector<int32_t>* out_v, std::vector<int32_t>* out_f) {
  int32_t cur = f;
  int cur_kd = kd;
  int32_t p = w.floc[3 * size_t(f) + kp], q = w.floc[3 * size_t(f) + kq];
  int n = 0;
  for (;;) {
    const int32_t g = w.adj[3 * size_t(cur) + cur_kd];
    if (g < 0 || w.visited[size_t(g)] || w.stamp[size_t(g)] == w.stamp_id) break;
    const int32_t* t = &w.floc[3 * size_t(g)];
    int kr = -1, kpp = -1;
    for (int k = 0; k < 3; ++k) {
      if (t[k] != p && t[k] != q) kr = k;
      if (t[k] == p) kpp = k;
    }
    if (kr < 0 || kpp < 0) break;
    w.stamp[size_t(g)] = w.stamp_id;
    const int32_t r = t[kr];
    if (out_v) { out_v->push_back(r); out_f->push_back(g); }
    ++n;
    cur = g;
    cur_kd = kpp;
    p = q;
    q = r;
  }
  return n;
}

void build_strips(StripWork& w, int32_t nfac) {
  build_adjacency(w, nfac);
  w.visited.assign(size_t(nfac), 0);
  w.stamp.assign(size_t(nfac), -1);
  w.seq_v.clear();
  w.seq_f.clear();
  w.strip_begin.clear();
  for (int32_t f = 0; f < nfac; ++f) {
    if (w.visited[size_t(f)]) continue;
    int best_k = 0, best_len = -1;
    for (int kd = 0; kd < 3; ++kd) {
      const int kp = (kd + 1) % 3, kq = (kd + 2) % 3;
      ++w.stamp_id;
      w.stamp[size_t(f)] = w.stamp_id;
      const int len = walk(w, f, kd, kp, kq, nullptr, nullptr) + walk(w, f, kq, kp, kd, nullptr, nullptr);
      if (len > best_len) { best_len = len; best_k = kd; }
    }
    const int kd = best_k, kp = (kd + 1) % 3, kq = (kd + 2) % 3;
    ++w.stamp_id;
    w.stamp[size_t(f)] = w.stamp_id;
    w.tmp_v.clear(); w.tmp_f.clear(); w.back_v.clear(); w.back_f.clear();
    walk(w, f, kd, kp, kq, &w.tmp_v, &w.tmp_f);
    walk(w, f, kq, kp, kd, &w.back_v, &w.back_f);
    w.strip_begin.push_back(int32_t(w.seq_f.size()));
    for (size_t i = w.back_v.size(); i-- > 0;) w.seq_v.push_back(w.back_v[i]);
    w.seq_v.push_back(w.floc[3 * size_t(f) + kd]);
    w.seq_v.push_back(w.floc[3 * size_t(f) + kp]);
    w.seq_v.push_back(w.floc[3 * size_t(f) + kq]);
    for (int32_t v : w.tmp_v) w.seq_v.push_back(v);
    for (size_t i = w.back_f.size(); i-- > 0;) w.seq_f.push_back(w.back_f[i]);
    w.seq_f.push_back(f);
    for (int32_t g : w.tmp_f) w.seq_f.push_back(g);
    w.visited[size_t(f)] = 1;
    for (int32_t g : w.tmp_f) w.visited[size_t(g)] = 1;
    for (int32_t g : w.back_f) w.visited[size_t(g)] = 1;
  }
  w.strip_begin.push_back(int32_t(w.seq_f.size()));
}

// Cut the strips into lane pieces and emit the step words, the event rows and the statistics.  The facets
// fill the lanes one after the other, n_steps each: a piece is the run of consecutive facets of one strip
// inside one lane; it begins with a restart step that also replaces the two other slots.
void emit_steps(StripWork& w, int32_t lanes, int32_t n_owned, const int32_t* patch_facets, const int32_t* tri,
                const uint8_t* body_mask, int32_t v_lo) {
  const int32_t nfac = int32_t(w.seq_f.size());
  const int32_t S_used = (nfac + lanes - 1) / lanes;
  const int32_t S = (S_used + 2) / 3 * 3;  // whole triples of steps (the kernels rotate three word registers); padded rows are no-ops
  w.n_steps = S;
  w.n_steps_used = S_used;
  // restarts per lane -> number of restart rows
  const size_t ns = w.strip_begin.size() - 1;
  int32_t max_restarts = 0;
  {
    int32_t lane = 0, s = 0, r = 0;
    for (size_t i = 0; i < ns; ++i) {
      int32_t rem = w.strip_begin[i + 1] - w.strip_begin[i];
      while (rem > 0) {
        if (s == S_used) { ++lane; s = 0; r = 0; }
        const int32_t m = std::min(rem, S_used - s);
        max_restarts = std::max(max_restarts, ++r);
        s += m;
        rem -= m;
      }
    }
  }
  const size_t aux0 = size_t(S + 3) * size_t(lanes);
  const size_t n_words = aux0 + 2 * size_t(max_restarts) * size_t(lanes);
  w.words.assign(n_words, 0u);
  w.word_facet.assign(n_words, -1);
  w.events.clear();
  w.n_pieces = 0;
  int32_t slot_v[3] = {-1, -1, -1};
  bool dirty[3] = {false, false, false};
  int32_t lane = 0, s = 0, r = 0;  // next free step / restart row of the current lane
  auto replace = [&](int k, int32_t x, size_t wi) {  // flush the slot's vertex, load x
    if (dirty[k]) w.events.push_back({slot_v[k], int32_t(wi)});
    w.words[wi] |= uint32_t(x);
    slot_v[k] = x;
    dirty[k] = false;
  };
  auto close_lane = [&]() {
    for (int k = 0; k < 3; ++k) {
      if (dirty[k]) w.events.push_back({slot_v[k], int32_t(size_t(S + k) * size_t(lanes) + size_t(lane))});
      dirty[k] = false;
      slot_v[k] = -1;
    }
  };
  const int32_t v_hi = v_lo + n_owned;
  for (size_t i = 0; i < ns; ++i) {
    int32_t j = w.strip_begin[i];
    const int32_t j_end = w.strip_begin[i + 1];
    const int32_t voff = 2 * int32_t(i);  // vertex sequence of strip i starts at strip_begin[i] + 2 i
    while (j < j_end) {
      if (s == S_used) {
        close_lane();
        ++lane;
        s = 0;
        r = 0;
      }
      const int32_t m = std::min(j_end - j, S_used - s);
      ++w.n_pieces;
      for (int32_t t = 0; t < m; ++t, ++s) {
        const int k = s % 3;
        const size_t wi = size_t(s) * size_t(lanes) + size_t(lane);
        if (t == 0) {  // restart: the strip's two leading vertices go to the slots that are replaced next
          w.words[wi] |= STEP_RESTART;
          for (int q = 1; q <= 2; ++q)
            replace((k + q) % 3, w.seq_v[size_t(j + voff + q - 1)], aux0 + size_t(2 * r + q - 1) * size_t(lanes) + size_t(lane));
          ++r;
        }
        replace(k, w.seq_v[size_t(j + voff + t + 2)], wi);
        const int32_t fl = w.seq_f[size_t(j + t)];   // patch-local facet number
        const int32_t f = patch_facets[fl];
        const int32_t* tg = tri + 3 * size_t(f);
        const int32_t* tl = &w.floc[3 * size_t(fl)];
        uint32_t word = STEP_LOAD | STEP_COMPUTE;
        if (tg[0] >= v_lo && tg[0] < v_hi) word |= STEP_PRIMARY;
        if (body_mask && body_mask[f]) word |= STEP_BODY;
        int i0 = 0;
        while (i0 < 3 && tl[i0] != slot_v[0]) ++i0;
        if (slot_v[1] != tl[(i0 + 1) % 3]) word |= STEP_NEG;
        for (int kk = 0; kk < 3; ++kk)
          if (slot_v[kk] < n_owned) dirty[kk] = true;
        w.word_facet[wi] = f;
        w.words[wi] |= word;
      }
      j += m;
    }
  }
  if (nfac > 0) close_lane();
  // event rows: grouped by owned vertex, in generation order (lane-major, then step)
  w.evt_count.assign(size_t(n_owned) + 1, 0);
  for (const StripWork::Event& e : w.events) ++w.evt_count[size_t(e.vertex) + 1];
  for (int32_t v = 0; v < n_owned; ++v) w.evt_count[size_t(v) + 1] += w.evt_count[size_t(v)];
  w.n_events = w.evt_count[size_t(n_owned)];
  w.evt_ptr.assign(size_t(n_owned) + 1, 0);
  for (int32_t v = 0; v <= n_owned; ++v) w.evt_ptr[size_t(v)] = uint16_t(std::min<int32_t>(w.evt_count[size_t(v)], 65535));
  if (w.n_events <= kMaxPatchEvents) {
    std::vector<int32_t>& cur = w.evt_count;  // reuse as running cursor
    for (const StripWork::Event& e : w.events) {
      const int32_t row = cur[size_t(e.vertex)]++;
      w.words[size_t(e.word)] |= uint32_t(row + 1) << STEP_EVENT_SHIFT;
    }
  }
  // statistics
  w.warp_compute = 0;
  w.gather_groups = 0;
  w.gather_excess = 0;
  for (int32_t st = 0; st < S; ++st) {
    for (int32_t l0 = 0; l0 < lanes; l0 += 32) {
      bool any = false;
      for (int32_t l = l0; l < l0 + 32; ++l) any |= (w.words[size_t(st) * lanes + l] & STEP_COMPUTE) != 0;
      w.warp_compute += any ? 1 : 0;
    }
    for (int32_t l0 = 0; l0 < lanes; l0 += 16) {
      int32_t seen[16][16];
      int cnt[16] = {0};
      bool any = false;
      for (int32_t l = l0; l < l0 + 16; ++l) {
        const uint32_t word = w.words[size_t(st) * lanes + l];
        if (!(word & STEP_LOAD)) continue;
        any = true;
        const int32_t idx = int32_t(word & STEP_INDEX_MASK);
        const int res = idx & 15;
        bool dup = false;
        for (int q = 0; q < cnt[res]; ++q) dup |= seen[res][q] == idx;
        if (!dup) seen[res][cnt[res]++] = idx;
      }
      if (!any) continue;
      int mx = 1;
      for (int q = 0; q < 16; ++q) mx = std::max(mx, cnt[q]);
      w.gather_groups += 1;
      w.gather_excess += mx - 1;
    }
  }
}

// Build everything for the patch owning [v_lo, v_lo + n).  Returns the local vertex count.
int64_t build_patch(int32_t v_lo, int32_t n, int32_t& stamp, const int32_t* tri, const uint8_t* body_mask,
                    const PackParams& prm, const std::vector<int64_t>& vptr, const std::vector<int32_t>& vfac,
                    Scratch& s, StripWork& w) {
  const int64_t n_local = collect(v_lo, n, stamp++, tri, vptr, vfac, s);
  if (n_local > prm.max_local) return n_local;
  const int32_t v_hi = v_lo + n;
  const size_t nfac = s.facets.size();
  w.floc.resize(3 * nfac);
  for (size_t i = 0; i < nfac; ++i) {
    const int32_t* t = tri + 3 * size_t(s.facets[i]);
    for (int k = 0; k < 3; ++k) {
      const int32_t u = t[k];
      w.floc[3 * i + k] = (u >= v_lo && u < v_hi) ? u - v_lo : n + s.vert_local[u];
    }
  }
  build_strips(w, int32_t(nfac));
  emit_steps(w, prm.threads, n, s.facets.data(), tri, body_mask, v_lo);
  return n_local;
}

inline int32_t even_down(int32_t n) { return n > 2 ? (n & ~1) : n; }

}  // namespace

// Packs the patches owning the vertex rows [v_begin, v_end) into `part` (offsets relative to `part`).
static int pack_range(const int32_t* tri, const uint8_t* body_mask, const PackParams& prm,
                      const std::vector<int64_t>& vptr, const std::vector<int32_t>& vfac, int32_t v_begin,
                      int32_t v_end, PackedMesh& part, Scratch& s, StripWork& w, int32_t& stamp) {
  const int32_t lanes = prm.threads;
  for (int32_t v_lo = v_begin; v_lo < v_end;) {
    int32_t n = std::min(prm.max_owned, v_end - v_lo);
    if (v_lo + n < v_end) n = even_down(n);  // even row ranges: the owned rows move as one bulk copy
    int64_t n_local = build_patch(v_lo, n, stamp, tri, body_mask, prm, vptr, vfac, s, w);
    while ((n_local > prm.max_local || w.n_events > prm.max_events) && n > 1) {
      n = std::max(1, even_down(n / 2));
      n_local = build_patch(v_lo, n, stamp, tri, body_mask, prm, vptr, vfac, s, w);
    }
    if (n_local > prm.max_local) return -2;
    if (w.n_events > prm.max_events) return -3;
    // Trim: the lanes run ceil(facets / lanes) steps; shrink the patch to the largest owned range whose
    // facets fill a whole number of steps (found by bisection on the facet count alone).
    if (prm.trim && v_lo + n < v_end && n >= 64) {
      const int64_t nfac0 = int64_t(s.facets.size());
      const int64_t steps_down = nfac0 / lanes;
      if (steps_down >= 1 && nfac0 > steps_down * lanes) {
        int32_t lo = 2, hi = n;  // invariant: facets(lo) <= target < facets(hi)
        const int64_t target = steps_down * lanes;
        while (hi - lo > 2) {
          const int32_t mid = even_down((lo + hi) / 2);
          if (mid <= lo) break;
          collect(v_lo, mid, stamp++, tri, vptr, vfac, s);
          if (int64_t(s.facets.size()) <= target) lo = mid; else hi = mid;
        }
        // vertices per step: trimmed lo / steps_down against untrimmed n / (steps_down + 1); a smaller patch
        // recomputes more ring facets, so it must win by a margin
        if (double(lo) / double(steps_down) > 1.02 * double(n) / double(steps_down + 1)) n = lo;
        n_local = build_patch(v_lo, n, stamp, tri, body_mask, prm, vptr, vfac, s, w);
      }
    }

    PatchHeader h;
    h.v_lo = v_lo;
    h.n_owned = n;
    h.halo_off = int32_t(part.halo_ids.size());
    h.n_halo = int32_t(s.halo.size());
    h.step_off = int64_t(part.steps.size());
    h.n_steps = w.n_steps;
    h.n_events = w.n_events;
    h.fac_off = int64_t(part.recs.size());
    h.n_fac = int32_t(s.facets.size());
    h.evt_off = int32_t(part.evt_ptr.size());
    part.patches.push_back(h);
    part.halo_ids.insert(part.halo_ids.end(), s.halo.begin(), s.halo.end());
    part.steps.insert(part.steps.end(), w.words.begin(), w.words.end());
    part.step_facet.insert(part.step_facet.end(), w.word_facet.begin(), w.word_facet.end());
    part.evt_ptr.insert(part.evt_ptr.end(), w.evt_ptr.begin(), w.evt_ptr.end());
    while (part.evt_ptr.size() % 8) part.evt_ptr.push_back(0);  // 16-byte granules (bulk copies)
    const int32_t v_hi = v_lo + n;
    for (size_t i = 0; i < s.facets.size(); ++i) {
      const int32_t f = s.facets[i];
      const int32_t* t = tri + 3 * size_t(f);
      FacetRec rec;
      rec.a = uint16_t(w.floc[3 * i]);
      rec.b = uint16_t(w.floc[3 * i + 1]);
      rec.c = uint16_t(w.floc[3 * i + 2]);
      rec.flags = REC_VALID;
      if (t[0] >= v_lo && t[0] < v_hi) rec.flags |= REC_PRIMARY;
      if (body_mask && body_mask[f]) rec.flags |= REC_BODY;
      part.recs.push_back(rec);
    }
    part.max_owned = std::max(part.max_owned, n);
    part.max_local = std::max(part.max_local, int32_t(n_local));
    part.max_steps = std::max(part.max_steps, w.n_steps);
    part.max_events = std::max(part.max_events, w.n_events);
    part.max_words = std::max(part.max_words, int32_t(w.words.size()));
    part.n_listed += int64_t(s.facets.size());
    part.n_strips += int64_t(w.strip_begin.size()) - 1;
    part.n_pieces += w.n_pieces;
    part.n_events += w.n_events;
    part.n_lane_steps += int64_t(lanes) * int64_t(w.n_steps_used);
    part.n_warp_compute += w.warp_compute;
    part.n_gather_groups += w.gather_groups;
    part.n_gather_excess += w.gather_excess;
    v_lo += n;
  }
  return 0;
}

int pack_patches(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* body_mask,
                 const PackParams& prm, PackedMesh& out, int32_t n_owned_vertices) {
  if (n_owned_vertices < 0 || n_owned_vertices > nv) n_owned_vertices = nv;
  if (nv < 0 || nf < 0 || (nf > 0 && !tri) || prm.threads < 32 || prm.threads % 32 || prm.max_owned <= 0 ||
      prm.max_local <= 0 || prm.max_local > int32_t(STEP_INDEX_MASK) || prm.max_events < 3 || prm.max_events > kMaxPatchEvents)
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
    StripWork w;
    s.facet_stamp.assign(size_t(nf), -1);
    s.vert_stamp.assign(size_t(nv), -1);
    s.vert_local.assign(size_t(nv), 0);
    int32_t stamp = 0;
    for (;;) {
      const int64_t c = next.fetch_add(1);
      if (c >= n_chunks) break;
      const int32_t vb = int32_t(c * chunk), ve = int32_t(std::min<int64_t>((c + 1) * chunk, n_owned_vertices));
      rcs[size_t(c)] = pack_range(tri, body_mask, prm, vptr, vfac, vb, ve, parts[size_t(c)], s, w, stamp);
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
    const int64_t step_base = int64_t(out.steps.size());
    const int64_t fac_base = int64_t(out.recs.size());
    const int32_t evt_base = int32_t(out.evt_ptr.size());
    for (PatchHeader h : part.patches) {
      h.halo_off += halo_base;
      h.step_off += step_base;
      h.fac_off += fac_base;
      h.evt_off += evt_base;
      out.patches.push_back(h);
    }
    out.halo_ids.insert(out.halo_ids.end(), part.halo_ids.begin(), part.halo_ids.end());
    out.steps.insert(out.steps.end(), part.steps.begin(), part.steps.end());
    out.step_facet.insert(out.step_facet.end(), part.step_facet.begin(), part.step_facet.end());
    out.evt_ptr.insert(out.evt_ptr.end(), part.evt_ptr.begin(), part.evt_ptr.end());
    out.recs.insert(out.recs.end(), part.recs.begin(), part.recs.end());
    out.max_owned = std::max(out.max_owned, part.max_owned);
    out.max_local = std::max(out.max_local, part.max_local);
    out.max_steps = std::max(out.max_steps, part.max_steps);
    out.max_events = std::max(out.max_events, part.max_events);
    out.max_words = std::max(out.max_words, part.max_words);
    out.n_listed += part.n_listed;
    out.n_strips += part.n_strips;
    out.n_pieces += part.n_pieces;
    out.n_events += part.n_events;
    out.n_lane_steps += part.n_lane_steps;
    out.n_warp_compute += part.n_warp_compute;
    out.n_gather_groups += part.n_gather_groups;
    out.n_gather_excess += part.n_gather_excess;
    part = PackedMesh();  // release
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
