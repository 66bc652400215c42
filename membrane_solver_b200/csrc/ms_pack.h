// Host-side packing of a triangle mesh into shared-memory sized vertex patches whose
// facets are walked as lane-local TRIANGLE STRIPS.
//
// This replaces, for the device path, what geometry/triangle_rows.py:10-35 and
// Mesh.triangle_row_cache (geometry/mesh.py:597-624) do for the reference: it
// turns the (nf,3) int32 triangle rows into the layout the kernels stream.
//
// A patch owns a contiguous range of vertex rows and lists every facet touching
// an owned vertex (ring facets are listed by each patch that touches them).  The
// facets of a patch are chained into zigzag strips (consecutive facets share an
// edge) and the strips are cut into equally long pieces, one per lane of the team
// of warps that processes the patch.  A lane keeps the three vertices of its
// current facet -- inputs AND partial sums -- in registers: every step brings in
// ONE new vertex (replacing the oldest of the three) instead of three, and the
// partial sums of a vertex leave the registers once, when the vertex leaves the
// strip, as a plain store into an EVENT row that no other lane ever writes.  The
// owned vertices then add up their event rows in index order.  No atomics, no
// read-modify-write on shared memory, no ordering between lanes: the order of
// every floating-point sum is fixed at pack time -> run-to-run reproducible results.
#pragma once

#include <stdint.h>

#include <vector>

namespace ms {

struct PatchHeader {  // 48 bytes
  int32_t v_lo;       // first owned vertex row
  int32_t n_owned;    // owned vertices [v_lo, v_lo + n_owned)
  int32_t halo_off;   // first entry of this patch in halo_ids
  int32_t n_halo;     // local index n_owned + j  ->  vertex row halo_ids[halo_off + j]
  int64_t step_off;   // first step word of this patch: word (s, lane) at step_off + s*lanes + lane,
                      // followed by the tail rows and the restart rows
  int32_t n_steps;    // steps per lane: a multiple of 3 (padded with no-op rows)
  int32_t n_events;   // event rows of this patch
  int64_t fac_off;    // first compact facet record (flat-vertex normal fallback only)
  int32_t n_fac;      // facets listed by this patch
  int32_t evt_off;    // first entry of this patch in evt_ptr (n_owned + 1 entries)
};
static_assert(sizeof(PatchHeader) == 48, "PatchHeader layout");

// One step of one lane.  Step s works on register slot s % 3.
//   bits  0..10  local index of the vertex to load into the slot
//   bit   11     LOAD: gather that vertex, zero the slot's partial sums
//   bit   12     COMPUTE: the three slots now hold a facet -> evaluate it
//   bit   13     PRIMARY: this patch counts the facet's scalars (area, energies, volume)
//   bit   14     BODY: the facet belongs to the body (volume terms)
//   bit   15     NEG: slot order (0,1,2) is an odd permutation of the facet's orientation
//   bit   16     RESTART: a new strip piece begins: the OTHER two slots, (s+1) % 3 and (s+2) % 3, are
//                replaced first, from the lane's next two restart words (index + event fields only)
//   bits 17..31  event row + 1 that receives the partial sums of the slot's PREVIOUS vertex (0: none)
// A zero word is a no-op.  Per patch the stream holds n_steps rows of `lanes` words, three rows of
// tail words (final flush of slots 0, 1, 2; event field only) and the restart rows (two per restart
// of the lane with the most restarts).
enum : uint32_t {
  STEP_INDEX_MASK = 0x7ffu,
  STEP_LOAD = 1u << 11,
  STEP_COMPUTE = 1u << 12,
  STEP_PRIMARY = 1u << 13,
  STEP_BODY = 1u << 14,
  STEP_NEG = 1u << 15,
  STEP_RESTART = 1u << 16,
  STEP_EVENT_SHIFT = 17,
};
constexpr int32_t kMaxPatchEvents = 32766;

// Compact facet record (patch-local vertex indices, facet orientation): only read by the
// area-weighted vertex normal of flat interior vertices (bending_utils.py:13-34).
struct FacetRec {
  uint16_t a, b, c;
  uint16_t flags;
};
static_assert(sizeof(FacetRec) == 8, "FacetRec layout");

enum : uint16_t { REC_VALID = 1, REC_PRIMARY = 2, REC_BODY = 4 };

struct PackParams {
  int32_t threads = 192;      // lanes of the team that walks one patch (a multiple of 32)
  int32_t max_owned = 448;    // owned vertices per patch
  int32_t max_local = 768;    // owned + halo vertices per patch (shared-memory budget, <= 2047)
  int32_t max_events = 1536;  // event rows per patch (shared-memory budget, <= 32766)
  int32_t trim = 1;           // shrink patches so that their facets fill the lanes' steps evenly
};

struct PackedMesh {
  int32_t nv = 0, nf = 0;
  int32_t n_owned_vertices = 0;  // rows [0, n_owned_vertices) belong to patches
  PackParams params;
  std::vector<PatchHeader> patches;
  std::vector<int32_t> halo_ids;
  std::vector<uint32_t> steps;      // step words, patch after patch
  std::vector<int32_t> step_facet;  // facet row evaluated by each step word (-1: none)
  std::vector<uint16_t> evt_ptr;    // per patch n_owned + 1 offsets: events of owned vertex i are rows [ptr[i], ptr[i+1])
  std::vector<FacetRec> recs;       // compact facet records, patch after patch
  int32_t max_owned = 0, max_local = 0, max_steps = 0, max_events = 0;
  int32_t max_words = 0;       // largest step-word count of a patch (steps + tail rows + restart rows)
  int64_t n_listed = 0;        // facet listings over all patches (>= valid facets)
  int64_t n_valid = 0;         // facets with all indices in range
  int64_t n_strips = 0;        // strips before cutting
  int64_t n_pieces = 0;        // strip pieces after cutting (each begins with a restart step)
  int64_t n_events = 0;        // event rows over all patches
  int64_t n_lane_steps = 0;    // sum over patches of lanes * steps in use (without the no-op padding rows)
  int64_t n_warp_compute = 0;  // (patch, warp, step) triples in which at least one lane evaluates a facet
  int64_t n_gather_groups = 0; // (patch, half-warp, step) groups that load at least one vertex
  int64_t n_gather_excess = 0; // extra shared-memory wavefronts of those gathers (bank conflicts)
};

// body_mask: nf bytes (nonzero = facet in the body) or nullptr (no facet flagged).
// Facets with an index outside [0,nv) are skipped, like surface_energy.f90:57-59; so are facets that name
// one vertex twice: they have zero area and volume and their curvature / gradient contributions vanish
// identically (every term carries a factor e_k = 0 or fK_i - fK_i = 0), in the reference as well.
// Returns 0, or a negative error code (-1 bad arguments, -2 a single vertex needs
// more than max_local local vertices, -3 a single vertex needs more than max_events event rows).
// n_owned_vertices (multi-GPU partitions): only vertex rows [0, n_owned_vertices) are owned
// by patches; rows beyond are ghost vertices of neighbouring partitions, referenced as halo
// only.  Facets without an owned vertex are not listed.  -1 = all vertices are owned.
int pack_patches(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* body_mask,
                 const PackParams& params, PackedMesh& out, int32_t n_owned_vertices = -1);

// Vertex -> corner incidence in CSR form, facet-major order, for the generic
// (stateless) kernels.  corner id = 3*facet + column.
void build_corner_csr(int32_t nv, int32_t nf, const int32_t* tri, std::vector<int32_t>& ptr,
                      std::vector<int32_t>& idx);

}  // namespace ms
