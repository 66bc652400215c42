// TEST-ONLY host emulator of the patch kernels.
//
// Runs the SAME per-record / per-vertex bodies (ms_patch_body.cuh) and the SAME
// packer (ms_pack.cpp) as the device path, serially on the host, so that the
// packing, the round schedule and the fp64 math can be validated against the
// oracle in the no-GPU test tier.  It is NOT a product path: nothing under
// membrane_solver_b200/ loads it, and the product fails loudly without a GPU.
#include <cstring>
#include <vector>

#include "../../membrane_solver_b200/csrc/ms_bt.cuh"
#include "../../membrane_solver_b200/csrc/ms_leaflet.cuh"
#include "../../membrane_solver_b200/csrc/ms_patch_body.cuh"

using namespace ms;

namespace {
int local_row(const PackedMesh& pk, const PatchHeader& h, int j) {
  return j < h.n_owned ? h.v_lo + j : pk.halo_ids[size_t(h.halo_off) + size_t(j - h.n_owned)];
}
}  // namespace

extern "C" {

// scalars8: PS_* sums with the volume slot already divided by 6.
// Optional outputs may be null.  Returns the pack return code.
int emul_eval(int32_t nv, int32_t nf, const int32_t* tri, const uint8_t* is_boundary,
              const uint8_t* body_mask, const double* pos, const double* tilts, const double* gamma,
              double gamma_u, const double* kappa, const double* c0, double kappa_u, double c0_u,
              double k_tilt, uint32_t modules, uint32_t flags, int32_t want_grad, int32_t threads,
              int32_t max_owned, int32_t max_local, double* scalars8, double* grad, double* volgrad,
              double* tilt_grad, double* seeds, double* k_vecs, double* a_vor, double* a_eff,
              double* e_vertex, int64_t* pack_stats /*[n_patches,n_slots,n_listed,max_rounds,max_local,lane_conflicts,hw_groups,hw_excess]*/,
              int32_t n_owned /* -1: all */, int32_t phase /* 0: A+B, 1: A only, 2: B only (seeds are input) */) {
  PackParams prm;
  prm.threads = threads;
  prm.max_owned = max_owned;
  prm.max_local = max_local;
  PackedMesh pk;
  const int rc = pack_patches(nv, nf, tri, body_mask, prm, pk, n_owned);
  if (rc) return rc;
  if (pack_stats) {
    pack_stats[0] = int64_t(pk.patches.size());
    pack_stats[1] = int64_t(pk.recs.size());
    pack_stats[2] = pk.n_listed;
    pack_stats[3] = pk.max_rounds;
    pack_stats[4] = pk.max_local;
    pack_stats[5] = pk.n_lane_conflicts;
    pack_stats[6] = pk.n_hw_groups;
    pack_stats[7] = pk.n_hw_excess;
  }
  const bool bt = (modules & MS_MOD_BENDING_TILT) != 0;
  const bool bending = (modules & (MS_MOD_BENDING | MS_MOD_BENDING_TILT)) != 0;
  // the coupling stage needs K, A_vor, A_eff of every vertex
  std::vector<double> kv_store, av_store, ae_store;
  if (bt) {
    if (!k_vecs) { kv_store.assign(3 * size_t(nv), 0.0); k_vecs = kv_store.data(); }
    if (!a_vor) { av_store.assign(size_t(nv), 0.0); a_vor = av_store.data(); }
    if (!a_eff) { ae_store.assign(size_t(nv), 0.0); a_eff = ae_store.data(); }
  }
  if (bt) modules = (modules & ~uint32_t(MS_MOD_BENDING_TILT)) | MS_MOD_BENDING;
  const bool do_tilt = (modules & MS_MOD_TILT) && tilts;
  const bool willmore = (flags & MS_FLAG_WILLMORE) != 0;
  std::vector<double> seed_store(size_t(nv) * kSeedStrideBody, 0.0);
  if (phase == 2 && seeds) std::memcpy(seed_store.data(), seeds, seed_store.size() * sizeof(double));
  double total[PS_COUNT] = {0};
  const int T = threads;

  // ---- pass A -------------------------------------------------------------
  if ((bending || !want_grad) && phase != 2) {
    for (const PatchHeader& h : pk.patches) {
      const int P = h.n_owned, L = h.n_owned + h.n_halo;
      DynamicStrides st;
      st.L = L;
      st.A = (P + 15) / 16 * 16 + kDumpRows;
      const int n_slots = h.n_rounds * T;
      const FacetRec* recs = pk.recs.data() + size_t(h.slot_off);
      std::vector<double> lpos(3 * size_t(L)), t2(size_t(L), 0.0), acc(5 * size_t(st.A), 0.0);
      std::vector<int32_t> bfl(size_t(L), 0);
      for (int j = 0; j < L; ++j) {
        const int row = local_row(pk, h, j);
        for (int k = 0; k < 3; ++k) lpos[3 * size_t(j) + k] = pos[3 * size_t(row) + k];
        bfl[size_t(j)] = is_boundary ? is_boundary[row] : 0;
        if (do_tilt) {
          const double* t = tilts + 3 * size_t(row);
          t2[size_t(j)] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        }
      }
      LocalA loc;
      loc.pos = lpos.data(); loc.bfl = bfl.data(); loc.t2 = do_tilt ? t2.data() : nullptr;
      loc.acc = acc.data(); loc.P = P;
      double sums[PS_COUNT] = {0};
      for (int slot = 0; slot < n_slots; ++slot) {
        const FacetRec rec = recs[slot];
        if (!(rec.flags & REC_VALID)) continue;
        const double gam = gamma ? gamma[pk.slot_facet[size_t(h.slot_off) + slot]] : gamma_u;
        const CornerA c = facet_compute_a(st, rec, gam, loc, modules, k_tilt, sums);
        facet_accumulate_a(st, rec, c, loc, modules);
      }
      if (bending) {
        for (int i = 0; i < P; ++i) {
          const size_t row = size_t(h.v_lo) + i;
          auto normal_of = [&](int v) { return vertex_normal_scan(st, recs, n_slots, lpos.data(), v); };
          const VertexSeed sd = vertex_body_a(st, i, loc, bfl[size_t(i)] != 0, normal_of, kappa ? kappa[row] : kappa_u,
                                              c0 ? c0[row] : c0_u, willmore);
          sums[PS_E_BENDING] += sd.E;
          double* o = seed_store.data() + row * kSeedStrideBody;
          o[0] = sd.fK.x; o[1] = sd.fK.y; o[2] = sd.fK.z; o[3] = sd.fAe; o[4] = sd.fAv;
          if (k_vecs) for (int k = 0; k < 3; ++k) k_vecs[3 * row + k] = acc[size_t(k) * st.A + i];
          if (a_vor) a_vor[row] = acc[3 * size_t(st.A) + i];
          if (a_eff) a_eff[row] = acc[4 * size_t(st.A) + i];
          if (e_vertex) e_vertex[row] = sd.E;
        }
      }
      for (int k = 0; k < PS_COUNT; ++k) total[k] += sums[k];
    }
  }
  double e_bt = 0.0;
  std::vector<double> bt_corner;
  std::vector<int32_t> csr_ptr, csr_idx;
  BtMesh bm;
  if (bt) {
    build_corner_csr(nv, nf, tri, csr_ptr, csr_idx);
    bm.nv = nv; bm.nf = nf; bm.tri = tri; bm.pos = pos; bm.tilts = tilts; bm.is_boundary = is_boundary;
    bm.kappa = kappa; bm.c0 = c0; bm.kappa_u = kappa_u; bm.c0_u = c0_u;
    bm.csr_ptr = csr_ptr.data(); bm.csr_idx = csr_idx.data();
    bt_corner.assign(12 * size_t(nf) + 1, 0.0);
    std::vector<double> base(size_t(nv) + 1, 0.0);
    for (int f = 0; f < nf; ++f) bt_facet_a(bm, f, 1.0, bt_corner.data());
    for (int v = 0; v < nv; ++v) bt_vertex(bm, v, k_vecs, a_vor, a_eff, bt_corner.data(), seed_store.data(), base.data());
    for (int f = 0; f < nf; ++f) e_bt += bt_facet_b(bm, f, base.data(), 1.0, tilt_grad ? bt_corner.data() : nullptr);
    total[PS_E_BENDING] = 0.0;
  }
  if (seeds && phase != 2) std::memcpy(seeds, seed_store.data(), seed_store.size() * sizeof(double));

  // ---- pass B -------------------------------------------------------------
  if (want_grad && phase != 1) {
    const bool scalars_here = !bending;
    const bool do_volume = (modules & MS_MOD_VOLUME) != 0;
    double total_b[PS_COUNT] = {0};
    for (const PatchHeader& h : pk.patches) {
      const int P = h.n_owned, L = h.n_owned + h.n_halo;
      DynamicStrides st;
      st.L = L;
      st.A = (P + 15) / 16 * 16 + kDumpRows;
      const int n_slots = h.n_rounds * T;
      const FacetRec* recs = pk.recs.data() + size_t(h.slot_off);
      std::vector<double> lpos(3 * size_t(L)), lseed(size_t(kSeedStrideBody) * size_t(L), 0.0),
          t2(size_t(L), 0.0), acc(6 * size_t(st.A), 0.0), accAb(size_t(st.A), 0.0);
      std::vector<int32_t> bfl(size_t(L), 0);
      for (int j = 0; j < L; ++j) {
        const int row = local_row(pk, h, j);
        for (int k = 0; k < 3; ++k) lpos[3 * size_t(j) + k] = pos[3 * size_t(row) + k];
        for (int k = 0; k < kSeedStrideBody; ++k)
          lseed[size_t(kSeedStrideBody) * j + k] = seed_store[size_t(row) * kSeedStrideBody + k];
        bfl[size_t(j)] = is_boundary ? is_boundary[row] : 0;
        if (do_tilt) {
          const double* t = tilts + 3 * size_t(row);
          t2[size_t(j)] = t[0] * t[0] + t[1] * t[1] + t[2] * t[2];
        }
      }
      LocalB loc;
      loc.pos = lpos.data(); loc.seed = lseed.data(); loc.bfl = bfl.data();
      loc.t2 = do_tilt ? t2.data() : nullptr;
      loc.acc = acc.data(); loc.accAb = accAb.data(); loc.P = P;
      double sums[PS_COUNT] = {0};
      for (int slot = 0; slot < n_slots; ++slot) {
        const FacetRec rec = recs[slot];
        if (!(rec.flags & REC_VALID)) continue;
        const double gam = gamma ? gamma[pk.slot_facet[size_t(h.slot_off) + slot]] : gamma_u;
        const FacetOutB o = bending
            ? facet_compute_b<true>(st, rec, gam, loc, modules, flags, k_tilt, scalars_here, sums)
            : facet_compute_b<false>(st, rec, gam, loc, modules, flags, k_tilt, scalars_here, sums);
        facet_accumulate_b(st, rec, o, loc, do_volume, do_tilt);
      }
      for (int i = 0; i < P; ++i)
        for (int k = 0; k < 3; ++k) {
          const size_t o = (size_t(h.v_lo) + i) * 3 + size_t(k);
          if (grad) grad[o] = acc[size_t(k) * st.A + i];
          if (volgrad && do_volume) volgrad[o] = (1.0 / 6.0) * acc[size_t(3 + k) * st.A + i];
          if (tilt_grad && do_tilt) tilt_grad[o] = k_tilt * tilts[o] * accAb[size_t(i)];
        }
      for (int k = 0; k < PS_COUNT; ++k) total_b[k] += sums[k];
    }
    if (scalars_here)
      for (int k = 0; k < PS_COUNT; ++k) total[k] = total_b[k];
    else
      total[PS_E_TILT] = total_b[PS_E_TILT];
  }
  if (bt && tilt_grad) {  // tilt gradient of the coupling term: fixed-order CSR gather, added to the tilt module's
    for (int v = 0; v < nv; ++v)
      for (int d = 0; d < 3; ++d) {
        double acc = 0.0;
        for (int j = csr_ptr[size_t(v)]; j < csr_ptr[size_t(v) + 1]; ++j) acc += bt_corner[3 * size_t(csr_idx[size_t(j)]) + d];
        double& o = tilt_grad[3 * size_t(v) + d];
        o = (do_tilt && want_grad) ? o + acc : acc;
      }
  }
  total[PS_E_BENDING_TILT] = e_bt;
  total[PS_VOLUME6] /= 6.0;
  for (int k = 0; k < 8; ++k) scalars8[k] = total[k];
  return 0;
}

// Leaflet modules (ms_leaflet.cuh): the three sweeps and the gathers in the device path's order.
// grad / tilt_grad (may be null) are ACCUMULATED into.  energies3 = {E_bending_tilt, E_tilt, E_tilt_smoothness}.
int emul_leaflet(int32_t nv, int32_t nf, const int32_t* tri, const double* pos, const double* tilts,
                 const uint8_t* keep, const uint8_t* is_boundary, const uint8_t* interior, const uint8_t* base_zero,
                 const double* kappa, const double* c0, double kappa_u, double c0_u, const double* row_weight,
                 const uint8_t* consistent, int32_t consistent_u, double k_tilt, double k_smooth, double sign,
                 int32_t with_bt, int32_t with_tilt, int32_t with_smooth, double* grad, double* tilt_grad,
                 double* energies3) {
  std::vector<int32_t> csr_ptr, csr_idx;
  build_corner_csr(nv, nf, tri, csr_ptr, csr_idx);
  LeafletMesh m{nv, nf, tri, pos, tilts, keep, is_boundary, interior, base_zero, kappa, c0, kappa_u, c0_u,
                row_weight, consistent, consistent_u, k_tilt, k_smooth, sign, csr_ptr.data(), csr_idx.data()};
  std::vector<double> corner(size_t(3 * kLfCornerA) * size_t(nf) + 1, 0.0), vbuf(size_t(kLfVertex) * size_t(nv) + 1, 0.0);
  std::vector<double> cs(9 * size_t(nf) + 1, 0.0), ct(9 * size_t(nf) + 1, 0.0);
  if (with_bt) {
    for (int f = 0; f < nf; ++f) lf_facet_a(m, f, corner.data());
    for (int v = 0; v < nv; ++v) lf_vertex(m, v, corner.data(), vbuf.data());
  }
  double e_bt = 0.0, e_tilt = 0.0, e_smooth = 0.0;
  for (int f = 0; f < nf; ++f) {
    const LfEnergies e = lf_facet_b(m, f, vbuf.data(), with_bt != 0, with_tilt != 0, with_smooth != 0,
                                    grad ? cs.data() : nullptr, tilt_grad ? ct.data() : nullptr);
    e_bt += e.e_bt;
    e_tilt += e.e_tilt;
    e_smooth += e.e_smooth;
  }
  for (int v = 0; v < nv; ++v)
    for (int d = 0; d < 3; ++d) {
      double a = 0.0, b = 0.0;
      for (int j = csr_ptr[size_t(v)]; j < csr_ptr[size_t(v) + 1]; ++j) {
        a += cs[3 * size_t(csr_idx[size_t(j)]) + d];
        b += ct[3 * size_t(csr_idx[size_t(j)]) + d];
      }
      if (grad) grad[3 * size_t(v) + d] += a;
      if (tilt_grad) tilt_grad[3 * size_t(v) + d] += b;
    }
  energies3[0] = e_bt;
  energies3[1] = e_tilt;
  energies3[2] = e_smooth;
  return 0;
}

}  // extern "C"
