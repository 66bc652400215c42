"""CPU tier: the packer (rounds, lane placement, partitions) and the per-facet device code,
compiled for the host (tests/emul), against the golden vectors of the real reference.
This validates the kernel *logic* without a GPU; the GPU tier validates the kernels."""

import numpy as np
import pytest

import ms_test_helpers as H
from ms_test_helpers import golden_ids, golden_module_files, rel_err

TOL = 1e-12
BENDING_TAGS = {"helfrich_analytic": (0, 0), "helfrich_c0": (0, 0), "helfrich_approx": (0, 2),
                "willmore_analytic": (1, 0)}
PACKS = [dict(), dict(threads=32, max_owned=16, max_local=120)]


@pytest.mark.parametrize("path", golden_module_files(), ids=golden_ids())
@pytest.mark.parametrize("pack", PACKS, ids=["default", "tiny-patches"])
def test_emulated_patch_path_vs_reference_golden(path, pack):
    g = dict(np.load(path))
    pos, tri = g["pos"], g["tri"]
    nf = tri.shape[0]
    body = np.zeros(nf, np.uint8)
    if "body_rows_0" in g:
        body[g["body_rows_0"]] = 1
    gamma = g["gamma"]
    out = H.emulate(pos, tri, modules=H.MOD_SURFACE | H.MOD_VOLUME, is_boundary=g["is_boundary"], body_mask=body,
                    gamma=gamma, **pack)
    assert abs(out["E_surface"] - float(g["E_surface"])) <= TOL * max(1.0, abs(float(g["E_surface"])))
    assert rel_err(out["grad"], g["g_surface"]) <= TOL
    if "g_volume" in g:
        assert abs(out["volume"] - float(g["volumes"][0])) <= TOL * max(1.0, abs(float(g["volumes"][0])))
        assert rel_err(out["volgrad"], g["g_volume"][0]) <= TOL
    for tag, (wil, apx) in BENDING_TAGS.items():
        kappa, c0 = g[f"param_{tag}"]
        out = H.emulate(pos, tri, modules=H.MOD_BENDING, flags=wil | apx, is_boundary=g["is_boundary"],
                        kappa_u=float(kappa), c0_u=float(c0), **pack)
        e = float(g[f"E_bending_{tag}"])
        assert abs(out["E_bending"] - e) <= TOL * max(1.0, abs(e)), tag
        grad = out["grad"]
        if apx:
            grad[g["is_boundary"]] = 0.0
        assert rel_err(grad, g[f"g_bending_{tag}"]) <= 2e-12, tag
        assert rel_err(out["e_vertex"], g[f"Ev_bending_{tag}"]) <= TOL
    assert rel_err(out["k_vecs"], g["k_vecs"]) <= TOL
    assert rel_err(out["a_vor"], g["a_vor"]) <= TOL
    assert rel_err(out["a_eff"], g["a_eff"]) <= TOL
    out = H.emulate(pos, tri, modules=H.MOD_TILT, tilts=g["tilts"], k_tilt=float(g["k_tilt"]), **pack)
    assert abs(out["E_tilt"] - float(g["E_tilt"])) <= TOL * max(1.0, abs(float(g["E_tilt"])))
    assert rel_err(out["grad"], g["g_tilt"]) <= TOL
    assert rel_err(out["tilt_grad"], g["tg_tilt"]) <= TOL


def test_lane_placement_properties():
    """Every valid facet appears exactly once per listing patch; slots come in whole rounds;
    the bank-aware placement leaves few half-warp residue clashes."""
    from membrane_solver_b200.synthetic import icosphere

    pos, tri = icosphere(30)
    out = H.emulate(pos, tri, modules=H.MOD_SURFACE)
    p = out["pack"]
    assert p["n_slots"] % 96 == 0 and p["n_slots"] >= p["n_listed"] >= tri.shape[0]
    # bank-aware placement: few half-warp gather groups need an extra shared-memory wavefront
    assert p["hw_excess"] <= 0.6 * p["hw_groups"]  # unplaced (random) lanes give ~2.0
    area = 0.5 * np.linalg.norm(np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]]), axis=1).sum()
    assert abs(out["area"] - area) <= 1e-13 * area


@pytest.mark.parametrize("path", golden_module_files(), ids=golden_ids())
def test_emulated_bending_tilt_vs_reference_golden(path):
    """Single-field bending_tilt (bending_tilt.py:151-482): energy, shape gradient (div treated as
    constant), exact tilt gradient, and the tilt-only evaluation."""
    g = dict(np.load(path))
    pos, tri = g["pos"], g["tri"]
    for tag, (wil, apx) in BENDING_TAGS.items():
        if wil:
            continue
        kappa, c0 = g[f"param_{tag}"]
        out = H.emulate(pos, tri, modules=H.MOD_BENDING_TILT, flags=apx, is_boundary=g["is_boundary"],
                        tilts=g["tilts"], kappa_u=float(kappa), c0_u=float(c0))
        e = float(g[f"E_bending_tilt_{tag}"])
        assert abs(out["E_bending_tilt"] - e) <= TOL * max(1.0, abs(e)), tag
        grad = out["grad"]
        if apx:
            grad[g["is_boundary"]] = 0.0
        assert rel_err(grad, g[f"g_bending_tilt_{tag}"]) <= 2e-12, tag
        assert rel_err(out["tilt_grad"], g[f"tg_bending_tilt_{tag}"]) <= TOL, tag
        only = H.emulate(pos, tri, modules=H.MOD_BENDING_TILT, flags=apx, want_grad=False, is_boundary=g["is_boundary"],
                         tilts=g["tilts"], kappa_u=float(kappa), c0_u=float(c0))
        assert abs(only["E_bending_tilt"] - e) <= TOL * max(1.0, abs(e)), tag
        assert rel_err(only["tilt_grad"], g[f"tgonly_bending_tilt_{tag}"]) <= TOL, tag


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_odd_topologies_against_the_oracle(seed):
    """Packer and per-facet code on meshes the benchmark shapes never produce: a surface with unused vertices, a few
    facets that name a vertex twice (zero area: never listed), a fan of valence 14 (14 rounds of 96 slots: close to the
    slot capacity of a patch; ms_ctx_set_topology narrows the rounds beyond that), a
    duplicated facet and two components, in random vertex order and with several pack geometries.  Surface and volume
    modules (they are defined facet by facet, whatever the connectivity) against the oracle at 1e-12."""
    pos, tri, gamma, body, want = H.odd_mesh(seed, nfan=14)
    nf = tri.shape[0]
    want_e, want_g, want_v, want_vg = want
    for pack in (dict(), dict(threads=32, max_owned=16, max_local=120), dict(threads=64, max_owned=48, max_local=200)):
        out = H.emulate(pos, tri, modules=H.MOD_SURFACE | H.MOD_VOLUME, body_mask=body, gamma=gamma, **pack)
        assert abs(out["E_surface"] - want_e) <= TOL * abs(want_e), pack
        assert rel_err(out["grad"], want_g) <= TOL, pack
        assert abs(out["volume"] - want_v) <= TOL * max(1.0, abs(want_v)), pack
        assert rel_err(out["volgrad"], want_vg) <= TOL, pack
        assert out["pack"]["n_listed"] >= nf - 2          # the two degenerate facets are never listed
