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


def test_strip_packing_properties():
    """Every listed facet is evaluated by exactly one step of its patch; the lanes' steps are filled
    almost completely (a strip piece restarts inside one step); every owned vertex leaves the lanes'
    registers in a handful of event rows."""
    from membrane_solver_b200.synthetic import icosphere

    pos, tri = icosphere(30)
    out = H.emulate(pos, tri, modules=H.MOD_SURFACE, threads=192, max_owned=448, max_local=704)
    p = out["pack"]
    assert p["n_lane_steps"] % 192 == 0 and p["n_lane_steps"] >= p["n_listed"] >= tri.shape[0]
    assert p["n_listed"] >= 0.9 * p["n_lane_steps"]            # lanes nearly full
    assert p["n_listed"] >= 0.9 * 32 * p["n_warp_compute"]     # warps that compute are nearly full
    assert p["n_events"] <= 4 * pos.shape[0]                   # partial sums leave the registers rarely
    assert p["n_pieces"] >= p["n_strips"] >= p["n_patches"]
    area = 0.5 * np.linalg.norm(np.cross(pos[tri[:, 1]] - pos[tri[:, 0]], pos[tri[:, 2]] - pos[tri[:, 0]]), axis=1).sum()
    assert abs(out["area"] - area) <= 1e-13 * area


def test_strip_packing_handles_hubs_soups_and_open_meshes():
    """Shapes that stress the strip builder: a fan around a high-valence hub, disconnected triangles
    (no shared edges: every strip has one facet), a non-manifold edge, degenerate and out-of-range rows."""
    rng = np.random.default_rng(5)
    n = 40
    ang = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    pos = np.vstack([[0.0, 0.0, 0.3], np.stack([np.cos(ang), np.sin(ang), 0.1 * np.sin(3 * ang)], axis=1)])
    tri = np.array([[0, 1 + i, 1 + (i + 1) % n] for i in range(n)], np.int32)
    soup_pos = rng.normal(size=(30, 3))
    soup_tri = np.arange(30, dtype=np.int32).reshape(10, 3) + pos.shape[0]
    fin = np.array([[1, 2, pos.shape[0] + 30], [3, 3, 4], [5, 6, 10_000]], np.int32)  # shares edge (1,2); degenerate; invalid
    pos = np.vstack([pos, soup_pos, [[0.5, 0.5, 1.0]]])
    tri = np.vstack([tri, soup_tri, fin])
    for pack in (dict(threads=32, max_owned=8, max_local=100), dict(threads=64, max_owned=64, max_local=200), dict()):
        out = H.emulate(pos, tri, modules=H.MOD_SURFACE | H.MOD_VOLUME, body_mask=np.ones(tri.shape[0], np.uint8), **pack)
        ok = (tri.max(axis=1) < pos.shape[0]) & (tri[:, 0] != tri[:, 1])
        t = tri[ok]
        nrm = np.cross(pos[t[:, 1]] - pos[t[:, 0]], pos[t[:, 2]] - pos[t[:, 0]])
        area = 0.5 * np.linalg.norm(nrm, axis=1).sum()
        vol = np.einsum("ij,ij->i", np.cross(pos[t[:, 1]], pos[t[:, 2]]), pos[t[:, 0]]).sum() / 6.0
        assert abs(out["area"] - area) <= 1e-13 * area
        assert abs(out["volume"] - vol) <= 1e-12 * max(1.0, abs(vol))
        vg = np.zeros_like(pos)
        np.add.at(vg, t[:, 0], np.cross(pos[t[:, 1]], pos[t[:, 2]]) / 6.0)
        np.add.at(vg, t[:, 1], np.cross(pos[t[:, 2]], pos[t[:, 0]]) / 6.0)
        np.add.at(vg, t[:, 2], np.cross(pos[t[:, 0]], pos[t[:, 1]]) / 6.0)
        assert rel_err(out["volgrad"], vg) <= 1e-12
        sg = np.zeros_like(pos)
        nhat = nrm / np.linalg.norm(nrm, axis=1)[:, None]
        for k in range(3):  # d(area)/dv_k = 1/2 n_hat x (v_{k+2} - v_{k+1})
            e = pos[t[:, (k + 2) % 3]] - pos[t[:, (k + 1) % 3]]
            np.add.at(sg, t[:, k], 0.5 * np.cross(nhat, e))
        assert rel_err(out["grad"], sg) <= 1e-12


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
