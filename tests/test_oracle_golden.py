"""Pin the CPU oracle against golden vectors produced by the real reference.

The fixtures come from ``tests/golden/generate_golden.py`` (which imports the
unmodified reference).  Both oracle variants are checked: pure NumPy, and NumPy
with the C restatement of the Fortran kernels injected through the loader seam.
"""

import numpy as np
import pytest

from oracle import ckernels
from oracle import ref_modules as ref
from ms_test_helpers import golden_ids, golden_module_files, rel_err

TOL = 1e-12
BENDING_TAGS = {
    "helfrich_analytic": ("helfrich", "analytic"),
    "helfrich_c0": ("helfrich", "analytic"),
    "helfrich_approx": ("helfrich", "approx"),
    "willmore_analytic": ("willmore", "analytic"),
}


@pytest.fixture(params=["numpy", "c_kernels"])
def variant(request):
    ref.use_c_kernels(ckernels if request.param == "c_kernels" else None)
    yield request.param
    ref.use_c_kernels(None)


def _close(a, b, tol=TOL):
    assert rel_err(a, b) <= tol, rel_err(a, b)


def test_kernel_vectors(kernels_golden, variant):
    g = kernels_golden
    for n in (4, 17):
        gu, gv = ref.grad_cotan(g[f"gc{n}_u"], g[f"gc{n}_v"])
        _close(gu, g[f"gc{n}_gu"])
        _close(gv, g[f"gc{n}_gv"])
        tu, tv = ref.grad_triangle_area(g[f"gc{n}_u"], g[f"gc{n}_v"])
        _close(tu, g[f"gc{n}_tu"])
        _close(tv, g[f"gc{n}_tv"])
    gu, gv = ref.grad_cotan(g["gcd_u"], g["gcd_v"])
    assert np.array_equal(gu[2], np.zeros(3)) and np.array_equal(gu[0], np.zeros(3))
    _close(gu, g["gcd_gu"])
    _close(gv, g["gcd_gv"])
    _close(ref.beltrami_laplacian(g["lap_w"], g["lap_tri"], g["lap_field"]), g["lap_out"])
    div, area, g0, g1, g2 = ref.p1_triangle_divergence(g["p1_pos"], g["p1_tilts"], g["p1_tri"])
    for a, b in ((div, "p1_div"), (area, "p1_area"), (g0, "p1_g0"), (g1, "p1_g1"), (g2, "p1_g2")):
        _close(a, g[b])
    k, a, w, va0, va1, va2 = ref.curvature_data(g["cd_pos"], g["cd_tri"])
    for x, name in ((k, "cd_k"), (a, "cd_a"), (w, "cd_w"), (va0, "cd_va0"), (va1, "cd_va1"), (va2, "cd_va2")):
        _close(x, g[name])


@pytest.mark.parametrize("path", golden_module_files(), ids=golden_ids())
def test_module_vectors(path, variant):
    g = dict(np.load(path))
    pos, tri = g["pos"], g["tri"]
    grad = np.zeros_like(pos)
    e = ref.surface_energy_and_gradient(pos, tri, g["gamma"], grad)
    assert abs(e - g["E_surface"]) <= TOL * max(1.0, abs(g["E_surface"]))
    _close(grad, g["g_surface"])

    if "g_volume" in g:
        for i in range(g["g_volume"].shape[0]):
            rows = g[f"body_rows_{i}"]
            gc = np.zeros_like(pos)
            ref.accumulate_volume_gradient(pos, tri[rows], gc, 1.0)
            _close(gc, g["g_volume"][i])
            v = ref.body_volume(pos, tri[rows])
            assert abs(v - g["volumes"][i]) <= TOL * max(1.0, abs(g["volumes"][i]))

    k, a, w, _, _, _ = ref.curvature_data(pos, tri)
    _close(k, g["k_vecs"])
    _close(a, g["a_vor"])
    _close(w, g["weights"])
    a_eff, va = ref.effective_areas(pos, tri, w, g["is_boundary"])
    _close(a_eff, g["a_eff"])
    _close(va, g["va_eff"])
    _close(ref.vertex_normals(pos, tri), g["normals"])

    for tag, (model, mode) in BENDING_TAGS.items():
        kappa, c0 = g[f"param_{tag}"]
        grad = np.zeros_like(pos)
        e = ref.bending_energy_and_gradient(pos, tri, kappa, c0, g["is_boundary"], grad, model, mode)
        assert abs(e - g[f"E_bending_{tag}"]) <= TOL * max(1.0, abs(g[f"E_bending_{tag}"]))
        _close(grad, g[f"g_bending_{tag}"], 1e-11)
        ev = ref.bending_energy_per_vertex(pos, tri, kappa, c0, g["is_boundary"], model)
        _close(ev, g[f"Ev_bending_{tag}"])
        if model == "helfrich":
            grad = np.zeros_like(pos)
            tg = np.zeros_like(pos)
            e = ref.bending_tilt_energy_and_gradient(pos, tri, g["tilts"], kappa, c0, g["is_boundary"],
                                                     grad, tg, mode)
            assert abs(e - g[f"E_bending_tilt_{tag}"]) <= TOL * max(1.0, abs(e))
            _close(grad, g[f"g_bending_tilt_{tag}"], 1e-11)
            _close(tg, g[f"tg_bending_tilt_{tag}"])
            tg2 = np.zeros_like(pos)
            ref.bending_tilt_energy_and_gradient(pos, tri, g["tilts"], kappa, c0, g["is_boundary"],
                                                 None, tg2, mode)
            _close(tg2, g[f"tgonly_bending_tilt_{tag}"])

    grad = np.zeros_like(pos)
    tg = np.zeros_like(pos)
    e = ref.tilt_energy_and_gradient(pos, tri, g["tilts"], float(g["k_tilt"]), grad, tg)
    assert abs(e - g["E_tilt"]) <= TOL * max(1.0, abs(g["E_tilt"]))
    _close(grad, g["g_tilt"])
    _close(tg, g["tg_tilt"])


def test_boundary_mask_from_triangles():
    for path in golden_module_files():
        g = dict(np.load(path))
        mask = ref.boundary_mask_from_triangles(g["tri"], g["pos"].shape[0])
        assert np.array_equal(mask, g["is_boundary"]), path


def test_minimizer_vectors(minimizer_golden):
    """Total energy and KKT-projected gradient through the reference Minimizer
    (runtime/minimizer.py:941-992, constraint_manager.py:294-301)."""
    g = minimizer_golden
    # cube: surface + volume penalty
    pos, tri = g["cube_pos"], g["cube_tri"]
    grad = np.zeros_like(pos)
    e = ref.surface_energy_and_gradient(pos, tri, g["cube_gamma"], grad)
    _, ev = ref.volume_penalty_energy_and_gradient(
        pos, tri[g["cube_body_rows_0"]], float(g["cube_kvol"]), float(g["cube_body_target_0"]), grad)
    grad[g["cube_fixed"]] = 0.0
    assert abs(e + ev - g["cube_E"]) <= TOL * abs(g["cube_E"])
    assert abs(e - g["cube_E_surface"]) <= TOL * abs(e)
    assert abs(ev - g["cube_E_volume"]) <= 1e-10 * max(abs(ev), 1e-3)
    _close(grad, g["cube_g"])
    # bending cube: bending + lagrange volume constraint
    pos, tri = g["bcube_pos"], g["bcube_tri"]
    grad = np.zeros_like(pos)
    e = ref.bending_energy_and_gradient(pos, tri, float(g["bcube_kappa"]), float(g["bcube_c0"]),
                                        g["bcube_is_boundary"], grad, str(g["bcube_model"]), "analytic")
    gc = np.zeros_like(pos)
    ref.accumulate_volume_gradient(pos, tri[g["bcube_body_rows_0"]], gc, 1.0)
    ref.kkt_project_single(grad, gc)
    grad[g["bcube_fixed"]] = 0.0
    assert abs(e - g["bcube_E"]) <= TOL * abs(g["bcube_E"])
    _close(grad, g["bcube_g"], 1e-11)
