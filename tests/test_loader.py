"""The KernelSpec seam (``membrane_solver_b200/loader.py`` = ``fortran_kernels/loader.py:15-20,30,85,139,193,247``).

CPU tier: the getters resolve (the library loads without a device), return ``KernelSpec(func, expects_transpose=False)``
and enforce the reference's strict no-copy contract (``surface.py:137-155``) before anything reaches the device.
GPU tier: every kernel, called through its ``KernelSpec`` with the reference's native-layout arrays, against the
inputs / outputs of the reference's own kernel-parity tests (``tests/test_fortran_kernels.py:46-376``, stored in
``tests/golden/kernels.npz``), and injected the way the reference's tests inject kernels
(``tests/test_surface_nocopy_guardrails.py:49-53``: the getter of the consuming module is replaced)."""

import types

import numpy as np
import pytest

from ms_test_helpers import rel_err

from membrane_solver_b200 import _lib as L
from membrane_solver_b200 import loader

TOL = 1e-12
GETTERS = ("get_surface_energy_kernel", "get_bending_grad_cotan_kernel", "get_bending_laplacian_kernel",
           "get_tilt_divergence_kernel", "get_tilt_curvature_kernel")


def test_getters_resolve_without_a_device():
    for name in GETTERS:
        spec = getattr(loader, name)()
        assert isinstance(spec, loader.KernelSpec) and callable(spec.func) and spec.expects_transpose is False


def test_strict_nocopy_contract_is_enforced_before_the_device(kernels_golden):
    g = kernels_golden
    pos, tri, gamma = (np.ascontiguousarray(g[k]) for k in ("sf_pos", "sf_tri", "sf_gamma"))
    grad = np.zeros_like(pos)
    k = loader.get_surface_energy_kernel().func
    with pytest.raises(TypeError):  # wrong dtype: no silent conversion
        k(pos.astype(np.float32), tri, gamma, grad)
    with pytest.raises(TypeError):
        k(pos, tri.astype(np.int64), gamma, grad)
    with pytest.raises(ValueError):  # F-ordered positions are not the native (n,3) layout
        k(np.asfortranarray(pos), tri, gamma, grad)
    with pytest.raises(ValueError):
        k(pos, tri, gamma, grad[:, ::-1])
    with pytest.raises(TypeError):
        k(pos.tolist(), tri, gamma, grad)
    lap = loader.get_bending_laplacian_kernel().func
    w, t, f = (np.ascontiguousarray(g[x]) for x in ("lap_w", "lap_tri", "lap_field"))
    with pytest.raises(ValueError):
        lap(w, t, f, np.zeros(f.shape[::-1]).T)


@pytest.fixture
def gpu():
    if L.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tier must run on the B200 box")
    return L


@pytest.mark.gpu
def test_kernelspecs_vs_reference_kernel_vectors(gpu, kernels_golden):
    from oracle import ref_modules as ref

    g = kernels_golden
    # surface_energy_and_gradient: accumulates into grad, returns the energy
    pos, tri, gamma = (np.ascontiguousarray(g[k]) for k in ("sf_pos", "sf_tri", "sf_gamma"))
    want_g = np.full_like(pos, 0.5)
    want_e = ref.surface_energy_and_gradient(pos, tri, gamma, want_g)
    grad = np.full_like(pos, 0.5)
    e = loader.get_surface_energy_kernel().func(pos, tri, gamma, grad, 1)
    assert abs(e - want_e) <= TOL * abs(want_e) and rel_err(grad, want_g) <= TOL
    grad1 = np.full_like(pos, 0.5)  # one-based rows (zero_based=0), as the Fortran callers may pass them
    e1 = loader.get_surface_energy_kernel().func(pos, np.ascontiguousarray(tri + 1), gamma, grad1, 0)
    assert e1 == e and np.array_equal(grad1, grad)
    # grad_cotan_batch: three input sets of the reference's tests
    for tag in ("gc4", "gc17", "gcd"):
        u, v = np.ascontiguousarray(g[f"{tag}_u"]), np.ascontiguousarray(g[f"{tag}_v"])
        gu, gv = np.empty_like(u), np.empty_like(v)
        loader.get_bending_grad_cotan_kernel().func(u, v, gu, gv)
        assert rel_err(gu, g[f"{tag}_gu"]) <= TOL and rel_err(gv, g[f"{tag}_gv"]) <= TOL, tag
    # apply_beltrami_laplacian: out is overwritten
    w, t, f = (np.ascontiguousarray(g[x]) for x in ("lap_w", "lap_tri", "lap_field"))
    out = np.full_like(f, np.nan)
    loader.get_bending_laplacian_kernel().func(w, t, f, out, 1)
    assert rel_err(out, g["lap_out"]) <= TOL
    # p1_triangle_divergence
    pos, tl, tri = (np.ascontiguousarray(g[k]) for k in ("p1_pos", "p1_tilts", "p1_tri"))
    nf = tri.shape[0]
    div, area = np.empty(nf), np.empty(nf)
    g0, g1, g2 = np.empty((nf, 3)), np.empty((nf, 3)), np.empty((nf, 3))
    loader.get_tilt_divergence_kernel().func(pos, tl, tri, div, area, g0, g1, g2, 1)
    for a, name in ((div, "p1_div"), (area, "p1_area"), (g0, "p1_g0"), (g1, "p1_g1"), (g2, "p1_g2")):
        assert rel_err(a, g[name]) <= TOL, name
    # compute_curvature_data, with and without the optional corner areas
    pos, tri = np.ascontiguousarray(g["cd_pos"]), np.ascontiguousarray(g["cd_tri"])
    nv, nf = pos.shape[0], tri.shape[0]
    k, a, w = np.empty((nv, 3)), np.empty(nv), np.empty((nf, 3))
    va = [np.empty(nf) for _ in range(3)]
    loader.get_tilt_curvature_kernel().func(pos, tri, k, a, w, 1, va[0], va[1], va[2])
    for x, name in ((k, "cd_k"), (a, "cd_a"), (w, "cd_w"), (va[0], "cd_va0"), (va[1], "cd_va1"), (va[2], "cd_va2")):
        assert rel_err(x, g[name]) <= TOL, name
    k2 = np.empty((nv, 3))
    loader.get_tilt_curvature_kernel().func(pos, tri, k2, a, w, 1)
    assert np.array_equal(k2, k)


@pytest.mark.gpu
def test_injection_recipe_of_the_reference_tests(gpu, kernels_golden, monkeypatch):
    """A consumer that resolves its kernel through a module-level getter -- the shape of modules/energy/surface.py
    -- gets the B200 kernel when the getter is replaced, exactly as tests/test_surface_nocopy_guardrails.py:49-53
    replaces it with a fake; a non-zero return code of the C ABI surfaces as B200Error (no silent fallback)."""
    from oracle import ref_modules as ref

    consumer = types.SimpleNamespace(get_surface_energy_kernel=lambda: None)

    def compute(positions, tri_rows, gamma, grad_arr):  # what surface.compute_energy_and_gradient_array does
        spec = consumer.get_surface_energy_kernel()
        if spec is None:
            raise RuntimeError("no kernel")
        assert spec.expects_transpose is False
        return spec.func(positions, tri_rows, gamma, grad_arr, 1)

    monkeypatch.setattr(consumer, "get_surface_energy_kernel", loader.get_surface_energy_kernel)
    g = kernels_golden
    pos, tri, gamma = (np.ascontiguousarray(g[k]) for k in ("sf_pos", "sf_tri", "sf_gamma"))
    grad = np.zeros_like(pos)
    want_g = np.zeros_like(pos)
    want_e = ref.surface_energy_and_gradient(pos, tri, gamma, want_g)
    assert abs(compute(pos, tri, gamma, grad) - want_e) <= TOL * abs(want_e)
    assert rel_err(grad, want_g) <= TOL
    with pytest.raises(L.B200Error):  # bad sizes and a null result pointer: refused by the ABI, not papered over
        bad = np.empty((2, 3))
        L.check(L.lib().ms_surface_energy_and_gradient(-1, tri.shape[0], L.dptr(pos), L.iptr(tri), L.dptr(gamma),
                                                       L.dptr(bad), None, 1))
