"""The reference-facing plugin layer (modules/energy twins, evaluation manager, residency),
driven through the array contract of SURVEY.md section 8b against the golden vectors of the real
reference.  CPU tier: the device is the host emulator (tests/fake_device.py); GPU tier
(``-m gpu``): the same tests through libms_b200.so."""

import numpy as np
import pytest

import ms_test_helpers as H
from ms_test_helpers import golden_ids, golden_module_files, rel_err

from membrane_solver_b200 import _lib as L
from membrane_solver_b200.geometry.array_mesh import ArrayBody, ArrayMesh, GlobalParams, ParamResolver
from membrane_solver_b200.modules.constraints import volume as volume_constraint
from membrane_solver_b200.modules.energy import bending, bending_tilt, surface, tilt, volume
from membrane_solver_b200.runtime import device_state
from membrane_solver_b200.runtime.energy_manager import EnergyModuleManager
from membrane_solver_b200.runtime.evaluation_manager import EvaluationManager

TOL = 1e-12
BENDING_TAGS = {"helfrich_analytic": ("helfrich", "analytic"), "helfrich_c0": ("helfrich", "analytic"),
                "helfrich_approx": ("helfrich", "approx"), "willmore_analytic": ("willmore", "analytic")}


def _backend_params():
    return [pytest.param("emulator", id="emulator"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=_backend_params())
def backend(request, monkeypatch):
    if request.param == "emulator":
        from fake_device import FakeDeviceMesh

        monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    else:
        if L.device_count() < 1:
            pytest.fail("no CUDA device visible: the gpu tier must run on the B200 box")
    return request.param


def _mesh(g, **gp):
    params = GlobalParams(surface_tension=1.0, volume_constraint_mode="lagrange", volume_stiffness=1000.0,
                          bending_modulus=0.0, **gp)
    bodies = {}
    if "body_rows_0" in g:
        t = float(g["body_target_0"])
        bodies[0] = ArrayBody(g["body_rows_0"], target_volume=None if np.isnan(t) else t)
    mesh = ArrayMesh(g["pos"], g["tri"], global_params=params, facet_params={"surface_tension": g["gamma"]},
                     bodies=bodies, fixed=g["fixed"], tilts=g["tilts"])
    return mesh, params, ParamResolver(params)


def _close(a, b, tol=TOL):
    assert abs(a - b) <= tol * max(1.0, abs(b)), (a, b)


@pytest.mark.parametrize("path", golden_module_files(), ids=golden_ids())
def test_plugin_modules_vs_reference_golden(backend, path):
    g = dict(np.load(path))
    mesh, gp, res = _mesh(g)
    pos, idx = mesh.positions_view(), mesh.vertex_index_to_row
    assert set(mesh.boundary_vertex_ids) == set(np.nonzero(g["is_boundary"])[0].tolist())

    grad = np.full_like(pos, 0.25)  # plugins accumulate into the caller's array
    e = surface.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx, grad_arr=grad)
    _close(e, float(g["E_surface"]))
    assert rel_err(grad - 0.25, g["g_surface"]) <= TOL
    _close(surface.compute_energy_array(mesh, gp, positions=pos, index_map=idx), float(g["E_surface"]))
    e_d, g_d = surface.compute_energy_and_gradient(mesh, gp, res)
    _close(e_d, float(g["E_surface"]))
    assert rel_err(np.array([g_d[v] for v in range(len(g_d))]), g["g_surface"]) <= TOL

    if "g_volume" in g:
        gcs = volume_constraint.constraint_gradients_array(mesh, gp, positions=pos, index_map=idx)
        if mesh.bodies[0].target_volume is None:
            assert gcs is None  # bodies without a target volume are not constrained
            mesh.bodies[0].target_volume = 0.9 * float(g["volumes"][0])
            gcs = volume_constraint.constraint_gradients_array(mesh, gp, positions=pos, index_map=idx)
        assert len(gcs) == 1 and rel_err(gcs[0], g["g_volume"][0]) <= TOL
        gp["volume_constraint_mode"] = "penalty"
        assert volume_constraint.constraint_gradients_array(mesh, gp, positions=pos, index_map=idx) is None
        grad = np.zeros_like(pos)
        e = volume.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx, grad_arr=grad)
        d = float(g["volumes"][0]) - mesh.bodies[0].target_volume
        _close(e, 0.5 * 1000.0 * d * d, 1e-11)
        assert rel_err(grad, 1000.0 * d * g["g_volume"][0]) <= 1e-10 if abs(d) > 1e-13 else True
        gp["volume_constraint_mode"] = "lagrange"
        assert volume.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx,
                                                        grad_arr=grad) == 0.0

    for tag, (model, mode) in BENDING_TAGS.items():
        kappa, c0 = (float(x) for x in g[f"param_{tag}"])
        gp.update(bending_modulus=kappa, spontaneous_curvature=c0, bending_energy_model=model,
                  bending_gradient_mode=mode)
        grad = np.zeros_like(pos)
        e = bending.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx, grad_arr=grad)
        _close(e, float(g[f"E_bending_{tag}"]))
        assert rel_err(grad, g[f"g_bending_{tag}"]) <= 2e-12, tag
        ev = bending.compute_energy_array(mesh, gp, pos, idx)
        assert rel_err(ev, g[f"Ev_bending_{tag}"]) <= TOL
        _close(bending.compute_total_energy(mesh, gp, pos, idx), float(g[f"E_bending_{tag}"]))
    gp.update(bending_modulus=0.0)
    assert bending.compute_total_energy(mesh, gp, pos, idx) == 0.0

    gp["tilt_rigidity"] = float(g["k_tilt"])
    grad, tg = np.zeros_like(pos), np.ones_like(pos)
    e = tilt.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx, grad_arr=grad,
                                               tilt_grad_arr=tg)
    _close(e, float(g["E_tilt"]))
    assert rel_err(grad, g["g_tilt"]) <= TOL
    assert rel_err(tg - 1.0, g["tg_tilt"]) <= TOL
    _close(tilt.compute_energy_array(mesh, gp, res, positions=pos, index_map=idx), float(g["E_tilt"]))
    with pytest.raises(ValueError):
        tilt.compute_energy_array(mesh, gp, res, positions=pos, index_map=idx, tilts=np.zeros((3, 3)))

    # bending_tilt (single field): energy, shape gradient, exact tilt gradient, tilt-only evaluation
    for tag, (model, mode) in BENDING_TAGS.items():
        if model != "helfrich":
            continue
        kappa, c0 = (float(x) for x in g[f"param_{tag}"])
        gp.update(bending_modulus=kappa, spontaneous_curvature=c0, bending_energy_model=model,
                  bending_gradient_mode=mode)
        grad, tg = np.zeros_like(pos), np.zeros_like(pos)
        e = bending_tilt.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx,
                                                           grad_arr=grad, tilt_grad_arr=tg)
        _close(e, float(g[f"E_bending_tilt_{tag}"]))
        assert rel_err(grad, g[f"g_bending_tilt_{tag}"]) <= 2e-12, tag
        assert rel_err(tg, g[f"tg_bending_tilt_{tag}"]) <= TOL, tag
        tg = np.zeros_like(pos)
        e = bending_tilt.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx,
                                                           grad_arr=None, tilts=g["tilts"], tilt_grad_arr=tg)
        _close(e, float(g[f"E_bending_tilt_{tag}"]))
        assert rel_err(tg, g[f"tgonly_bending_tilt_{tag}"]) <= TOL, tag
    # the evaluation manager's tilt entry point: tilt + bending_tilt, tilt gradients summed
    kappa, c0 = (float(x) for x in g["param_helfrich_analytic"])
    gp.update(bending_modulus=kappa, spontaneous_curvature=c0, bending_energy_model="helfrich",
              bending_gradient_mode="analytic")
    mods = [tilt, bending_tilt, surface]
    ev = EvaluationManager(mesh=mesh, global_params=gp, param_resolver=res, energy_modules=mods,
                           energy_module_names=["tilt", "bending_tilt", "surface"])
    tg = np.full_like(pos, 7.0)
    e = ev.compute_energy_and_tilt_gradient_array(positions=pos, tilts=g["tilts"], tilt_grad_arr=tg)
    _close(e, float(g["E_tilt"]) + float(g["E_bending_tilt_helfrich_analytic"]))
    assert rel_err(tg, g["tg_tilt"] + g["tgonly_bending_tilt_helfrich_analytic"]) <= TOL
    _close(ev.compute_total_energy_array_with_tilts(positions=pos, tilts=g["tilts"]),
           float(g["E_tilt"]) + float(g["E_bending_tilt_helfrich_analytic"]) + float(g["E_surface"]))
    # tilt-dependent energy only (evaluation_manager.py:303-384): surface does not take part; at another tilt field
    # the tilt magnitude scales quadratically
    _close(ev.compute_energy_array_with_tilts(positions=pos, tilts=g["tilts"]),
           float(g["E_tilt"]) + float(g["E_bending_tilt_helfrich_analytic"]))
    ev_t = EvaluationManager(mesh=mesh, global_params=gp, param_resolver=res, energy_modules=[tilt, surface],
                             energy_module_names=["tilt", "surface"],
                             experimental_energy_scale_fn=lambda name: 0.5 if name == "tilt" else 1.0)
    _close(ev_t.compute_energy_array_with_tilts(positions=pos, tilts=3.0 * g["tilts"]), 0.5 * 9.0 * float(g["E_tilt"]))


def test_bending_finite_difference_mode_is_refused(backend):
    g = dict(np.load(golden_module_files()[0]))
    mesh, gp, res = _mesh(g, bending_gradient_mode="fd")
    gp["bending_modulus"] = 1.0
    with pytest.raises(L.B200Error):
        bending.compute_energy_and_gradient_array(mesh, gp, res, positions=mesh.positions_view(),
                                                  index_map=mesh.vertex_index_to_row,
                                                  grad_arr=np.zeros_like(mesh.positions_view()))


@pytest.mark.parametrize("case", ["cube", "bcube"])
def test_fused_evaluation_manager_vs_reference_minimizer(backend, case):
    """EvaluationManager twin: one fused device pass equals the reference Minimizer's
    compute_energy_and_gradient_array (module sum + KKT projection + fixed mask)."""
    z = np.load(H.GOLDEN + "/minimizer.npz")
    g = {k[len(case) + 1:]: z[k] for k in z.files if k.startswith(case + "_")}
    names = [str(n) for n in g["modules"]]
    gp = GlobalParams(surface_tension=1.0, volume_constraint_mode=str(g["mode"]), volume_stiffness=float(g["kvol"]),
                      bending_modulus=float(g["kappa"]), spontaneous_curvature=float(g["c0"]),
                      bending_energy_model=str(g["model"]), bending_gradient_mode="analytic")
    mesh = ArrayMesh(g["pos"], g["tri"], global_params=gp, facet_params={"surface_tension": g["gamma"]},
                     bodies={0: ArrayBody(g["body_rows_0"], target_volume=float(g["body_target_0"]))},
                     fixed=g["fixed"])
    mgr = EnergyModuleManager(names)
    with pytest.raises(ImportError):
        EnergyModuleManager(["line_tension"])
    with pytest.raises(KeyError):
        mgr.get_module("nope")
    ev = EvaluationManager(mesh=mesh, global_params=gp, param_resolver=ParamResolver(gp),
                           energy_modules=[mgr.get_module(n) for n in names], energy_module_names=names)
    pos = mesh.positions_view()
    if len(g["constraints"]):  # lagrange volume constraint: projected gradient on the device
        e, grad, res = ev.compute_energy_and_projected_gradient(positions=pos)
    else:
        e, grad = ev.compute_energy_and_gradient_array(positions=pos)
        grad[np.asarray(g["fixed"], bool)] = 0.0
    _close(e, float(g["E"]))
    assert rel_err(grad, g["g"]) <= 5e-12
    bd = ev.compute_energy_breakdown(positions=pos)
    assert set(bd) == set(names)
    for n in names:
        if f"E_{n}" in g:
            _close(bd[n], float(g[f"E_{n}"]), 1e-11)
    _close(ev.compute_energy_array_total(positions=pos), float(g["E"]), 1e-11)


def test_residency_follows_the_version_counters(backend):
    """Topology is packed once per topology version; positions travel per evaluation
    (SURVEY.md section 3.5)."""
    g = dict(np.load(golden_module_files()[0]))
    mesh, gp, res = _mesh(g)
    pos, idx = mesh.positions_view(), mesh.vertex_index_to_row
    for _ in range(3):
        surface.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=idx,
                                                  grad_arr=np.zeros_like(pos))
    st = mesh._b200_state
    assert st.uploads == 1
    e0 = surface.compute_energy_array(mesh, gp, positions=pos, index_map=idx)
    e1 = surface.compute_energy_array(mesh, gp, positions=1.1 * pos, index_map=idx)  # trial positions
    _close(e1, 1.21 * e0, 1e-12)
    assert st.uploads == 1
    mesh.bodies.clear()
    mesh.facet_params["surface_tension"] = g["gamma"][:-2]
    mesh.set_triangles(g["tri"][:-2])   # what refine / equiangulate do: topology version bump
    surface.compute_energy_array(mesh, gp, positions=pos, index_map=idx)
    assert st.uploads == 2 and st.nf == g["tri"].shape[0] - 2


def test_fixed_flags_follow_their_own_counter(backend):
    """Fixing / releasing vertices bumps only ``_fixed_flags_version`` (geometry/mesh.py:211-231): the device
    mask follows it without a topology upload, and the projected gradient zeroes exactly the new rows."""
    g = dict(np.load(H.GOLDEN + "/modules_cube_r2_jit.npz"))
    mesh, gp, res = _mesh(g)
    pos = mesh.positions_view()
    ev = EvaluationManager(mesh=mesh, global_params=gp, param_resolver=res, energy_modules=[surface],
                           energy_module_names=["surface"])
    mesh.set_fixed(np.zeros(len(pos), bool))
    _, g0, _ = ev.compute_energy_and_projected_gradient(positions=pos)
    st = mesh._b200_state
    assert st.uploads == 1 and np.all(np.abs(g0).sum(axis=1) > 0)
    fixed = np.zeros(len(pos), bool)
    fixed[[1, 4, 7]] = True
    mesh.set_fixed(fixed)
    _, g1, _ = ev.compute_energy_and_projected_gradient(positions=pos)
    assert st.uploads == 1 and st.fixed_uploads >= 1
    assert np.all(g1[fixed] == 0.0) and np.array_equal(g1[~fixed], g0[~fixed])
    fixed2 = np.zeros(len(pos), bool)
    fixed2[2] = True
    mesh.set_fixed(fixed2)
    _, g2, _ = ev.compute_energy_and_projected_gradient(positions=pos)
    assert st.uploads == 1 and np.all(g2[2] == 0.0) and np.array_equal(g2[1], g0[1])


def test_enforce_constraint_on_arrays(backend):
    """constraints/volume.py:69-149 on dense arrays: Newton projection onto V = V0 with
    the device's volume and dV/dx; the step is checked against the closed form of the oracle's volume gradient, the
    target is reached, fixed rows stay, and penalty mode is a no-op unless the projection is forced."""
    from oracle import ref_modules as ref

    g = dict(np.load(H.GOLDEN + "/modules_cube_r2_jit.npz"))
    mesh, gp, _ = _mesh(g)
    pos0 = 1.05 * np.array(mesh.positions_view())
    fixed = np.zeros(pos0.shape[0], bool)
    fixed[::13] = True
    mesh.set_fixed(fixed)
    mesh.set_positions(pos0)
    target = float(g["body_target_0"])
    # one iteration = one closed-form Newton step
    tri_body = np.asarray(g["tri"])[np.asarray(g["body_rows_0"])]
    vol, gv = ref.body_volume(pos0, tri_body), np.zeros_like(pos0)
    ref.accumulate_volume_gradient(pos0, tri_body, gv, 1.0)
    step = (vol - target) / (np.vdot(gv, gv) + 1e-12) * gv
    step[fixed] = 0.0
    volume_constraint.enforce_constraint(mesh, global_params=gp, max_iter=1)
    assert np.max(np.abs(np.array(mesh.positions_view()) - (pos0 - step))) <= 1e-13
    volume_constraint.enforce_constraint(mesh, global_params=gp, context="mesh_operation")
    pos = np.array(mesh.positions_view())
    v_end = ref.body_volume(pos, tri_body)
    assert abs(v_end - target) <= 1e-11
    assert np.array_equal(pos[fixed], pos0[fixed])
    # penalty mode: nothing moves unless forced
    gp.set("volume_constraint_mode", "penalty")
    mesh.set_positions(pos0)
    volume_constraint.enforce_constraint(mesh, global_params=gp)
    assert np.array_equal(np.array(mesh.positions_view()), pos0)
    volume_constraint.enforce_constraint(mesh, global_params=gp, force_projection=True)
    assert not np.array_equal(np.array(mesh.positions_view()), pos0)
    assert mesh._b200_state.uploads == 1 if hasattr(mesh, "_b200_state") else True


def test_array_refinement_matches_reference_hierarchy():
    """1 -> 4 refinement on arrays: counts of the 24 * 4^k hierarchy, closedness,
    orientation (volume stays +1), inherited fixed flags."""
    from membrane_solver_b200.geometry.refine import cube_mesh, refine_triangles

    pos, tri = cube_mesh()
    assert pos.shape == (14, 3) and tri.shape == (24, 3)
    fixed = np.zeros(14, bool)
    fixed[[0, 1]] = True
    for k in range(1, 4):
        pos, tri, fixed, parent = refine_triangles(pos, tri, fixed)
        assert tri.shape[0] == 24 * 4**k and pos.shape[0] == 12 * 4**k + 2
        assert parent.shape == (tri.shape[0],) and parent.max() == 24 * 4 ** (k - 1) - 1
        v = pos[tri]
        vol = np.einsum("ij,ij->i", np.cross(v[:, 1], v[:, 2]), v[:, 0]).sum() / 6.0
        area = 0.5 * np.linalg.norm(np.cross(v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]), axis=1).sum()
        assert abs(vol - 1.0) < 1e-14 and abs(area - 6.0) < 1e-13
        assert not ArrayMesh(pos, tri).boundary_vertex_ids          # closed
        assert fixed.sum() == 2 + (2**k - 1)                         # the refined edge 0-1 stays fixed


def test_array_vertex_averaging_matches_reference():
    """Mesh maintenance next to the refinement: ``geometry.vertex_average`` against the
    reference's ``runtime/vertex_average.py`` on a jittered cube (closed) and catenoid (fixed rims)."""
    import os

    from membrane_solver_b200.geometry.vertex_average import vertex_average_arrays
    from ms_test_helpers import GOLDEN

    g = np.load(os.path.join(GOLDEN, "vertex_average.npz"))
    for name in ("cube", "catenoid"):
        pos0, tri, fixed = g[name + "_pos0"], g[name + "_tri"], g[name + "_fixed"]
        out = vertex_average_arrays(pos0, tri, movable=~fixed)
        assert np.max(np.abs(out - g[name + "_pos1"])) <= 1e-13
        assert np.array_equal(out[fixed], pos0[fixed])
        assert np.max(np.abs(out - pos0)) > 1e-3
    # pin groups: a grouped vertex only averages over neighbours of its own group
    pos0, tri = g["cube_pos0"], g["cube_tri"]
    group = np.full(len(pos0), -1)
    group[:5] = 0
    out = vertex_average_arrays(pos0, tri, group=group)
    assert np.max(np.abs(out[5:] - g["cube_pos1"][5:])) <= 1e-13     # ungrouped vertices are unaffected
    assert not np.allclose(out[:5], g["cube_pos1"][:5])
    assert vertex_average_arrays(pos0, np.zeros((0, 3), np.int32)).tolist() == pos0.tolist()
