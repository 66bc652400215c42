"""Leaflet tilt modules (BASELINE config 4: bending_tilt_in/out, tilt_in/out).

Golden vectors: ``tests/golden/leaflet.npz`` -- outputs of the REAL reference on its caveolin free-disk mesh
(``tests/golden/generate_golden.py leaflet``).  CPU tier: the oracle restatement and the host emulator of
the device code against those vectors; GPU tier: the CUDA path through the C ABI.
"""

import os

import numpy as np
import pytest

from ms_test_helpers import GOLDEN, emulate_leaflet, rel_err

TOL = 1e-12
STATES = ("r0", "r1", "r1c")
LEAFLETS = ("in", "out")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "leaflet.npz"))


def _inputs(gold, state, leaf):
    pre = f"{state}_{leaf}_"
    return dict(pos=gold[f"{state}_pos"], tri=gold[f"{state}_tri"], is_boundary=gold[f"{state}_is_boundary"],
                tilts=gold[pre + "tilts"], keep=gold[pre + "keep"], interior=gold[pre + "interior"],
                base_zero=gold[pre + "base_zero"], kappa=gold[pre + "kappa"], c0=gold[pre + "c0"],
                sign=float(gold[pre + "sign"]), k_tilt=float(gold[pre + "k_tilt"]),
                k_smooth=float(gold[pre + "k_smooth"]), consistent=bool(gold[pre + "consistent"]))


def _close(a, b):
    assert abs(a - b) <= TOL * max(1.0, abs(b)), (a, b)


@pytest.mark.parametrize("leaf", LEAFLETS)
@pytest.mark.parametrize("state", STATES)
def test_oracle_matches_reference(gold, state, leaf):
    from oracle import ref_leaflet as rl

    i = _inputs(gold, state, leaf)
    pre = f"{state}_{leaf}_"
    g, tg = np.zeros_like(i["pos"]), np.zeros_like(i["pos"])
    e = rl.leaflet_bending_tilt_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["kappa"], i["c0"], sign=i["sign"],
                                                    keep=i["keep"], interior=i["interior"], base_zero=i["base_zero"],
                                                    is_boundary=i["is_boundary"], grad=g, tilt_grad=tg)
    _close(e, float(gold[pre + "E_bt"]))
    assert rel_err(g, gold[pre + "g_bt"]) <= TOL and rel_err(tg, gold[pre + "tg_bt"]) <= TOL
    tg = np.zeros_like(i["pos"])
    e = rl.leaflet_bending_tilt_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["kappa"], i["c0"], sign=i["sign"],
                                                    keep=i["keep"], interior=i["interior"], base_zero=i["base_zero"],
                                                    is_boundary=i["is_boundary"], grad=None, tilt_grad=tg)
    _close(e, float(gold[pre + "E_bt_tiltonly"]))
    assert rel_err(tg, gold[pre + "tg_bt_tiltonly"]) <= TOL
    g, tg = np.zeros_like(i["pos"]), np.zeros_like(i["pos"])
    e = rl.leaflet_tilt_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["k_tilt"], keep=i["keep"],
                                            consistent=i["consistent"], grad=g, tilt_grad=tg)
    _close(e, float(gold[pre + "E_tilt"]))
    assert rel_err(g, gold[pre + "g_tilt"]) <= TOL and rel_err(tg, gold[pre + "tg_tilt"]) <= TOL
    tg = np.zeros_like(i["pos"])
    e = rl.leaflet_tilt_smoothness_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["k_smooth"], keep=i["keep"],
                                                       tilt_grad=tg)
    _close(e, float(gold[pre + "E_smooth"]))
    assert rel_err(tg, gold[pre + "tg_smooth"]) <= TOL
    assert not np.any(gold[pre + "g_smooth"])          # the reference's smoothness term has no shape gradient


@pytest.mark.parametrize("leaf", LEAFLETS)
@pytest.mark.parametrize("state", STATES)
def test_emulated_device_code_matches_reference(gold, state, leaf):
    i = _inputs(gold, state, leaf)
    pre = f"{state}_{leaf}_"
    common = dict(sign=i["sign"], keep=i["keep"], is_boundary=i["is_boundary"], interior=i["interior"],
                  base_zero=i["base_zero"], kappa=i["kappa"], c0=i["c0"], consistent_u=i["consistent"],
                  k_tilt=i["k_tilt"], k_smooth=i["k_smooth"])
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], with_bt=False, with_smooth=True, **common)
    _close(r["E_smooth"], float(gold[pre + "E_smooth"]))
    assert rel_err(r["tilt_grad"], gold[pre + "tg_smooth"]) <= TOL and not np.any(r["grad"])
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], with_bt=True, with_tilt=False, **common)
    _close(r["E_bt"], float(gold[pre + "E_bt"]))
    assert rel_err(r["grad"], gold[pre + "g_bt"]) <= TOL
    assert rel_err(r["tilt_grad"], gold[pre + "tg_bt"]) <= TOL
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], with_bt=False, with_tilt=True, **common)
    _close(r["E_tilt"], float(gold[pre + "E_tilt"]))
    assert rel_err(r["grad"], gold[pre + "g_tilt"]) <= TOL
    assert rel_err(r["tilt_grad"], gold[pre + "tg_tilt"]) <= TOL
    # both modules of the leaflet in one sweep = the sum of the two
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], with_bt=True, with_tilt=True, **common)
    assert rel_err(r["grad"], gold[pre + "g_bt"] + gold[pre + "g_tilt"]) <= TOL
    assert rel_err(r["tilt_grad"], gold[pre + "tg_bt"] + gold[pre + "tg_tilt"]) <= TOL
    # tilt-only evaluation (inner relaxation loop)
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], with_bt=True, with_tilt=True, want_grad=False, **common)
    _close(r["E_bt"], float(gold[pre + "E_bt_tiltonly"]))
    assert rel_err(r["tilt_grad"], gold[pre + "tg_bt_tiltonly"] + gold[pre + "tg_tilt_tiltonly"]) <= TOL


def test_emulator_matches_oracle_with_row_weights_and_mixed_mass_modes(gold):
    """Options the caveolin fixture does not exercise: active-row weights and a per-facet mass mode."""
    from oracle import ref_leaflet as rl

    i = _inputs(gold, "r1", "out")
    rng = np.random.default_rng(5)
    nv, nf = i["pos"].shape[0], i["tri"].shape[0]
    w = rng.choice([0.0, 0.5, 1.0], size=nv)
    cons = rng.random(nf) < 0.5
    g, tg = np.zeros_like(i["pos"]), np.zeros_like(i["pos"])
    e = rl.leaflet_tilt_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["k_tilt"], keep=i["keep"],
                                            row_weights=w, consistent=cons, grad=g, tilt_grad=tg)
    r = emulate_leaflet(i["pos"], i["tri"], i["tilts"], sign=1.0, keep=i["keep"], row_weight=w, consistent=cons,
                        k_tilt=i["k_tilt"], with_bt=False, with_tilt=True)
    _close(r["E_tilt"], e)
    assert rel_err(r["grad"], g) <= TOL and rel_err(r["tilt_grad"], tg) <= TOL


# ------------------------------------------------------------------ GPU tier (C ABI)
def _device_leaflet(i, leaf, *, order_hint=None, **over):
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh

    dm = DeviceMesh(0)
    dm.set_topology(i["pos"].shape[0], i["tri"], is_boundary=i["is_boundary"].astype(np.uint8), order_hint=order_hint)
    which = L.LEAFLET_IN if leaf == "in" else L.LEAFLET_OUT
    args = dict(div_sign=i["sign"], kappa=i["kappa"], c0=i["c0"], k_tilt=i["k_tilt"], k_smooth=i["k_smooth"], facet_keep=i["keep"],
                interior=i["interior"], base_zero=i["base_zero"], consistent=i["consistent"])
    args.update(over)
    dm.set_leaflet(which, **args)
    dm.set_positions(i["pos"])
    dm.upload(L.ARR_TILTS_IN if leaf == "in" else L.ARR_TILTS_OUT, i["tilts"])
    return dm, which, (L.ARR_TILT_GRAD_IN if leaf == "in" else L.ARR_TILT_GRAD_OUT)


@pytest.mark.gpu
@pytest.mark.parametrize("hint", [False, True], ids=["plain", "reordered"])
@pytest.mark.parametrize("leaf", LEAFLETS)
@pytest.mark.parametrize("state", STATES)
def test_device_matches_reference(gold, state, leaf, hint):
    from membrane_solver_b200 import _lib as L

    i = _inputs(gold, state, leaf)
    pre = f"{state}_{leaf}_"
    dm, which, arr_tg = _device_leaflet(i, leaf, order_hint=i["pos"] if hint else None)
    _, _, e_s = dm.eval_leaflet(which, L.MOD_TILT_SMOOTHNESS)
    _close(e_s, float(gold[pre + "E_smooth"]))
    assert rel_err(dm.download(arr_tg), gold[pre + "tg_smooth"]) <= TOL and not np.any(dm.download(L.ARR_GRAD))
    e_bt, e_t, _ = dm.eval_leaflet(which, L.MOD_BENDING_TILT)
    _close(e_bt, float(gold[pre + "E_bt"]))
    assert e_t == 0.0
    assert rel_err(dm.download(L.ARR_GRAD), gold[pre + "g_bt"]) <= TOL
    assert rel_err(dm.download(arr_tg), gold[pre + "tg_bt"]) <= TOL
    e_bt, e_t, _ = dm.eval_leaflet(which, L.MOD_TILT)
    _close(e_t, float(gold[pre + "E_tilt"]))
    assert rel_err(dm.download(L.ARR_GRAD), gold[pre + "g_tilt"]) <= TOL
    assert rel_err(dm.download(arr_tg), gold[pre + "tg_tilt"]) <= TOL
    # both modules in one sweep, accumulated on top of the previous results
    g_prev, tg_prev = dm.download(L.ARR_GRAD), dm.download(arr_tg)
    e_bt, e_t, _ = dm.eval_leaflet(which, L.MOD_TILT | L.MOD_BENDING_TILT, accumulate=L.ACC_GRAD | L.ACC_TILT_GRAD)
    _close(e_bt, float(gold[pre + "E_bt"]))
    _close(e_t, float(gold[pre + "E_tilt"]))
    assert rel_err(dm.download(L.ARR_GRAD) - g_prev, gold[pre + "g_bt"] + gold[pre + "g_tilt"]) <= 4 * TOL
    assert rel_err(dm.download(arr_tg) - tg_prev, gold[pre + "tg_bt"] + gold[pre + "tg_tilt"]) <= 4 * TOL
    # tilt-only evaluation: the shape gradient array is left alone
    g_before = dm.download(L.ARR_GRAD)
    e_bt, e_t, e_s = dm.eval_leaflet(which, L.MOD_TILT | L.MOD_BENDING_TILT | L.MOD_TILT_SMOOTHNESS, want_grad=False)
    _close(e_bt, float(gold[pre + "E_bt_tiltonly"]))
    _close(e_s, float(gold[pre + "E_smooth_tiltonly"]))
    assert rel_err(dm.download(arr_tg), gold[pre + "tg_bt_tiltonly"] + gold[pre + "tg_tilt_tiltonly"]
                   + gold[pre + "tg_smooth_tiltonly"]) <= TOL
    assert np.array_equal(dm.download(L.ARR_GRAD), g_before)
    # repeatable bit for bit (fixed-order gathers, no atomics)
    dm.eval_leaflet(which, L.MOD_TILT | L.MOD_BENDING_TILT)
    g1, t1 = dm.download(L.ARR_GRAD), dm.download(arr_tg)
    dm.eval_leaflet(which, L.MOD_TILT | L.MOD_BENDING_TILT)
    assert np.array_equal(dm.download(L.ARR_GRAD), g1) and np.array_equal(dm.download(arr_tg), t1)
    dm.close()


@pytest.mark.gpu
def test_device_row_weights_mixed_mass_modes_and_errors(gold):
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh
    from oracle import ref_leaflet as rl

    i = _inputs(gold, "r1", "out")
    rng = np.random.default_rng(5)
    nv, nf = i["pos"].shape[0], i["tri"].shape[0]
    w = rng.choice([0.0, 0.5, 1.0], size=nv)
    cons = rng.random(nf) < 0.5
    g, tg = np.zeros_like(i["pos"]), np.zeros_like(i["pos"])
    e = rl.leaflet_tilt_energy_and_gradient(i["pos"], i["tri"], i["tilts"], i["k_tilt"], keep=i["keep"],
                                            row_weights=w, consistent=cons, grad=g, tilt_grad=tg)
    dm, which, arr_tg = _device_leaflet(i, "out", tilt_row_weight=w, facet_consistent=cons)
    _, e_t, _ = dm.eval_leaflet(which, L.MOD_TILT)
    _close(e_t, e)
    assert rel_err(dm.download(L.ARR_GRAD), g) <= TOL and rel_err(dm.download(arr_tg), tg) <= TOL
    # uniform parameters as scalars = the same as arrays
    dm.set_leaflet(which, div_sign=1.0, kappa=1.0, c0=0.0, k_tilt=i["k_tilt"], facet_keep=i["keep"],
                   interior=i["interior"], base_zero=i["base_zero"])
    e_bt, _, _ = dm.eval_leaflet(which, L.MOD_BENDING_TILT)
    _close(e_bt, float(gold["r1_out_E_bt"]))
    dm.close()
    # loud failures: leaflet not configured, tilt field missing, foreign module bits
    dm = DeviceMesh(0)
    dm.set_topology(nv, i["tri"])
    dm.set_positions(i["pos"])
    with pytest.raises(L.B200Error):
        dm.eval_leaflet(L.LEAFLET_IN, L.MOD_TILT)
    dm.set_leaflet(L.LEAFLET_IN, div_sign=-1.0, kappa=1.0, k_tilt=1.0)
    with pytest.raises(L.B200Error):
        dm.eval_leaflet(L.LEAFLET_IN, L.MOD_TILT)
    dm.upload(L.ARR_TILTS_IN, i["tilts"])
    with pytest.raises(L.B200Error):
        dm.eval_leaflet(L.LEAFLET_IN, L.MOD_SURFACE)
    dm.eval_leaflet(L.LEAFLET_IN, L.MOD_TILT)
    dm.close()


# ------------------------------------------------------------------ plugin level
def _array_mesh(gold, state):
    from membrane_solver_b200.geometry.array_mesh import ArrayMesh

    leaflets, tilts = {}, {}
    for leaf in LEAFLETS:
        i = _inputs(gold, state, leaf)
        leaflets[leaf] = dict(keep_bt=i["keep"], keep_tilt=i["keep"], interior=i["interior"], base_zero=i["base_zero"],
                              kappa=i["kappa"], c0=i["c0"], k_tilt=i["k_tilt"], k_smooth=i["k_smooth"],
                              consistent=i["consistent"])
        tilts[leaf] = i["tilts"]
    return ArrayMesh(gold[f"{state}_pos"], gold[f"{state}_tri"], tilts_in=tilts["in"], tilts_out=tilts["out"],
                     leaflets=leaflets)


def _options_mesh(gold, state):
    """ArrayMesh carrying the mesh OPTIONS (vertex presets / group tags, global parameters) instead of ready-made
    masks: the selections are derived by membrane_solver_b200/modules/energy/leaflet_selection.py."""
    import json

    from membrane_solver_b200.geometry.array_mesh import ArrayMesh, GlobalParams

    vopts = {int(k): v for k, v in json.loads(str(gold[f"{state}_vertex_options_json"])).items()}
    gp = GlobalParams(json.loads(str(gold[f"{state}_global_params_json"])))
    return ArrayMesh(gold[f"{state}_mesh_pos"], gold[f"{state}_tri"], global_params=gp, vertex_options=vopts,
                     tilts_in=gold[f"{state}_in_tilts"], tilts_out=gold[f"{state}_out_tilts"]), gp


@pytest.mark.parametrize("state", STATES)
def test_own_leaflet_selection_matches_reference_masks(gold, state):
    """leaflet_presence.py:34-170, bt_selection.py:140-330, bt_params.py:40-318, tilt_params.py:6-24 re-derived on
    arrays: every mask / parameter equals what the reference's helpers produced for its caveolin mesh."""
    from membrane_solver_b200.geometry.array_mesh import ParamResolver
    from membrane_solver_b200.modules.energy import _leaflet as LF
    from membrane_solver_b200.modules.energy import leaflet_selection as LS

    mesh, gp = _options_mesh(gold, state)
    assert not LS.needs_reference_helpers(gp)
    assert set(mesh.boundary_vertex_ids) == set(np.nonzero(gold[f"{state}_is_boundary"])[0].tolist())
    for leaf in LEAFLETS:
        spec = LF.selection(mesh, gp, ParamResolver(gp), leaf)
        pre = f"{state}_{leaf}_"
        assert np.array_equal(spec["keep_bt"], gold[pre + "keep"]) and np.array_equal(spec["keep_tilt"], gold[pre + "keep"])
        assert np.array_equal(spec["interior"], gold[pre + "interior"])
        assert np.array_equal(spec["base_zero"], gold[pre + "base_zero"])
        assert np.array_equal(spec["kappa"], gold[pre + "kappa"]) and np.array_equal(spec["c0"], gold[pre + "c0"])
        assert spec["k_tilt"] == float(gold[pre + "k_tilt"]) and spec["k_smooth"] == float(gold[pre + "k_smooth"])
        assert spec["consistent"] == bool(gold[pre + "consistent"])
        assert LF.selection(mesh, gp, ParamResolver(gp), leaf) is spec          # cached on the version counters
    # the options that matter are really in play on this mesh (config 4): an absent outer leaflet on the disk,
    # a tagged rim ring without base term, assume-J0 rows on the inner leaflet
    assert not gold[f"{state}_out_keep"].all() and gold[f"{state}_in_keep"].all()
    assert gold[f"{state}_in_base_zero"].any()
    # variations the golden mesh does not use: radius clipping of the assume-J0 rows, a region mode, overrides
    gp2 = type(gp)(dict(gp, bending_tilt_assume_J0_presets_radius_max=3.0,
                        bending_tilt_base_term_region_mode="physical_disk_split_v1",
                        bending_tilt_base_term_region_radius=5.0))
    pos = mesh.positions_view()
    r = np.linalg.norm(pos[:, :2], axis=1)
    bz_in = LS.base_zero_mask(mesh, gp2, "in", pos)
    assert np.array_equal(bz_in, gold[f"{state}_in_base_zero"] & ~(r > 3.0 + 1e-12))
    bz_out = LS.base_zero_mask(mesh, gp2, "out", pos)
    assert np.array_equal(bz_out, r <= 5.0 + 1e-12)
    with pytest.raises(ValueError):
        LS.base_zero_mask(mesh, type(gp)({"bending_tilt_base_term_region_mode": "physical_disk_split_v1"}), "out", pos)
    assert LS.needs_reference_helpers(type(gp)({"rim_slope_match_mode": "shared_rim_staggered_v1"}))


def test_per_vertex_leaflet_parameter_overrides():
    """bt_params.py:233-318: leaflet key beats the generic key; c0 falls back from the leaflet key to
    spontaneous_curvature to intrinsic_curvature; values that are not numbers are skipped."""
    from membrane_solver_b200.geometry.array_mesh import ArrayMesh, GlobalParams
    from membrane_solver_b200.modules.energy import leaflet_selection as LS

    pos = np.zeros((5, 3))
    tri = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 4]], np.int32)
    vopts = {0: {"bending_modulus": 2.0, "bending_modulus_in": 7.0, "spontaneous_curvature": 0.3},
             1: {"bending_modulus": 3.0, "intrinsic_curvature": 0.4},
             2: {"spontaneous_curvature_out": -0.2, "spontaneous_curvature": 0.9, "bending_modulus_out": "soft"},
             3: {"preset": "disk"}}
    gp = GlobalParams({"bending_modulus": 1.0, "bending_modulus_out": 1.5, "spontaneous_curvature": 0.1,
                       "spontaneous_curvature_in": 0.05})
    mesh = ArrayMesh(pos, tri, global_params=gp, vertex_options=vopts)
    k_in, c_in = LS.per_vertex_params(mesh, gp, "in")
    k_out, c_out = LS.per_vertex_params(mesh, gp, "out")
    assert k_in.tolist() == [7.0, 3.0, 1.0, 1.0, 1.0] and k_out.tolist() == [2.0, 3.0, 1.5, 1.5, 1.5]
    assert c_in.tolist() == [0.3, 0.4, 0.9, 0.05, 0.05] and c_out.tolist() == [0.3, 0.4, -0.2, 0.1, 0.1]
    gp_abs = GlobalParams({"leaflet_out_absent_presets": ["disk", " "], "leaflet_in_absent_presets": None})
    assert LS.absent_vertex_mask(mesh, gp_abs, "out").tolist() == [False, False, False, True, False]
    assert not LS.absent_vertex_mask(mesh, gp_abs, "in").any()
    assert LS.present_triangle_mask(tri, LS.absent_vertex_mask(mesh, gp_abs, "out")).tolist() == [True, False, False]


def test_plugins_derive_their_selections_from_mesh_options(gold, monkeypatch):
    """End to end on the emulated device: the leaflet plugins on a mesh that carries only OPTIONS reproduce the
    reference's energies and gradients (the selections come from leaflet_selection.py, not from given masks)."""
    import importlib

    from fake_device import FakeDeviceMesh

    from membrane_solver_b200.geometry.array_mesh import ParamResolver
    from membrane_solver_b200.runtime import device_state

    monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    state = "r1"
    mesh, gp = _options_mesh(gold, state)
    pos = gold[f"{state}_pos"]                  # evaluation positions (jittered); the options use the mesh's own
    res = ParamResolver(gp)
    for leaf in LEAFLETS:
        for name, tag in ((f"bending_tilt_{leaf}", "bt"), (f"tilt_{leaf}", "tilt"), (f"tilt_smoothness_{leaf}", "smooth")):
            mod = importlib.import_module(f"membrane_solver_b200.modules.energy.{name}")
            g = np.zeros_like(pos)
            tg = np.zeros_like(pos)
            e = mod.compute_energy_and_gradient_array(mesh, gp, res, positions=pos, index_map=mesh.vertex_index_to_row,
                                                      grad_arr=g, **{f"tilt_{leaf}_grad_arr": tg})
            pre = f"{state}_{leaf}_"
            _close(e, float(gold[pre + f"E_{tag}"]))
            assert rel_err(g, gold[pre + f"g_{tag}"]) <= 4 * TOL and rel_err(tg, gold[pre + f"tg_{tag}"]) <= 4 * TOL


def _plugin_checks(gold, state):
    import importlib

    mesh = _array_mesh(gold, state)
    pos = mesh.positions_view()
    idx = mesh.vertex_index_to_row
    for leaf in LEAFLETS:
        for name, tag in ((f"bending_tilt_{leaf}", "bt"), (f"tilt_{leaf}", "tilt"), (f"tilt_smoothness_{leaf}", "smooth")):
            mod = importlib.import_module(f"membrane_solver_b200.modules.energy.{name}")
            assert mod.USES_TILT_LEAFLETS
            pre = f"{state}_{leaf}_"
            g = np.full_like(pos, 0.25)                       # plugins ADD into caller-owned arrays
            tg = {"in": np.full_like(pos, -0.5), "out": np.full_like(pos, 0.75)}
            e = mod.compute_energy_and_gradient_array(mesh, mesh.global_params, None, positions=pos, index_map=idx,
                                                      grad_arr=g, tilts_in=mesh.tilts_in_view(),
                                                      tilts_out=mesh.tilts_out_view(), tilt_in_grad_arr=tg["in"],
                                                      tilt_out_grad_arr=tg["out"])
            _close(e, float(gold[pre + f"E_{tag}"]))
            assert rel_err(g - 0.25, gold[pre + f"g_{tag}"]) <= 4 * TOL
            other = "out" if leaf == "in" else "in"
            assert rel_err(tg[leaf] - (-0.5 if leaf == "in" else 0.75), gold[pre + f"tg_{tag}"]) <= 4 * TOL
            assert np.all(tg[other] == (-0.5 if other == "in" else 0.75))     # the other leaflet is untouched
            # tilt-only evaluation and the energy-only entry point (tilts default to the mesh's own)
            t2 = np.zeros_like(pos)
            kw = {f"tilt_{leaf}_grad_arr": t2}
            e2 = mod.compute_energy_and_gradient_array(mesh, mesh.global_params, None, positions=pos, index_map=idx,
                                                       grad_arr=None, **kw)
            _close(e2, float(gold[pre + f"E_{tag}_tiltonly"]))
            assert rel_err(t2, gold[pre + f"tg_{tag}_tiltonly"]) <= TOL
            _close(mod.compute_energy_array(mesh, mesh.global_params, None, positions=pos, index_map=idx),
                   float(gold[pre + f"E_{tag}"]))
            ed, gd, tgd = mod.compute_energy_and_gradient(mesh, mesh.global_params, None)
            _close(ed, float(gold[pre + f"E_{tag}"]))
            assert rel_err(np.array([gd[v] for v in range(len(pos))]), gold[pre + f"g_{tag}"]) <= TOL
    st = mesh._b200_state
    assert st.uploads == 1                                    # one topology upload for all of the above
    with pytest.raises(ValueError):
        importlib.import_module("membrane_solver_b200.modules.energy.tilt_in").compute_energy_and_gradient_array(
            mesh, mesh.global_params, None, positions=pos, index_map=idx, grad_arr=None, tilts_in=np.zeros((3, 3)))
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.geometry.array_mesh import GlobalParams

    for key, val in (("theory_parity_lane", "stage_a_emergent"), ("tilt_transport_model", "connection_v1"),
                     ("bending_gradient_mode", "approx"), ("bending_tilt_in_update_mode", "radial_cross_term_off_v1")):
        gp = GlobalParams({key: val})
        with pytest.raises(L.B200Error):
            importlib.import_module("membrane_solver_b200.modules.energy.bending_tilt_in").compute_energy_array(
                mesh, gp, None, positions=pos, index_map=idx)


@pytest.mark.parametrize("state", ["r0", "r1c"])
def test_plugins_on_emulated_device(gold, state, monkeypatch):
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200.runtime import device_state

    monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    _plugin_checks(gold, state)


@pytest.mark.gpu
@pytest.mark.parametrize("state", ["r0", "r1", "r1c"])
def test_plugins_on_device(gold, state):
    _plugin_checks(gold, state)


def _manager_checks(gold, state):
    """The B200 EvaluationManager's leaflet entry points: all four twins in one device pass."""
    import importlib

    from membrane_solver_b200.runtime.evaluation_manager import EvaluationManager

    mesh = _array_mesh(gold, state)
    names = ["bending_tilt_in", "bending_tilt_out", "tilt_in", "tilt_out"]
    mods = [importlib.import_module(f"membrane_solver_b200.modules.energy.{n}") for n in names]
    ev = EvaluationManager(mesh=mesh, global_params=mesh.global_params, param_resolver=None, energy_modules=mods,
                           energy_module_names=names)
    pos = mesh.positions_view()
    ti, to = mesh.tilts_in_view(), mesh.tilts_out_view()
    gi, go = np.ones_like(pos), np.ones_like(pos)            # overwritten, not accumulated (":651-652")
    e = ev.compute_energy_and_leaflet_tilt_gradients_array(positions=pos, tilts_in=ti, tilts_out=to,
                                                           tilt_in_grad_arr=gi, tilt_out_grad_arr=go, tilt_only=True)
    want_e = sum(float(gold[f"{state}_{leaf}_E_{tag}"]) for leaf in LEAFLETS for tag in ("bt", "tilt"))
    _close(e, want_e)
    for leaf, got in (("in", gi), ("out", go)):
        want = gold[f"{state}_{leaf}_tg_bt_tiltonly"] + gold[f"{state}_{leaf}_tg_tilt_tiltonly"]
        assert rel_err(got, want) <= TOL
    _close(ev.compute_tilt_dependent_energy_with_leaflet_tilts(positions=pos, tilts_in=ti, tilts_out=to), want_e)
    _close(ev.compute_energy_array_with_leaflet_tilts(positions=pos, tilts_in=ti, tilts_out=to), want_e)
    # a per-module experimental scale splits the sweep and scales energy and gradient alike
    ev2 = EvaluationManager(mesh=mesh, global_params=mesh.global_params, param_resolver=None, energy_modules=mods,
                            energy_module_names=names,
                            experimental_energy_scale_fn=lambda n: 0.5 if n == "tilt_in" else 1.0)
    gi2, go2 = np.zeros_like(pos), np.zeros_like(pos)
    e2 = ev2.compute_energy_and_leaflet_tilt_gradients_array(positions=pos, tilts_in=ti, tilts_out=to,
                                                             tilt_in_grad_arr=gi2, tilt_out_grad_arr=go2)
    _close(e2, want_e - 0.5 * float(gold[f"{state}_in_E_tilt"]))
    assert rel_err(gi2, gold[f"{state}_in_tg_bt"] + 0.5 * gold[f"{state}_in_tg_tilt"]) <= TOL
    assert rel_err(go2, go) <= TOL
    return mesh


def test_manager_leaflet_entry_points_on_emulated_device(gold, monkeypatch):
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200.runtime import device_state

    monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    mesh = _manager_checks(gold, "r1")
    fake = mesh._b200_state.dm
    assert fake.topology_uploads == 1


@pytest.mark.gpu
@pytest.mark.parametrize("state", ["r1", "r1c"])
def test_manager_leaflet_entry_points_on_device(gold, state):
    _manager_checks(gold, state)


# ------------------------------------------------------------------ P1 operators (row a17)
def test_oracle_p1_vertex_divergence_matches_reference():
    from oracle import ref_modules as rm

    g = np.load(os.path.join(GOLDEN, "p1_vertex.npz"))
    div_v, area_v = rm.p1_vertex_divergence(g["pos"], g["tilts"], g["tri"])
    assert rel_err(div_v, g["div_v"]) <= TOL and rel_err(area_v, g["area_v"]) <= TOL


@pytest.mark.gpu
def test_device_p1_operators_match_reference():
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.geometry import tilt_operators as ops
    from oracle import ref_modules as rm

    g = np.load(os.path.join(GOLDEN, "p1_vertex.npz"))
    pos, tri, tilts = g["pos"], g["tri"], g["tilts"]
    div_v, area_v = ops.p1_vertex_divergence(n_vertices=len(pos), positions=pos, tilts=tilts, tri_rows=tri)
    assert rel_err(div_v, g["div_v"]) <= TOL and rel_err(area_v, g["area_v"]) <= TOL
    div, area, g0, g1, g2 = ops.p1_triangle_divergence(positions=pos, tilts=tilts, tri_rows=tri)
    want = rm.p1_triangle_divergence(pos, tilts, tri)
    for a, b in zip((div, area, g0, g1, g2), want):
        assert rel_err(a, b) <= TOL
    area2, h0, _, _ = ops.p1_triangle_shape_gradients(positions=pos, tri_rows=tri)
    assert np.array_equal(area2, area) and np.array_equal(h0, g0)
    assert ops.p1_vertex_divergence(n_vertices=0, positions=pos, tilts=tilts, tri_rows=tri)[0].size == 0
    e_div, e_area = ops.p1_vertex_divergence(n_vertices=len(pos), positions=pos, tilts=tilts,
                                             tri_rows=np.zeros((0, 3), np.int32))
    assert not e_div.any() and not e_area.any()
    with pytest.raises(L.B200Error):
        ops.p1_triangle_divergence(positions=pos, tilts=tilts, tri_rows=tri, transport_model="connection_v1")


# ------------------------------------------------------------------ leaflet tilt relaxation (row f3)
RELAX_CASES = ("gd5", "gd4small", "gdreject", "cg6", "cg5plain", "cg6big")


@pytest.fixture(scope="module")
def relax_gold():
    return np.load(os.path.join(GOLDEN, "tilt_relaxation.npz"))


def _relax_leaflets(g, case):
    p = case + "_"
    return {k: dict(keep=g[p + f"{k}_keep"], interior=g[p + f"{k}_interior"], base_zero=g[p + f"{k}_base_zero"],
                    kappa=g[p + f"{k}_kappa"], c0=g[p + f"{k}_c0"], k_tilt=float(g[p + f"{k}_k_tilt"]))
            for k in LEAFLETS}


def _relax_stats_match(st, g, case):
    p = case + "_"
    assert st["accepted_steps"] == int(g[p + "accepted_steps"])
    assert st["backtracking_steps"] == int(g[p + "backtracking_steps"])
    for k in ("initial_energy", "final_energy", "initial_gradient_norm", "final_gradient_norm"):
        assert abs(st[k] - float(g[p + k])) <= 1e-11 * max(1.0, abs(float(g[p + k]))), k


@pytest.mark.parametrize("case", RELAX_CASES)
def test_oracle_tilt_relaxation_matches_reference(relax_gold, case):
    from oracle import ref_leaflet as rl

    g, p = relax_gold, case + "_"
    common = dict(is_boundary=g[p + "is_boundary"], fixed={"in": g[p + "fixed_in"], "out": g[p + "fixed_out"]},
                  max_iters=int(g[p + "steps"]), step_size=float(g[p + "step_size"]))
    start = {"in": g[p + "tilts_in0"], "out": g[p + "tilts_out0"]}
    if str(g[p + "solver"]) == "cg":
        t, st = rl.relax_leaflet_tilts_cg(g[p + "pos"], g[p + "tri"], start, _relax_leaflets(g, case),
                                          preconditioner=bool(g[p + "preconditioner"]),
                                          gd_fallback=bool(g[p + "gd_fallback"]),
                                          k_smooth={"in": float(g[p + "k_smooth_in"]), "out": float(g[p + "k_smooth_out"])},
                                          **common)
    else:
        t, st = rl.relax_leaflet_tilts_gd(g[p + "pos"], g[p + "tri"], start, _relax_leaflets(g, case), **common)
    _relax_stats_match(st, g, case)
    assert np.max(np.abs(t["in"] - g[p + "tilts_in1"])) <= 1e-12
    assert np.max(np.abs(t["out"] - g[p + "tilts_out1"])) <= 1e-12


def _device_relaxation(g, case, factory):
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.runtime.device_tilt_relaxer import DeviceTiltRelaxer

    p = case + "_"
    pos, tri = g[p + "pos"], g[p + "tri"]
    dm = factory(0)
    dm.set_topology(pos.shape[0], tri, is_boundary=g[p + "is_boundary"].astype(np.uint8))
    dm.set_positions(pos)
    for leaf, d in _relax_leaflets(g, case).items():
        which = L.LEAFLET_IN if leaf == "in" else L.LEAFLET_OUT
        dm.set_leaflet(which, div_sign=-1.0 if leaf == "in" else 1.0, kappa=d["kappa"], c0=d["c0"], k_tilt=d["k_tilt"],
                       facet_keep=d["keep"].astype(np.uint8), interior=d["interior"].astype(np.uint8),
                       base_zero=d["base_zero"].astype(np.uint8))
        dm.set_leaflet_fixed(which, g[p + f"fixed_{leaf}"].astype(np.uint8))
        dm.upload(L.ARR_TILTS_IN if leaf == "in" else L.ARR_TILTS_OUT, g[p + f"tilts_{leaf}0"])
    st = DeviceTiltRelaxer(dm).relax(
        max_iters=int(g[p + "steps"]), step_size=float(g[p + "step_size"]), solver=str(g[p + "solver"]),
        preconditioner=bool(g[p + "preconditioner"]), gd_fallback=bool(g[p + "gd_fallback"]),
        k_smooth={"in": float(g[p + "k_smooth_in"]), "out": float(g[p + "k_smooth_out"])},
        area_kept_only={"out": not bool(np.all(g[p + "out_keep"]))})
    _relax_stats_match(st, g, case)
    assert st["stop_reason"] == str(g[p + "stop_reason"])
    # north_star: trajectories within 1e-9
    assert np.max(np.abs(dm.download(L.ARR_TILTS_IN) - g[p + "tilts_in1"])) <= 1e-10
    assert np.max(np.abs(dm.download(L.ARR_TILTS_OUT) - g[p + "tilts_out1"])) <= 1e-10
    fixed = g[p + "fixed_in"]
    if case != "gdreject":
        moved = np.abs(dm.download(L.ARR_TILTS_IN) - g[p + "tilts_in0"]).max(axis=1)
        assert moved[~fixed].max() > 1e-4
    dm.close()


@pytest.mark.parametrize("case", RELAX_CASES)
def test_tilt_relaxer_on_emulated_device(relax_gold, case):
    from fake_device import FakeDeviceMesh

    _device_relaxation(relax_gold, case, FakeDeviceMesh)


@pytest.mark.gpu
@pytest.mark.parametrize("case", RELAX_CASES)
def test_tilt_relaxer_on_device(relax_gold, case):
    from membrane_solver_b200.context import DeviceMesh

    _device_relaxation(relax_gold, case, DeviceMesh)


@pytest.mark.gpu
def test_device_leaflet_on_a_mesh_above_the_single_launch_threshold():
    """Meshes up to 65 536 facets are evaluated in one cooperative launch, larger ones by separate sweeps with
    two-stage energy sums: the larger path against the oracle (72 000 facets, random leaflet selections)."""
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh
    from membrane_solver_b200.synthetic import icosphere
    from oracle import ref_leaflet as rl

    pos, tri = icosphere(60)
    nv, nf = pos.shape[0], tri.shape[0]
    assert nf > 65536
    rng = np.random.default_rng(21)
    pos = pos * (1.0 + 0.02 * rng.standard_normal((nv, 1)))
    tilts = 0.1 * rng.standard_normal((nv, 3))
    keep = rng.random(nf) > 0.1
    interior = rng.random(nv) > 0.05
    base_zero = rng.random(nv) < 0.03
    kappa = 1.0 + 0.2 * rng.random(nv)
    c0 = 0.05 * rng.standard_normal(nv)
    dm = DeviceMesh(0)
    dm.set_topology(nv, tri)
    dm.set_positions(pos)
    dm.set_leaflet(L.LEAFLET_OUT, div_sign=1.0, kappa=kappa, c0=c0, k_tilt=3.0, facet_keep=keep.astype(np.uint8),
                   interior=interior.astype(np.uint8), base_zero=base_zero.astype(np.uint8), consistent=True)
    dm.upload(L.ARR_TILTS_OUT, tilts)
    g, tg = np.zeros_like(pos), np.zeros_like(pos)
    e_bt = rl.leaflet_bending_tilt_energy_and_gradient(pos, tri, tilts, kappa, c0, sign=1.0, keep=keep, interior=interior,
                                                       base_zero=base_zero, is_boundary=None, grad=g, tilt_grad=tg)
    e_t = rl.leaflet_tilt_energy_and_gradient(pos, tri, tilts, 3.0, keep=keep, consistent=True, grad=g, tilt_grad=tg)
    got_bt, got_t, _ = dm.eval_leaflet(L.LEAFLET_OUT, L.MOD_TILT | L.MOD_BENDING_TILT)
    _close(got_bt, e_bt)
    _close(got_t, e_t)
    assert rel_err(dm.download(L.ARR_GRAD), g) <= 2e-12
    assert rel_err(dm.download(L.ARR_TILT_GRAD_OUT), tg) <= TOL
    dm.close()


def _single_field_smoothness(factory_patch=None):
    from membrane_solver_b200.geometry.array_mesh import ArrayMesh, GlobalParams
    from membrane_solver_b200.modules.energy import tilt_smoothness as mod

    g = np.load(os.path.join(GOLDEN, "p1_vertex.npz"))
    gp = GlobalParams({"tilt_smoothness_rigidity": float(g["smooth_k"])})
    mesh = ArrayMesh(g["pos"], g["tri"], global_params=gp, tilts=g["tilts"])
    pos = mesh.positions_view()
    tg = np.full_like(pos, 0.5)
    grad = np.full_like(pos, -1.0)
    e = mod.compute_energy_and_gradient_array(mesh, gp, None, positions=pos, index_map=mesh.vertex_index_to_row,
                                              grad_arr=grad, tilts=g["tilts"], tilt_grad_arr=tg)
    _close(e, float(g["smooth_E"]))
    assert rel_err(tg - 0.5, g["smooth_tg"]) <= 4 * TOL
    assert np.all(grad == -1.0)                          # no shape gradient
    _close(mod.compute_energy_array(mesh, gp, None, positions=pos, index_map=mesh.vertex_index_to_row),
           float(g["smooth_E"]))
    assert mod.compute_energy_array(mesh, GlobalParams({}), None, positions=pos, index_map=None) == 0.0


def test_single_field_tilt_smoothness_on_emulated_device(monkeypatch):
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200.runtime import device_state
    from oracle import ref_leaflet as rl

    g = np.load(os.path.join(GOLDEN, "p1_vertex.npz"))
    tg = np.zeros_like(g["pos"])
    e = rl.leaflet_tilt_smoothness_energy_and_gradient(g["pos"], g["tri"], g["tilts"], float(g["smooth_k"]), tilt_grad=tg)
    _close(e, float(g["smooth_E"]))
    assert rel_err(tg, g["smooth_tg"]) <= TOL
    monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    _single_field_smoothness()


@pytest.mark.gpu
def test_single_field_tilt_smoothness_on_device():
    _single_field_smoothness()


@pytest.mark.parametrize("solver", ["gd", "cg"])
def test_tilt_relaxer_stop_conditions_on_emulated_device(relax_gold, solver):
    """Loop control of the relaxer: tolerance reached before the first step, every row fixed, zero step size."""
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.runtime.device_tilt_relaxer import DeviceTiltRelaxer

    g, p = relax_gold, "gd5_"
    pos, tri = g[p + "pos"], g[p + "tri"]

    def build(fixed_all=False):
        dm = FakeDeviceMesh(0)
        dm.set_topology(pos.shape[0], tri, is_boundary=g[p + "is_boundary"].astype(np.uint8))
        dm.set_positions(pos)
        for leaf, d in _relax_leaflets(g, "gd5").items():
            which = L.LEAFLET_IN if leaf == "in" else L.LEAFLET_OUT
            dm.set_leaflet(which, div_sign=-1.0 if leaf == "in" else 1.0, kappa=d["kappa"], c0=d["c0"], k_tilt=d["k_tilt"],
                           facet_keep=d["keep"].astype(np.uint8), interior=d["interior"].astype(np.uint8),
                           base_zero=d["base_zero"].astype(np.uint8))
            fx = np.ones(pos.shape[0], np.uint8) if fixed_all else g[p + f"fixed_{leaf}"].astype(np.uint8)
            dm.set_leaflet_fixed(which, fx)
            dm.upload(L.ARR_TILTS_IN if leaf == "in" else L.ARR_TILTS_OUT, g[p + f"tilts_{leaf}0"])
        return dm

    kw = dict(solver=solver, k_smooth={"in": 1.0, "out": 1.0}, area_kept_only={"out": True})
    dm = build()
    st = DeviceTiltRelaxer(dm).relax(max_iters=5, step_size=0.15, tol=1e9, **kw)
    assert st["stop_reason"] == "converged" and st["accepted_steps"] == 0
    assert abs(st["initial_gradient_norm"] - float(g[p + "initial_gradient_norm"])) <= 1e-9
    st = DeviceTiltRelaxer(dm).relax(max_iters=5, step_size=0.0, **kw)
    assert st["stop_reason"] == "step_size_zero"
    dm = build(fixed_all=True)
    before = dm.download(L.ARR_TILTS_IN)
    st = DeviceTiltRelaxer(dm).relax(max_iters=5, step_size=0.15, **kw)
    assert st["stop_reason"] == "zero_gradient" and st["accepted_steps"] == 0
    # only the tangent projection touched the fields
    n = dm.vnormals
    assert np.allclose(dm.download(L.ARR_TILTS_IN), before - (before * n).sum(axis=1)[:, None] * n, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("state", ["r1", "r1c"])
def test_device_leaflet_pair_equals_two_evaluations(gold, state):
    """Both leaflets in one call (one cooperative launch on this mesh): gradients bitwise those of two evaluations,
    energies equal to rounding (different grouping of the block sums), results read back with one copy."""
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh

    ins = {leaf: _inputs(gold, state, leaf) for leaf in LEAFLETS}
    i = ins["in"]
    dm = DeviceMesh(0)
    dm.set_topology(i["pos"].shape[0], i["tri"], is_boundary=i["is_boundary"].astype(np.uint8))
    dm.set_positions(i["pos"])
    for leaf, which, arr in (("in", L.LEAFLET_IN, L.ARR_TILTS_IN), ("out", L.LEAFLET_OUT, L.ARR_TILTS_OUT)):
        d = ins[leaf]
        dm.set_leaflet(which, div_sign=d["sign"], kappa=d["kappa"], c0=d["c0"], k_tilt=d["k_tilt"], k_smooth=d["k_smooth"],
                       facet_keep=d["keep"], interior=d["interior"], base_zero=d["base_zero"], consistent=d["consistent"])
        dm.upload(arr, d["tilts"])
    mods = L.MOD_TILT | L.MOD_BENDING_TILT | L.MOD_TILT_SMOOTHNESS
    e_in = dm.eval_leaflet(L.LEAFLET_IN, mods)
    e_out = dm.eval_leaflet(L.LEAFLET_OUT, mods, accumulate=L.ACC_GRAD)
    want = (dm.download(L.ARR_GRAD), dm.download(L.ARR_TILT_GRAD_IN), dm.download(L.ARR_TILT_GRAD_OUT))
    dm.upload(L.ARR_GRAD, np.full_like(i["pos"], 3.0))
    dm.eval_leaflet_pair(mods, want_grad=True, want_tilt_grad=True)
    res = dm.leaflet_results()
    for k in range(3):
        _close(res[L.LEAFLET_IN, k], e_in[k])
        _close(res[L.LEAFLET_OUT, k], e_out[k])
    got = (dm.download(L.ARR_GRAD), dm.download(L.ARR_TILT_GRAD_IN), dm.download(L.ARR_TILT_GRAD_OUT))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # tilt-only: the shape gradient array is left alone; gradient norms join the same read-back
    dm.eval_leaflet_pair(mods, want_grad=False, want_tilt_grad=True)
    dm.leaflet_gradient_norm2(L.LEAFLET_IN, read=False)
    dm.leaflet_gradient_norm2(L.LEAFLET_OUT, read=False)
    res = dm.leaflet_results()
    assert np.array_equal(dm.download(L.ARR_GRAD), want[0])
    tgi = gold[f"{state}_in_tg_bt_tiltonly"] + gold[f"{state}_in_tg_tilt_tiltonly"] + gold[f"{state}_in_tg_smooth_tiltonly"]
    assert rel_err(dm.download(L.ARR_TILT_GRAD_IN), tgi) <= TOL
    assert abs(res[L.LEAFLET_IN, 3] - float((dm.download(L.ARR_TILT_GRAD_IN) ** 2).sum())) <= 1e-10 * res[L.LEAFLET_IN, 3]
    dm.close()


# ------------------------------------------------ leaflet tilt relaxation WITH tilt constraint modules (row f3)
CONSTRAINED_CASES = ("cgd4", "ccg5", "cgd6i2", "ccg4pass")


@pytest.fixture(scope="module")
def constrained_gold():
    return np.load(os.path.join(GOLDEN, "tilt_relaxation_constrained.npz"))


class _ReplayedConstraintManager:
    """The reference's constraint manager as recorded by ``generate_golden.py tilt_relaxation_constrained_vectors``:
    call k of each hook checks that the arrays the device hands over are the ones the reference handed to its
    constraint manager (1e-10 of the array's scale) and answers with what the reference's modules
    (``tilt_thetaB_boundary_in``, ``rim_slope_match_out``) answered -- a few rim rows change."""

    def __init__(self, g, case):
        self.g, self.case, self.kg, self.kr = g, case, 0, 0

    def _same(self, got, key):
        want = self.g[key]
        scale = max(1.0, float(np.abs(want).max()))
        assert float(np.abs(got - want).max()) <= 1e-10 * scale, (key, float(np.abs(got - want).max()))

    def _same_norms(self, arrays, key):
        want = self.g[key]
        got = np.array([np.linalg.norm(x) for x in arrays])
        assert np.all(np.abs(got - want) <= 1e-10 * np.maximum(1.0, want)), (key, got, want)

    def gradient_hook(self, g_in, g_out, t_in, t_out):
        q = f"{self.case}_ghook_{self.kg}_"
        assert self.kg < int(self.g[self.case + "_n_ghook"])
        if q + "g_in" in self.g.files:      # the first calls are stored in full, later ones as norms + changed rows
            self._same(g_in, q + "g_in"), self._same(g_out, q + "g_out")
            self._same(t_in, q + "t_in"), self._same(t_out, q + "t_out")
        self._same_norms((g_in, g_out, t_in, t_out), q + "norms")
        self._same(g_in[self.g[q + "rows_in"]], q + "old_in"), self._same(g_out[self.g[q + "rows_out"]], q + "old_out")
        g_in[self.g[q + "rows_in"]] = self.g[q + "new_in"]
        g_out[self.g[q + "rows_out"]] = self.g[q + "new_out"]
        self.kg += 1

    def refresh_hook(self, t_in, t_out):
        q = f"{self.case}_refresh_{self.kr}_"
        assert self.kr < int(self.g[self.case + "_n_refresh"])
        if q + "t_in" in self.g.files:
            self._same(t_in, q + "t_in"), self._same(t_out, q + "t_out")
        self._same_norms((t_in, t_out), q + "norms")
        self._same(t_in[self.g[q + "rows_in"]], q + "old_in"), self._same(t_out[self.g[q + "rows_out"]], q + "old_out")
        t_in, t_out = np.array(t_in), np.array(t_out)
        t_in[self.g[q + "rows_in"]] = self.g[q + "new_in"]
        t_out[self.g[q + "rows_out"]] = self.g[q + "new_out"]
        self.kr += 1
        return t_in, t_out


def _constrained_device_relaxation(g, case, factory):
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.runtime.device_tilt_relaxer import DeviceTiltRelaxer

    p = case + "_"
    pos, tri = g[p + "pos"], g[p + "tri"]
    dm = factory(0)
    dm.set_topology(pos.shape[0], tri, is_boundary=g[p + "is_boundary"].astype(np.uint8))
    dm.set_positions(pos)
    for leaf, d in _relax_leaflets(g, case).items():
        which = L.LEAFLET_IN if leaf == "in" else L.LEAFLET_OUT
        dm.set_leaflet(which, div_sign=-1.0 if leaf == "in" else 1.0, kappa=d["kappa"], c0=d["c0"], k_tilt=d["k_tilt"],
                       facet_keep=d["keep"].astype(np.uint8), interior=d["interior"].astype(np.uint8),
                       base_zero=d["base_zero"].astype(np.uint8))
        dm.set_leaflet_fixed(which, g[p + f"fixed_{leaf}"].astype(np.uint8))
        dm.upload(L.ARR_TILTS_IN if leaf == "in" else L.ARR_TILTS_OUT, g[p + f"tilts_{leaf}0"])
    cm = _ReplayedConstraintManager(g, case)
    relaxer = DeviceTiltRelaxer(dm, gradient_hook=cm.gradient_hook, refresh_hook=cm.refresh_hook,
                                projection_interval=int(g[p + "interval"]), projection_cadence=str(g[p + "cadence"]))
    st = relaxer.relax(max_iters=int(g[p + "steps"]), step_size=float(g[p + "step_size"]), solver=str(g[p + "solver"]),
                       preconditioner=True, gd_fallback=False,
                       k_smooth={"in": float(g[p + "k_smooth_in"]), "out": float(g[p + "k_smooth_out"])},
                       area_kept_only={"out": not bool(np.all(g[p + "out_keep"]))})
    # every recorded call of the reference's constraint manager was consumed, in order
    assert (cm.kg, cm.kr) == (int(g[p + "n_ghook"]), int(g[p + "n_refresh"]))
    assert relaxer.hook_calls == {"gradient": cm.kg, "refresh": cm.kr}
    _relax_stats_match(st, g, case)
    assert st["stop_reason"] == str(g[p + "stop_reason"])
    assert np.max(np.abs(dm.download(L.ARR_TILTS_IN) - g[p + "tilts_in1"])) <= 1e-10
    assert np.max(np.abs(dm.download(L.ARR_TILTS_OUT) - g[p + "tilts_out1"])) <= 1e-10
    dm.close()


@pytest.mark.parametrize("case", CONSTRAINED_CASES)
def test_constrained_tilt_relaxer_on_emulated_device(constrained_gold, case):
    from fake_device import FakeDeviceMesh

    _constrained_device_relaxation(constrained_gold, case, FakeDeviceMesh)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CONSTRAINED_CASES)
def test_constrained_tilt_relaxer_on_device(constrained_gold, case):
    from membrane_solver_b200.context import DeviceMesh

    _constrained_device_relaxation(constrained_gold, case, DeviceMesh)


@pytest.mark.parametrize("backend", [pytest.param("emulator", id="emulator"),
                                     pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)])
@pytest.mark.parametrize("case", ["gd5", "cg6"])
def test_mesh_level_relax_leaflet_tilts(relax_gold, case, backend, monkeypatch):
    """``device_tilt_relaxer.relax_leaflet_tilts``: the mesh-level call (global parameters in, relaxed fields written
    back to the mesh) lands on the reference's relaxed fields."""
    from membrane_solver_b200.geometry.array_mesh import ArrayMesh, GlobalParams, ParamResolver
    from membrane_solver_b200.runtime import device_state
    from membrane_solver_b200.runtime.device_tilt_relaxer import relax_leaflet_tilts

    if backend == "emulator":
        from fake_device import FakeDeviceMesh

        monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    g, p = relax_gold, case + "_"
    leaflets = {}
    for leaf, d in _relax_leaflets(g, case).items():
        leaflets[leaf] = dict(keep_bt=d["keep"], keep_tilt=d["keep"], interior=d["interior"], base_zero=d["base_zero"],
                              kappa=d["kappa"], c0=d["c0"], k_tilt=d["k_tilt"])
    gp = GlobalParams(tilt_solver=str(g[p + "solver"]), tilt_inner_steps=int(g[p + "steps"]),
                      tilt_step_size=float(g[p + "step_size"]), tilt_tol=0.0,
                      bending_modulus_in=float(g[p + "k_smooth_in"]), bending_modulus_out=float(g[p + "k_smooth_out"]))
    mesh = ArrayMesh(g[p + "pos"], g[p + "tri"], global_params=gp, tilts_in=g[p + "tilts_in0"],
                     tilts_out=g[p + "tilts_out0"], leaflets=leaflets, tilt_fixed_in=g[p + "fixed_in"],
                     tilt_fixed_out=g[p + "fixed_out"])
    mesh._boundary = set(np.flatnonzero(g[p + "is_boundary"]).tolist())
    st = relax_leaflet_tilts(mesh, gp, ParamResolver(gp))
    _relax_stats_match(st, g, case)
    assert np.max(np.abs(mesh.tilts_in_view() - g[p + "tilts_in1"])) <= 1e-10
    assert np.max(np.abs(mesh.tilts_out_view() - g[p + "tilts_out1"])) <= 1e-10
