"""Drop-in proof against the REAL reference (build container only: ``/root/reference`` does not
exist on the GPU box, where this file skips).

The reference's own ``Minimizer`` / ``EnergyModuleManager`` / ``ConstraintModuleManager`` run
unmodified; ``membrane_solver_b200.runtime.energy_manager.install()`` binds the B200 plugin twins
under the reference's module names.  On the CPU tier the device is the host emulator, so what is
proven here is the CONTRACT (names, signatures, accumulate semantics, version-keyed residency)
and the per-facet code; the kernels themselves are proven by the gpu tier."""

import importlib
import os
import sys

import numpy as np
import pytest

REF = os.environ.get("MEMBRANE_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "runtime")),
                                reason="the reference tree is only present in the build container")

CASES = [("meshes/cube.json", 1, 6), ("meshes/bending_cube.yaml", 1, 3), ("meshes/catenoid.json", 1, 4)]


def _ref_imports():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from geometry.geom_io import load_data, parse_geometry
    from runtime.constraint_manager import ConstraintModuleManager
    from runtime.energy_manager import EnergyModuleManager
    from runtime.minimizer import Minimizer
    from runtime.refinement import refine_triangle_mesh
    from runtime.steppers.gradient_descent import GradientDescent

    return load_data, parse_geometry, ConstraintModuleManager, EnergyModuleManager, Minimizer, refine_triangle_mesh, GradientDescent


def _build(path, levels, seed=5):
    load_data, parse_geometry, CMM, EMM, Minimizer, refine, GD = _ref_imports()
    mesh = parse_geometry(load_data(os.path.join(REF, path)))
    for _ in range(levels):
        mesh = refine(mesh)
    rng = np.random.default_rng(seed)
    for v in mesh.vertices.values():
        if not getattr(v, "fixed", False):
            v.position = np.asarray(v.position, dtype=float) + 0.01 * rng.normal(size=3)
    mesh.increment_version()
    gp = mesh.global_parameters
    mini = Minimizer(mesh, gp, GD(), EMM(mesh.energy_modules), CMM(mesh.constraint_modules), quiet=True)
    return mesh, mini


@pytest.fixture
def b200_installed(monkeypatch):
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200.runtime import device_state, energy_manager

    _ref_imports()
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.startswith("modules.energy.")}
    import modules.constraints.volume as ref_cv

    saved_cv = (ref_cv.constraint_gradients_array, ref_cv.constraint_gradients)
    monkeypatch.setattr(device_state, "DEVICE_MESH_FACTORY", FakeDeviceMesh)
    yield energy_manager.install
    for name in ("surface", "volume", "bending", "tilt"):
        key = f"modules.energy.{name}"
        if saved.get(key) is not None:
            sys.modules[key] = saved[key]
        else:
            sys.modules.pop(key, None)
            importlib.import_module(key)
        import modules.energy as pkg

        setattr(pkg, name, sys.modules[key])
    ref_cv.constraint_gradients_array, ref_cv.constraint_gradients = saved_cv


@pytest.mark.parametrize("path,levels,steps", CASES, ids=[c[0].split("/")[-1] for c in CASES])
def test_reference_minimizer_runs_on_b200_plugins(b200_installed, path, levels, steps):
    # 1. the unmodified reference
    mesh_ref, mini_ref = _build(path, levels)
    e_ref, g_ref = mini_ref.compute_energy_and_gradient_array()
    bd_ref = mini_ref.compute_energy_breakdown()
    res_ref = mini_ref.minimize(n_steps=steps)
    # 2. the same driver code with the B200 plugins bound under the reference's names
    bound = b200_installed()
    assert "modules.energy.surface" in bound
    mesh, mini = _build(path, levels)
    for name, mod in zip(mini.energy_module_names if hasattr(mini, "energy_module_names") else mesh.energy_modules,
                         mini.energy_modules):
        if name in ("surface", "volume", "bending", "tilt"):
            assert mod.__name__.startswith("membrane_solver_b200."), (name, mod.__name__)
    e, g = mini.compute_energy_and_gradient_array()
    assert abs(e - e_ref) <= 1e-12 * max(1.0, abs(e_ref))
    assert np.max(np.abs(g - g_ref)) <= 1e-11 * max(1.0, np.max(np.abs(g_ref)))
    bd = mini.compute_energy_breakdown()
    for k, v in bd_ref.items():
        assert abs(bd[k] - v) <= 1e-11 * max(1.0, abs(v)), k
    # topology was packed once; every evaluation re-sent positions only
    assert mesh._b200_state.uploads == 1
    # 3. minimised-energy trajectory (BASELINE.json north_star: agree within 1e-9)
    res = mini.minimize(n_steps=steps)
    e_end_ref, e_end = mini_ref.compute_energy(), mini.compute_energy()
    assert abs(e_end - e_end_ref) <= 1e-9 * max(1.0, abs(e_end_ref)), (e_end, e_end_ref)
    p_ref = np.array([mesh_ref.vertices[v].position for v in sorted(mesh_ref.vertices)])
    p = np.array([mesh.vertices[v].position for v in sorted(mesh.vertices)])
    assert np.max(np.abs(p - p_ref)) <= 1e-9
    _ = (res, res_ref)
