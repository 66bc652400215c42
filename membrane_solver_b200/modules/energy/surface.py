"""Surface-tension plugin on the B200 path.

Twin of ``modules/energy/surface.py:100-239``: ``E = sum_f gamma_f A_f`` over facets with
``|n| >= 1e-12`` and ``dE/dv0 = gamma/2 (v1 - v2) x n_hat`` (cyclic), accumulated into the
caller's ``grad_arr``.  The per-facet surface tensions come from
``mesh.get_facet_parameter_array("surface_tension")`` exactly like the reference; the kernel is
the fused patch kernel of ``libms_b200.so`` with only the surface bit set.  A polygonal
(non-triangle) mesh raises -- there is no CPU fallback on this path.
"""

from __future__ import annotations

import numpy as np

from . import _common as C

B200_MODULE = C.L.MOD_SURFACE
USES_TILT = False


def b200_configure(state, mesh, global_params, param_resolver) -> dict:
    """Send this module's parameters to the device state; returns eval-option overrides."""
    state.set_gamma(mesh.get_facet_parameter_array("surface_tension"))
    return {}


def b200_energy(result) -> float:
    return result.e_surface


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr) -> float:
    want_grad = grad_arr is not None
    tmp = C.scratch_like(positions) if want_grad else None
    _, res = C.device_eval(mesh, positions, B200_MODULE, want_grad=want_grad, grad=tmp,
                           configure=lambda st: b200_configure(st, mesh, global_params, param_resolver))
    if want_grad:
        C.accumulate(grad_arr, tmp)
    return res.e_surface


def compute_energy_array(mesh, global_params, param_resolver=None, *, positions, index_map) -> float:
    """Energy only (the reference has no such entry point for surface and pays a full gradient
    evaluation instead, ``evaluation_manager.py:201-210``)."""
    _, res = C.device_eval(mesh, positions, B200_MODULE, want_grad=False,
                           configure=lambda st: b200_configure(st, mesh, global_params, param_resolver))
    return res.e_surface


def calculate_surface_energy(mesh, global_params) -> float:
    """``surface.calculate_surface_energy`` (``surface.py:64-97``)."""
    return compute_energy_array(mesh, global_params, positions=mesh.positions_view(),
                                index_map=mesh.vertex_index_to_row)


compute_energy_and_gradient = C.dict_api(compute_energy_and_gradient_array)

__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient",
           "calculate_surface_energy"]
