"""Body-volume penalty plugin on the B200 path.

Twin of ``modules/energy/volume.py:94-128`` (penalty mode only; in ``lagrange`` mode the
volume enters through ``modules/constraints/volume.py``): ``E = 1/2 k (V - V0)^2`` per body with
``V = sum_f (v1 x v2).v0 / 6`` and ``grad += k (V - V0) dV/dx`` (``geometry/body.py:192-252``).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _common as C
from ...runtime.device_state import body_entries, get_state, positions_array, triangle_rows

B200_MODULE = C.L.MOD_VOLUME
USES_TILT = False


def body_volume_and_gradient(mesh, body_index: int, positions, want_grad: bool = True):
    """(V, dV/dx or None) of one body of the mesh on the device."""
    pos = positions_array(positions)
    st = get_state(mesh, pos)
    bodies = body_entries(mesh)
    _, rows, _ = bodies[body_index]
    if rows is None:
        raise C.L.B200Error("the B200 path needs triangulated bodies (body rows are not available)")
    if len(bodies) == 1 and st.body_rows is not None:
        gC = C.scratch_like(pos) if want_grad else None
        opts = st.dm.options(C.L.MOD_VOLUME, want_grad=want_grad)
        res = st.dm.eval_host(opts, pos, volgrad=gC)
        return res.volume, gC
    # several bodies share the mesh: evaluate this body's facet rows with the stateless shim
    tri = np.ascontiguousarray(triangle_rows(mesh)[rows], dtype=np.int32)
    gC = np.zeros_like(pos) if want_grad else None
    vol = ctypes.c_double(0.0)
    C.L.check(C.L.lib().ms_volume_and_gradient(pos.shape[0], tri.shape[0], C.L.dptr(pos), C.L.iptr(tri), 1.0,
                                               C.L.dptr(gC), ctypes.byref(vol)))
    return float(vol.value), gC


def _stiffness(body, global_params, param_resolver):
    k = param_resolver.get(body, "volume_stiffness") if param_resolver is not None else None
    if k is None:
        k = C.gp_get(global_params, "volume_stiffness", 0.0)
    return float(k)


def _target(body):
    t = getattr(body, "target_volume", None)
    if t is None:
        t = (getattr(body, "options", None) or {}).get("target_volume", 0)
    return float(t)


def b200_configure(state, mesh, global_params, param_resolver) -> dict:
    bodies = body_entries(mesh)
    if len(bodies) != 1 or state.body_rows is None:
        return {"unfused": True}
    body = bodies[0][0]
    return {"constraint_mode": 1, "k_vol": _stiffness(body, global_params, param_resolver),
            "v_target": _target(body)}


def b200_energy(result, k_vol: float = 0.0, v_target: float = 0.0) -> float:
    d = result.volume - v_target
    return 0.5 * k_vol * d * d


def compute_energy_and_gradient_array(mesh, global_params, param_resolver, *, positions, index_map, grad_arr) -> float:
    if C.gp_get(global_params, "volume_constraint_mode", "lagrange") != "penalty":
        return 0.0
    energy = 0.0
    for i, (body, _, _) in enumerate(body_entries(mesh)):
        k, v0 = _stiffness(body, global_params, param_resolver), _target(body)
        vol, gC = body_volume_and_gradient(mesh, i, positions, want_grad=grad_arr is not None)
        delta = vol - v0
        energy += 0.5 * k * delta * delta
        if grad_arr is not None:
            np.add(grad_arr, (k * delta) * gC, out=grad_arr)
    return float(energy)


def compute_energy_array(mesh, global_params, param_resolver=None, *, positions, index_map) -> float:
    return compute_energy_and_gradient_array(mesh, global_params, param_resolver, positions=positions,
                                             index_map=index_map, grad_arr=None)


def calculate_volume_energy(mesh, global_params) -> float:
    """``volume.calculate_volume_energy`` (``volume.py:13-38``)."""
    return compute_energy_array(mesh, global_params, None, positions=mesh.positions_view(),
                                index_map=mesh.vertex_index_to_row)


compute_energy_and_gradient = C.dict_api(compute_energy_and_gradient_array)

__all__ = ["compute_energy_and_gradient_array", "compute_energy_array", "compute_energy_and_gradient",
           "calculate_volume_energy", "body_volume_and_gradient"]
