"""Volume constraint gradients on the B200 path.

Twin of ``modules/constraints/volume.py:13-66``: in ``lagrange`` mode every body with a target
volume contributes its dense ``dV/dx`` to the KKT projection of the constraint manager
(``runtime/constraint_manager.py:174-315``).  ``enforce_constraint`` is the twin of the hard projection of
the vertex positions (``constraints/volume.py:69-149``) on dense arrays: volume and ``dV/dx`` come from the device,
the Newton step ``x -= (V - V0) / (|dV/dx|^2 + 1e-12) dV/dx`` is one array expression, and the positions go back to
the mesh once per iteration -- instead of the reference's per-vertex dict loops.
"""

from __future__ import annotations

import numpy as np

from ..energy import _common as C
from ..energy.volume import body_volume_and_gradient
from ...runtime.device_state import body_entries


def _constrained(mesh):
    out = []
    for i, (body, _, _) in enumerate(body_entries(mesh)):
        target = getattr(body, "target_volume", None)
        if target is None:
            target = (getattr(body, "options", None) or {}).get("target_volume")
        if target is not None:
            out.append(i)
    return out


def constraint_gradients_array(mesh, global_params, *, positions, index_map):
    """Dense ``dV/dx`` of every constrained body, or None (``constraints/volume.py:43-66``)."""
    if C.gp_get(global_params, "volume_constraint_mode", "lagrange") != "lagrange":
        return None
    grads = [body_volume_and_gradient(mesh, i, positions)[1] for i in _constrained(mesh)]
    return grads or None


def constraint_gradients(mesh, global_params):
    """Dict form (``constraints/volume.py:13-40``)."""
    arrs = constraint_gradients_array(mesh, global_params, positions=mesh.positions_view(),
                                      index_map=mesh.vertex_index_to_row)
    if arrs is None:
        return None
    ids = list(mesh.vertex_ids)
    return [{int(v): g[r].copy() for r, v in enumerate(ids)} for g in arrs]


def _write_positions(mesh, positions: np.ndarray) -> None:
    """Hand the projected positions back to the mesh and bump its version (``mesh.increment_version()``)."""
    setter = getattr(mesh, "set_positions", None)
    if setter is not None:                     # ArrayMesh
        setter(positions)
        return
    vertices = mesh.vertices
    for row, vid in enumerate(mesh.vertex_ids):
        vertices[int(vid)].position = positions[row].copy()
    mesh.increment_version()


def enforce_constraint(mesh, tol: float = 1e-12, max_iter: int = 3, global_params=None, force_projection: bool = False,
                       **kwargs) -> None:
    """Hard volume projection of every body with a target volume (``constraints/volume.py:69-149``): same modes
    (``force_projection`` / ``lagrange`` / ``projection``), same iteration budget (12 for the ``finalize`` and
    ``mesh_operation`` contexts), same step and tolerance, fixed vertices stay."""
    mode = C.gp_get(global_params, "volume_constraint_mode", "lagrange") if global_params is not None else "projection"
    if not (force_projection or mode in ("lagrange", "projection")):
        return
    if kwargs.get("context", "minimize") in ("finalize", "mesh_operation"):
        max_iter = max(int(max_iter), 12)
    entries = body_entries(mesh)
    fm = getattr(mesh, "fixed_mask", None)
    fixed = None
    if fm is not None:
        fixed = np.asarray(fm() if callable(fm) else fm, dtype=bool)
        if not fixed.any():
            fixed = None
    for i in _constrained(mesh):
        body, _, target = entries[i]
        if target is None:
            target = (getattr(body, "options", None) or {}).get("target_volume")
        for _ in range(int(max_iter)):
            positions = np.array(mesh.positions_view(), dtype=np.float64)
            vol, gc = body_volume_and_gradient(mesh, i, positions)
            delta = float(vol) - float(target)
            if abs(delta) < tol:
                break
            lam = delta / (float(np.vdot(gc, gc)) + 1e-12)
            step = lam * gc
            if fixed is not None:
                step[fixed] = 0.0
            _write_positions(mesh, positions - step)


__all__ = ["constraint_gradients_array", "constraint_gradients", "enforce_constraint"]
