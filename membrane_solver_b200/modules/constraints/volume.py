"""Volume constraint gradients on the B200 path.

Twin of ``modules/constraints/volume.py:13-66``: in ``lagrange`` mode every body with a target
volume contributes its dense ``dV/dx`` to the KKT projection of the constraint manager
(``runtime/constraint_manager.py:174-315``).  ``enforce_constraint`` (the Newton projection of
the vertex positions, ``constraints/volume.py:69-149``) is a next-row item
and stays with the reference.
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


__all__ = ["constraint_gradients_array", "constraint_gradients"]
