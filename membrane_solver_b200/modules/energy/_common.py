"""Shared plumbing of the B200 energy plugins: parameter lookup, device evaluation, legacy dict API."""

from __future__ import annotations

import numpy as np

from ... import _lib as L
from ...runtime.device_state import get_state, positions_array


def gp_get(global_params, name, default=None):
    """``global_params.get(name, default)`` for GlobalParameters objects and plain dicts."""
    getter = getattr(global_params, "get", None)
    if getter is not None:
        val = getter(name, default)
        return default if val is None else val
    return getattr(global_params, name, default)


def energy_model(global_params) -> str:
    """``bending_params._energy_model`` (``modules/energy/bending_params.py:18-21``)."""
    model = str(gp_get(global_params, "bending_energy_model", "helfrich") or "helfrich").lower().strip()
    return "helfrich" if model == "helfrich" else "willmore"


def gradient_mode(global_params) -> str:
    """``bending_params._gradient_mode`` (``bending_params.py:24-31``)."""
    mode = str(gp_get(global_params, "bending_gradient_mode", "analytic") or "analytic").lower().strip()
    if mode in {"fd", "finite_difference"}:
        return "finite_difference"
    return "analytic" if mode == "analytic" else "approx"


def spontaneous_curvature(global_params) -> float:
    """``bending_params._spontaneous_curvature`` (``bending_params.py:34-38``)."""
    val = gp_get(global_params, "spontaneous_curvature")
    if val is None:
        val = gp_get(global_params, "intrinsic_curvature", 0.0)
    return float(val or 0.0)


def per_vertex_bending_params(mesh, global_params, model: str):
    """(kappa, c0) per vertex row with ``vertex.options`` overrides (``bending_params.py:41-115``).

    Uniform values come back as Python floats so that the device uses its uniform path."""
    kappa_default = float(gp_get(global_params, "bending_modulus", 0.0) or 0.0)
    c0_default = spontaneous_curvature(global_params) if model == "helfrich" else 0.0
    key = (int(getattr(mesh, "_vertex_ids_version", 0) or 0), model, kappa_default, c0_default)
    cached = getattr(mesh, "_b200_bending_param_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    rows_k, vals_k, rows_c, vals_c = [], [], [], []
    index = mesh.vertex_index_to_row
    for vid, vertex in (getattr(mesh, "vertices", None) or {}).items():
        opts = getattr(vertex, "options", None) or {}
        if not opts:
            continue
        row = index.get(int(vid))
        if row is None:
            continue
        if "bending_modulus" in opts:
            try:
                vals_k.append(float(opts["bending_modulus"]))
                rows_k.append(row)
            except (TypeError, ValueError):
                pass
        if model == "helfrich":
            name = "spontaneous_curvature" if "spontaneous_curvature" in opts else (
                "intrinsic_curvature" if "intrinsic_curvature" in opts else None)
            if name is not None:
                try:
                    vals_c.append(float(opts[name]))
                    rows_c.append(row)
                except (TypeError, ValueError):
                    pass
    n = len(mesh.vertex_ids)
    kappa, c0 = kappa_default, c0_default
    if rows_k:
        kappa = np.full(n, kappa_default)
        kappa[np.asarray(rows_k, dtype=np.int64)] = vals_k
    if rows_c:
        c0 = np.full(n, c0_default)
        c0[np.asarray(rows_c, dtype=np.int64)] = vals_c
    try:
        mesh._b200_bending_param_cache = (key, kappa, c0)
    except AttributeError:
        pass
    return kappa, c0


def max_abs(x) -> float:
    return float(np.max(np.abs(x))) if np.ndim(x) else abs(float(x))


def device_eval(mesh, positions, modules: int, *, flags: int = 0, want_grad: bool = True,
                grad: np.ndarray | None = None, volgrad: np.ndarray | None = None,
                tilt_grad: np.ndarray | None = None, constraint_mode: int = -1, k_vol: float = 0.0,
                v_target: float = 0.0, diagnostics: bool = False, configure=None):
    """One evaluation through the C ABI with host buffers.  Returns (state, EvalResult)."""
    pos = positions_array(positions)
    st = get_state(mesh, pos)
    if configure is not None:
        configure(st)
    opts = st.dm.options(modules, flags=flags, want_grad=want_grad, constraint_mode=constraint_mode, k_vol=k_vol,
                         v_target=v_target, diagnostics=diagnostics)
    res = st.dm.eval_host(opts, pos, grad=grad, volgrad=volgrad, tilt_grad=tilt_grad)
    return st, res


def scratch_like(positions) -> np.ndarray:
    return np.empty((int(np.shape(positions)[0]), 3), dtype=np.float64)


def accumulate(dst: np.ndarray, src: np.ndarray) -> None:
    """``dst += src`` honouring the reference contract that plugins add into caller-owned arrays."""
    if dst.shape != src.shape:
        raise ValueError(f"gradient array has shape {dst.shape}, expected {src.shape}")
    np.add(dst, src, out=dst)


def dict_api(array_fn):
    """Build the legacy ``compute_energy_and_gradient(mesh, gp, resolver, *, compute_gradient)``
    on top of an array plugin function (``evaluation_manager.py:169-176`` uses it when present)."""

    def compute_energy_and_gradient(mesh, global_params, param_resolver, *, compute_gradient: bool = True):
        positions = mesh.positions_view()
        grad = np.zeros_like(positions) if compute_gradient else None
        energy = array_fn(mesh, global_params, param_resolver, positions=positions,
                          index_map=mesh.vertex_index_to_row, grad_arr=grad)
        if not compute_gradient:
            return float(energy), {}
        return float(energy), {int(vid): grad[row].copy() for row, vid in enumerate(mesh.vertex_ids)}

    return compute_energy_and_gradient


__all__ = ["L", "get_state", "positions_array", "gp_get", "energy_model", "gradient_mode", "spontaneous_curvature", "per_vertex_bending_params",
           "max_abs", "device_eval", "scratch_like", "accumulate", "dict_api"]
