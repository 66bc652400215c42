"""Device-resident leaflet tilt relaxation at frozen geometry.

Twin of the gradient-descent solver of ``TiltRelaxationManager.relax_leaflet_tilts``
(``runtime/steppers/tilt_relaxation.py:426-1057``: setup ``:630-668``, gradients ``:825-872``, loop ``:894-1055``)
and of its preconditioned conjugate-gradient solver (``:1057-1440``, Jacobi diagonal of
``runtime/preconditioners.py:64-146``, optional gradient-descent fallback) for a mesh whose tilt fields are
constrained by fixed rows and, optionally, by tilt CONSTRAINT modules through two host hooks (no axisymmetric
projection).  Positions, both tilt fields, their gradients and the trial fields stay on the device; per
iteration the host sees three scalars (energy, gradient norm, trial energy) and keeps the loop control:

    E0, g = tilt-only evaluation of the leaflet modules;  g[fixed] = 0;  stop on |g| == 0 or |g| < tol
    step = tilt_step_size;  up to 12 trials:  t' = P(t - step g), fixed rows kept;  accept if E(t') <= E0 else halve
    after an accepted step the fields pass through the tangent projection once more (the reference's per-step
    refresh writes them to the mesh and reads them back projected, ``:803-823``)

Tilt constraint modules (configs[3] ships ``tilt_thetaB_boundary_in`` and ``rim_slope_match_out``) enter the
reference's loop at exactly two places, and both act on a few rim rows through the constraint manager:

* ``gradient_hook(g_in, g_out, t_in, t_out)`` -- ``ConstraintModuleManager.apply_tilt_gradient_modifications_array``
  (``runtime/constraint_manager.py:651-825``, called at ``tilt_relaxation.py:848-861``): the KKT projection of the two
  tilt gradients in tilt space, BEFORE the fixed rows are zeroed and the norm is taken;
* ``refresh_hook(t_in, t_out) -> (t_in, t_out)`` -- ``enforce_tilt_constraints`` (``constraint_manager.py:827-841``):
  before the loop (``tilt_relaxation.py:612-618``), after every ``projection_interval``-th accepted step
  (``:803-823,1047-1052``; or once per pass with ``projection_cadence="per_pass"``, ``:1416-1417``) each time followed
  by the tangent projection, and once more at the very end WITHOUT projection (``:1473-1478``).

With hooks the four (nv,3) arrays the hook reads and the arrays it changes cross PCIe at those points (108 KB on
the caveolin mesh); every energy / gradient evaluation and every trial stays on the device.  The hooks are plain
callables, so the reference's own constraint manager serves as both (``reference_constraint_hooks``), and a recorded
sequence of hook results replays the reference's constrained trajectory without the reference
(``tests/golden/tilt_relaxation_constrained.npz``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

from .. import _lib as L

_WHICH = {"in": L.LEAFLET_IN, "out": L.LEAFLET_OUT}
_TILTS = {"in": L.ARR_TILTS_IN, "out": L.ARR_TILTS_OUT}
_GRADS = {"in": L.ARR_TILT_GRAD_IN, "out": L.ARR_TILT_GRAD_OUT}


def reference_constraint_hooks(mesh, global_params, constraint_manager, positions):
    """(gradient_hook, refresh_hook) that call the reference's own constraint manager on its mesh object: the
    drop-in form, for a process in which the reference is importable."""
    import numpy as np

    def gradient_hook(g_in, g_out, t_in, t_out):
        if hasattr(constraint_manager, "apply_tilt_gradient_modifications_array"):
            constraint_manager.apply_tilt_gradient_modifications_array(
                g_in, g_out, mesh, global_params, positions=positions, tilts_in=t_in, tilts_out=t_out)

    def refresh_hook(t_in, t_out):
        mesh.set_tilts_in_from_array(t_in)
        mesh.set_tilts_out_from_array(t_out)
        if hasattr(constraint_manager, "enforce_tilt_constraints"):
            constraint_manager.enforce_tilt_constraints(mesh, global_params=global_params)
        return np.array(mesh.tilts_in_view()), np.array(mesh.tilts_out_view())

    return gradient_hook, refresh_hook


@dataclass
class DeviceTiltRelaxer:
    """``dm``: a DeviceMesh with positions uploaded, both leaflets described (``set_leaflet``), their fixed rows
    (``set_leaflet_fixed``) and tilt fields (``ARR_TILTS_IN`` / ``_OUT``) in place."""

    dm: object
    leaflets: tuple = ("in", "out")
    modules: int = L.MOD_TILT | L.MOD_BENDING_TILT
    stats: dict = field(default_factory=dict)
    gradient_hook: object = None      # see the module docstring
    refresh_hook: object = None
    projection_interval: int = 1      # tilt_projection_interval
    projection_cadence: str = "per_step"   # tilt_projection_cadence: "per_step" | "per_pass"
    hook_calls: dict = field(default_factory=lambda: {"gradient": 0, "refresh": 0})

    def _energy(self, want_tilt_grad: bool) -> float:
        """Both leaflets are launched back to back; ONE synchronisation brings their energies."""
        self._launch(want_tilt_grad)
        res = self.dm.leaflet_results()
        return float(sum(res[_WHICH[name], :3].sum() for name in self.leaflets))

    def _launch(self, want_tilt_grad: bool) -> None:
        if tuple(self.leaflets) == ("in", "out"):       # both leaflets: one call (one launch on small meshes)
            self.dm.eval_leaflet_pair(self.modules, want_grad=False, want_tilt_grad=want_tilt_grad)
            return
        for name in self.leaflets:
            self.dm.eval_leaflet(_WHICH[name], self.modules, want_grad=False, want_tilt_grad=want_tilt_grad, read=False)

    def _fields(self, arrays) -> list:
        return [self.dm.download(arrays[name]) for name in ("in", "out")]

    def _apply_gradient_hook(self) -> None:
        """tilt_relaxation.py:848-861: the constraint manager edits both tilt gradients (host), then they go back."""
        g_in, g_out = self._fields(_GRADS)
        t_in, t_out = self._fields(_TILTS)
        self.gradient_hook(g_in, g_out, t_in, t_out)
        self.dm.upload(_GRADS["in"], g_in)
        self.dm.upload(_GRADS["out"], g_out)
        self.hook_calls["gradient"] += 1

    def _refresh(self, project: bool = True) -> None:
        """tilt_relaxation.py:803-823: fields -> mesh, enforce_tilt_constraints, fields <- mesh, tangent projection."""
        if self.refresh_hook is not None:
            t_in, t_out = self.refresh_hook(*self._fields(_TILTS))
            self.dm.upload(_TILTS["in"], t_in)
            self.dm.upload(_TILTS["out"], t_out)
            self.hook_calls["refresh"] += 1
        if project:
            for name in self.leaflets:
                self.dm.leaflet_project_tilts(_WHICH[name])

    def _after_accepted_step(self, accepted_steps: int) -> None:
        if self.projection_cadence == "per_step" and accepted_steps % max(1, int(self.projection_interval)) == 0:
            self._refresh()

    def _finish(self, st: dict) -> dict:
        if self.projection_cadence == "per_pass":
            self._refresh()
        if self.refresh_hook is not None:        # tilt_relaxation.py:1473-1478: enforced once more, not projected
            self._refresh(project=False)
        return st

    def _energy_and_gradient_norm(self) -> tuple[float, float]:
        self._launch(True)
        if self.gradient_hook is not None:
            if tuple(self.leaflets) != ("in", "out"):
                raise L.B200Error("tilt constraint hooks need both leaflets")
            res0 = self.dm.leaflet_results()      # energies of this evaluation, before the arrays are touched
            self._apply_gradient_hook()
            for name in self.leaflets:
                self.dm.leaflet_gradient_norm2(_WHICH[name], read=False)
            res = self.dm.leaflet_results()
            e = float(sum(res0[_WHICH[name], :3].sum() for name in self.leaflets))
            return e, math.sqrt(float(sum(res[_WHICH[name], 3] for name in self.leaflets)))
        for name in self.leaflets:
            self.dm.leaflet_gradient_norm2(_WHICH[name], read=False)
        res = self.dm.leaflet_results()
        e = float(sum(res[_WHICH[name], :3].sum() for name in self.leaflets))
        return e, math.sqrt(float(sum(res[_WHICH[name], 3] for name in self.leaflets)))

    def relax(self, *, max_iters: int, step_size: float, tol: float = 0.0, solver: str = "gd",
              preconditioner: bool = True, k_smooth: dict | None = None, area_kept_only: dict | None = None,
              gd_fallback: bool = False) -> dict:
        """``solver``: "gd" or "cg" (``tilt_solver``).  CG only: ``preconditioner`` (``tilt_cg_preconditioner`` jacobi /
        none), ``k_smooth`` = {"in": bending_modulus_in, "out": ...}, ``area_kept_only`` = {"out": True} when the
        outer leaflet has absent vertices (its barycentric areas then run over its own facets,
        ``tilt_relaxation.py:679-697``), ``gd_fallback`` (``tilt_cg_rejection_fallback``)."""
        if solver == "cg":
            return self._relax_cg(max_iters=max_iters, step_size=step_size, tol=tol, preconditioner=preconditioner,
                                  k_smooth=k_smooth or {}, area_kept_only=area_kept_only or {}, gd_fallback=gd_fallback)
        dm = self.dm
        st = dict(accepted_steps=0, rejected_steps=0, backtracking_steps=0, stop_reason="completed_max_iters",
                  initial_energy=0.0, final_energy=0.0, initial_gradient_norm=0.0, final_gradient_norm=0.0)
        self.stats = st
        if step_size <= 0.0:
            st["stop_reason"] = "step_size_zero"
            return st
        dm.update_vertex_normals()
        self._refresh()          # tilt_relaxation.py:612-618,662-663: enforce (with hooks), then project
        for _ in range(int(max_iters)):
            e0, gnorm = self._energy_and_gradient_norm()
            if st["accepted_steps"] == 0 and st["rejected_steps"] == 0:
                st["initial_energy"], st["initial_gradient_norm"] = e0, gnorm
            st["final_energy"], st["final_gradient_norm"] = e0, gnorm
            if gnorm == 0.0:
                st["stop_reason"] = "zero_gradient"
                break
            if tol > 0.0 and gnorm < tol:
                st["stop_reason"] = "converged"
                break
            step, accepted, e1 = float(step_size), False, e0
            for attempt in range(12):
                if attempt:
                    st["backtracking_steps"] += 1
                for name in self.leaflets:
                    dm.leaflet_make_trial(_WHICH[name], step)
                    dm.leaflet_swap_trial(_WHICH[name])        # evaluate at the trial fields
                e1 = self._energy(False)
                if e1 <= e0:
                    accepted = True
                    break
                for name in self.leaflets:
                    dm.leaflet_swap_trial(_WHICH[name])        # rejected: back to the base fields
                step *= 0.5
                if step < 1e-16:
                    break
            if not accepted:
                st["rejected_steps"] += 1
                st["stop_reason"] = "line_search_rejected"
                break
            st["accepted_steps"] += 1
            st["step_size_last_accepted"] = step
            self._after_accepted_step(st["accepted_steps"])
            st["final_energy"] = e1
        return self._finish(st)

    # -- preconditioned conjugate gradients (tilt_relaxation.py:1057-1440) ------------------------------
    def _gradients(self) -> tuple[float, float]:
        return self._energy_and_gradient_norm()

    def _line_search(self, e0: float, step_size: float, along_direction: bool, st: dict):
        step = float(step_size)
        for attempt in range(12):
            if attempt:
                st["backtracking_steps"] += 1
            for name in self.leaflets:
                self.dm.leaflet_make_trial(_WHICH[name], step, along_direction)
                self.dm.leaflet_swap_trial(_WHICH[name])
            e1 = self._energy(False)
            if e1 <= e0:
                return True, e1, step
            for name in self.leaflets:
                self.dm.leaflet_swap_trial(_WHICH[name])
            step *= 0.5
            if step < 1e-16:
                break
        return False, e0, step

    def _relax_cg(self, *, max_iters, step_size, tol, preconditioner, k_smooth, area_kept_only, gd_fallback) -> dict:
        dm = self.dm
        st = dict(accepted_steps=0, rejected_steps=0, backtracking_steps=0, stop_reason="completed_max_iters",
                  initial_energy=0.0, final_energy=0.0, initial_gradient_norm=0.0, final_gradient_norm=0.0,
                  cg_fallback_accepted_count=0)
        self.stats = st
        if step_size <= 0.0:
            st["stop_reason"] = "step_size_zero"
            return st
        dm.update_vertex_normals()
        self._refresh()
        e0, gnorm = self._gradients()
        st["initial_energy"], st["initial_gradient_norm"] = e0, gnorm
        st["final_energy"], st["final_gradient_norm"] = e0, gnorm
        if gnorm == 0.0 or (tol > 0.0 and gnorm < tol):
            st["stop_reason"] = "zero_gradient" if gnorm == 0.0 else "converged"
            return self._finish(st)
        if preconditioner:
            for name in self.leaflets:
                dm.leaflet_build_preconditioner(_WHICH[name], float(k_smooth.get(name, 0.0)),
                                                bool(area_kept_only.get(name, False)))

        def rz() -> float:
            for n in self.leaflets:
                dm.leaflet_rz(_WHICH[n], preconditioner, read=False)
            res = dm.leaflet_results()
            return float(sum(res[_WHICH[n], 4] for n in self.leaflets))

        rz_old = rz()
        for name in self.leaflets:
            dm.leaflet_cg_direction(_WHICH[name], 0.0, True, preconditioner)
        for _ in range(int(max_iters)):
            if gnorm == 0.0:
                st["stop_reason"] = "zero_gradient"
                break
            if tol > 0.0 and gnorm < tol:
                st["stop_reason"] = "converged"
                break
            accepted, e1, step = self._line_search(e0, step_size, True, st)
            fallback = False
            if not accepted and gd_fallback:
                accepted, e1, step = self._line_search(e0, step_size, False, st)
                fallback = accepted
            if not accepted:
                st["rejected_steps"] += 1
                st["stop_reason"] = "line_search_rejected"
                st["final_energy"], st["final_gradient_norm"] = e0, gnorm
                break
            st["accepted_steps"] += 1
            st["cg_fallback_accepted_count"] += int(fallback)
            st["step_size_last_accepted"] = step
            self._after_accepted_step(st["accepted_steps"])
            e0, gnorm = self._gradients()
            st["final_energy"], st["final_gradient_norm"] = e0, gnorm
            if gnorm == 0.0 or (tol > 0.0 and gnorm < tol):
                st["stop_reason"] = "zero_gradient" if gnorm == 0.0 else "converged"
                break
            rz_new = rz()
            if fallback:
                for name in self.leaflets:
                    dm.leaflet_cg_direction(_WHICH[name], 0.0, True, preconditioner)
                rz_old = rz_new
                continue
            if rz_old == 0.0:
                st["stop_reason"] = "cg_breakdown"
                break
            beta = rz_new / rz_old
            for name in self.leaflets:
                dm.leaflet_cg_direction(_WHICH[name], beta, False, preconditioner)
            rz_old = rz_new
        return self._finish(st)


def _flag_mask(mesh, attr: str):
    import numpy as np

    given = getattr(mesh, f"{attr}_mask", None)          # ArrayMesh: ready-made boolean rows
    if given is not None:
        return np.asarray(given() if callable(given) else given, dtype=bool)
    vertices = mesh.vertices
    return np.fromiter((bool(getattr(vertices[int(v)], attr, False)) for v in mesh.vertex_ids), dtype=bool,
                       count=len(mesh.vertex_ids))


def relax_leaflet_tilts(mesh, global_params, param_resolver=None, *, constraint_manager=None, positions=None,
                        mode: str = "nested") -> dict:
    """Mesh-level entry: the call of ``TiltRelaxationManager.relax_leaflet_tilts``
    (``runtime/steppers/tilt_relaxation.py:426-1500``) on the device, for the module set ``tilt_in/out`` +
    ``bending_tilt_in/out``.  Reads the same global parameters (``tilt_step_size``, ``tilt_tol``, ``tilt_inner_steps``
    / ``tilt_coupled_steps``, ``tilt_solver``, ``tilt_cg_max_iters``, ``tilt_cg_preconditioner``,
    ``tilt_cg_rejection_fallback``, ``tilt_projection_cadence`` / ``_interval``), takes the fixed rows from the
    ``tilt_fixed_in`` / ``tilt_fixed_out`` vertex flags, runs the constraint manager's tilt hooks when it has any, and
    writes the relaxed fields back to the mesh.  Returns the relaxation statistics."""
    import numpy as np

    from ..modules.energy import _common as C
    from ..modules.energy import _leaflet as LF

    stats = {"stop_reason": "not_run", "accepted_steps": 0}
    mode_norm = str(mode or "").strip().lower()
    if mode_norm not in ("nested", "coupled"):
        stats["stop_reason"] = "mode_unknown"
        return stats
    step_size = float(global_params.get("tilt_step_size", 0.0) or 0.0)
    if step_size <= 0.0:
        stats["stop_reason"] = "step_size_zero"
        return stats
    tol = max(0.0, float(global_params.get("tilt_tol", 0.0) or 0.0))
    n_inner = int(global_params.get("tilt_inner_steps", 0) or 0)
    if mode_norm == "coupled":
        n_inner = int(global_params.get("tilt_coupled_steps", global_params.get("tilt_inner_steps", 0)) or 0)
    if n_inner <= 0:
        return stats
    solver = str(global_params.get("tilt_solver", "cg") or "cg").strip().lower()
    solver = solver if solver in ("gd", "cg") else "gd"
    max_iters = n_inner
    if solver == "cg":
        max_iters = int(global_params.get("tilt_cg_max_iters", n_inner) or 0)
        if max_iters <= 0:
            stats["stop_reason"] = "max_iters_zero"
            return stats
    cadence = str(global_params.get("tilt_projection_cadence", "per_step") or "per_step").strip().lower()
    if cadence not in ("per_step", "per_pass"):
        raise ValueError("tilt_projection_cadence must be 'per_step' or 'per_pass'.")
    interval = int(global_params.get("tilt_projection_interval", 1) or 1)
    if interval < 1:
        raise ValueError("tilt_projection_interval must be >= 1.")
    fallback = str(global_params.get("tilt_cg_rejection_fallback", "off") or "off").strip().lower()
    if fallback not in ("off", "gd"):
        raise ValueError("tilt_cg_rejection_fallback must be 'off' or 'gd'.")
    precond = str(global_params.get("tilt_cg_preconditioner", "jacobi") or "jacobi").strip().lower()

    fixed = {"in": _flag_mask(mesh, "tilt_fixed_in"), "out": _flag_mask(mesh, "tilt_fixed_out")}
    if not (np.any(~fixed["in"]) or np.any(~fixed["out"])):
        stats["stop_reason"] = "no_free_rows"
        return stats
    pos = C.positions_array(mesh.positions_view() if positions is None else positions)
    state = C.get_state(mesh, pos)
    dm = state.dm
    specs = {}
    for leaf in ("in", "out"):
        LF.refuse_unsupported(global_params, leaf, bending_tilt=True)
        specs[leaf] = LF.selection(mesh, global_params, param_resolver, leaf)
        state.set_leaflet(leaf, L.MOD_TILT | L.MOD_BENDING_TILT, specs[leaf], LF.SIGN[leaf])
        dm.set_leaflet_fixed(_WHICH[leaf], fixed[leaf].astype(np.uint8))
    dm.set_positions(pos)
    dm.upload(_TILTS["in"], np.ascontiguousarray(mesh.tilts_in_view(), dtype=np.float64))
    dm.upload(_TILTS["out"], np.ascontiguousarray(mesh.tilts_out_view(), dtype=np.float64))
    hooks = (None, None)
    if constraint_manager is not None and getattr(constraint_manager, "modules", None):
        hooks = reference_constraint_hooks(mesh, global_params, constraint_manager, pos)
    relaxer = DeviceTiltRelaxer(dm, gradient_hook=hooks[0], refresh_hook=hooks[1], projection_interval=interval,
                                projection_cadence=cadence)

    def kept_only(leaf):
        keep = specs[leaf].get("keep_tilt")
        return keep is not None and not bool(np.all(keep))

    stats = relaxer.relax(max_iters=max_iters, step_size=step_size, tol=tol, solver=solver,
                          preconditioner=precond not in ("none", "off", "false"), gd_fallback=fallback == "gd",
                          k_smooth={leaf: float(global_params.get(f"bending_modulus_{leaf}")
                                                or global_params.get("bending_modulus") or 0.0) for leaf in ("in", "out")},
                          area_kept_only={leaf: kept_only(leaf) for leaf in ("in", "out")})
    mesh.set_tilts_in_from_array(dm.download(_TILTS["in"]))
    mesh.set_tilts_out_from_array(dm.download(_TILTS["out"]))
    stats = dict(stats, solver=solver, max_iters=max_iters, mode=mode_norm, hook_calls=dict(relaxer.hook_calls))
    return stats
