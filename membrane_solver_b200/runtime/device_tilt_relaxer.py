"""Device-resident leaflet tilt relaxation at frozen geometry.

Twin of the gradient-descent solver of ``TiltRelaxationManager.relax_leaflet_tilts``
(``runtime/steppers/tilt_relaxation.py:426-1057``: setup ``:630-668``, gradients ``:825-872``, loop ``:894-1055``)
and of its preconditioned conjugate-gradient solver (``:1057-1440``, Jacobi diagonal of
``runtime/preconditioners.py:64-146``, optional gradient-descent fallback) for a mesh whose tilt fields are
constrained only by fixed rows (no tilt constraint modules, no axisymmetric projection).  Positions, both tilt fields, their gradients and the trial fields stay on the device; per
iteration the host sees three scalars (energy, gradient norm, trial energy) and keeps the loop control:

    E0, g = tilt-only evaluation of the leaflet modules;  g[fixed] = 0;  stop on |g| == 0 or |g| < tol
    step = tilt_step_size;  up to 12 trials:  t' = P(t - step g), fixed rows kept;  accept if E(t') <= E0 else halve
    after an accepted step the fields pass through the tangent projection once more (the reference's per-step
    refresh writes them to the mesh and reads them back projected, ``:803-823``)
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

from .. import _lib as L

_WHICH = {"in": L.LEAFLET_IN, "out": L.LEAFLET_OUT}


@dataclass
class DeviceTiltRelaxer:
    """``dm``: a DeviceMesh with positions uploaded, both leaflets described (``set_leaflet``), their fixed rows
    (``set_leaflet_fixed``) and tilt fields (``ARR_TILTS_IN`` / ``_OUT``) in place."""

    dm: object
    leaflets: tuple = ("in", "out")
    modules: int = L.MOD_TILT | L.MOD_BENDING_TILT
    stats: dict = field(default_factory=dict)

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

    def _energy_and_gradient_norm(self) -> tuple[float, float]:
        self._launch(True)
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
        for name in self.leaflets:
            dm.leaflet_project_tilts(_WHICH[name])
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
            for name in self.leaflets:
                dm.leaflet_project_tilts(_WHICH[name])
            st["final_energy"] = e1
        return st

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
        for name in self.leaflets:
            dm.leaflet_project_tilts(_WHICH[name])
        e0, gnorm = self._gradients()
        st["initial_energy"], st["initial_gradient_norm"] = e0, gnorm
        st["final_energy"], st["final_gradient_norm"] = e0, gnorm
        if gnorm == 0.0 or (tol > 0.0 and gnorm < tol):
            st["stop_reason"] = "zero_gradient" if gnorm == 0.0 else "converged"
            return st
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
            for name in self.leaflets:
                dm.leaflet_project_tilts(_WHICH[name])
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
        return st
