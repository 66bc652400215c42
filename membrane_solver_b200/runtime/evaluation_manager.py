"""Energy / gradient evaluation orchestration on the B200 path.

Twin of ``runtime/evaluation_manager.py`` (entry points ``:134-225``) with the same constructor
and method names.  Modules that carry ``B200_MODULE`` are evaluated TOGETHER in one fused device
pass (one upload of the positions, pass A + pass B, one download of the gradient); any other
module in the list is called through the reference's array contract with the same
signature-adapting dispatch (``evaluation_manager.py:45-124``).
"""

from __future__ import annotations

import inspect
from typing import Any, Callable

import numpy as np

from .. import _lib as L
from .device_state import get_state, positions_array


class EvaluationManager:
    def __init__(self, *, mesh, global_params, param_resolver, energy_modules: list[Any],
                 energy_module_names: list[str], energy_context_fn: Callable | None = None,
                 experimental_energy_scale_fn: Callable[[str], float] | None = None):
        self.mesh = mesh
        self.global_params = global_params
        self.param_resolver = param_resolver
        self.energy_modules = energy_modules
        self.energy_module_names = energy_module_names
        self.energy_context_fn = energy_context_fn
        self.experimental_energy_scale_fn = experimental_energy_scale_fn or (lambda name: 1.0)
        self._specs: dict[int, dict] = {}

    # -- reference-style dispatch for modules without a B200 twin -----------------
    def _call_fn(self, fn: Callable, **kwargs) -> Any:
        key = id(getattr(fn, "__func__", fn))
        spec = self._specs.get(key)
        if spec is None:
            params = inspect.signature(fn).parameters
            var_kw = any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())
            spec = {"names": list(params), "var_kw": var_kw}
            self._specs[key] = spec
        names = spec["names"]
        call = dict(kwargs)
        if "resolver" in names:
            call.setdefault("resolver", self.param_resolver)
        elif "param_resolver" in names or spec["var_kw"]:
            call.setdefault("param_resolver", self.param_resolver)
        if ("ctx" in names or spec["var_kw"]) and self.energy_context_fn is not None:
            call.setdefault("ctx", self.energy_context_fn())
        if not spec["var_kw"]:
            call = {k: v for k, v in call.items() if k in names}
        args = []
        if names and names[0] not in call:
            args.append(self.mesh)
        if len(names) > 1 and names[1] not in call:
            args.append(self.global_params)
        return fn(*args, **call)

    @staticmethod
    def _coerce(value) -> float:
        arr = np.asarray(value, dtype=float)
        return float(arr) if arr.ndim == 0 else float(np.sum(arr))

    # -- the fused device pass -----------------------------------------------------
    def _split(self):
        fused, other = [], []
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            (fused if hasattr(mod, "B200_MODULE") else other).append((name, mod))
        return fused, other

    def _fused_eval(self, positions, fused, *, want_grad: bool, grad=None, project: bool = False):
        """Evaluate the B200 modules in one pass.  Returns ({name: energy}, EvalResult) or None when
        the module set needs the sequential route (approx-mode boundary zeroing, several bodies)."""
        pos = positions_array(positions)
        st = get_state(self.mesh, pos)
        mask, flags, extra = 0, 0, {}
        bits = 0
        for _, mod in fused:
            bits |= mod.B200_MODULE
        if (bits & L.MOD_BENDING) and (bits & L.MOD_BENDING_TILT):
            return None  # both write the seed array: evaluate them one after the other
        for name, mod in fused:
            if name == "volume" and L.MOD_VOLUME & mask:
                continue
            if name == "volume" and self.global_params.get("volume_constraint_mode", "lagrange") != "penalty":
                continue  # lagrange mode: the volume enters through the constraint projection
            if name == "tilt" and float(self.param_resolver.get(None, "tilt_rigidity") or 0.0) == 0.0:
                continue
            cfg = mod.b200_configure(st, self.mesh, self.global_params, self.param_resolver)
            if cfg.get("unfused"):
                return None
            mask |= mod.B200_MODULE
            flags |= cfg.pop("flags", 0)
            extra.update(cfg)
        if (flags & L.FLAG_APPROX) and st.boundary is not None:
            return None
        constraint_mode = extra.get("constraint_mode", -1)
        if project and constraint_mode < 0 and st.body_rows is not None:
            if self.global_params.get("volume_constraint_mode", "lagrange") == "lagrange":
                mask |= L.MOD_VOLUME
                constraint_mode = 0
        opts = st.dm.options(mask, flags=flags, want_grad=want_grad, constraint_mode=constraint_mode,
                             k_vol=extra.get("k_vol", 0.0), v_target=extra.get("v_target", 0.0),
                             apply_fixed=project)
        res = st.dm.eval_host(opts, pos, grad=grad)
        energies = {}
        for name, mod in fused:
            if not (mod.B200_MODULE & mask) or (name == "volume" and constraint_mode != 1):
                energies[name] = 0.0
            elif name == "volume":
                energies[name] = mod.b200_energy(res, extra.get("k_vol", 0.0), extra.get("v_target", 0.0))
            else:
                energies[name] = mod.b200_energy(res)
        return energies, res

    # -- entry points (same names as the reference) --------------------------------
    def compute_energy_and_gradient_array(self, *, positions):
        """Raw module energy and dense shape gradient (``evaluation_manager.py:134-151``)."""
        index_map = self.mesh.vertex_index_to_row
        fused, other = self._split()
        grad = np.zeros_like(np.asarray(positions, dtype=np.float64))
        total = 0.0
        done = None
        if fused:
            done = self._fused_eval(positions, fused, want_grad=True, grad=grad)
        if done is not None:
            total += sum(done[0].values())
        else:
            other = fused + other
        for _, mod in other:
            total += self._call_fn(mod.compute_energy_and_gradient_array, positions=positions, index_map=index_map,
                                   grad_arr=grad)
        return float(total), grad

    def compute_energy_and_projected_gradient(self, *, positions):
        """Energy and the gradient after the single-constraint KKT projection and the fixed-vertex
        mask, all on the device (``minimizer.py:941-992`` + ``constraint_manager.py:294-301``)."""
        fused, other = self._split()
        if other:
            raise L.B200Error("compute_energy_and_projected_gradient needs every loaded module on the B200 path")
        grad = np.zeros_like(np.asarray(positions, dtype=np.float64))
        done = self._fused_eval(positions, fused, want_grad=True, grad=grad, project=True)
        if done is None:
            raise L.B200Error("this module set cannot be evaluated in one fused pass")
        return float(sum(done[0].values())), grad, done[1]

    def compute_energy_breakdown(self, *, positions) -> dict[str, float]:
        """Per-module energies (``evaluation_manager.py:153-182``)."""
        index_map = self.mesh.vertex_index_to_row
        fused, other = self._split()
        out: dict[str, float] = {}
        done = self._fused_eval(positions, fused, want_grad=False) if fused else None
        if done is None:
            other = fused + other
        else:
            out.update(done[0])
        for name, mod in other:
            dummy = np.zeros_like(np.asarray(positions, dtype=np.float64))
            out[name] = float(self._call_fn(mod.compute_energy_and_gradient_array, positions=positions,
                                            index_map=index_map, grad_arr=dummy))
        return {name: float(self.experimental_energy_scale_fn(str(name))) * out[name]
                for name in self.energy_module_names}

    def compute_total_energy_array_with_tilts(self, *, positions, tilts) -> float:
        """Total energy for fixed positions and a given tilt field (``evaluation_manager.py:227-301``)."""
        index_map = self.mesh.vertex_index_to_row
        total = 0.0
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            scale = float(self.experimental_energy_scale_fn(str(name)))
            kwargs = {"positions": positions, "index_map": index_map}
            if getattr(mod, "USES_TILT", False):
                kwargs["tilts"] = tilts
            if hasattr(mod, "compute_energy_array"):
                total += scale * self._coerce(self._call_fn(mod.compute_energy_array, **kwargs))
            else:
                dummy = np.zeros_like(np.asarray(positions, dtype=np.float64))
                total += scale * self._coerce(self._call_fn(mod.compute_energy_and_gradient_array, grad_arr=dummy,
                                                            **kwargs))
        return float(total)

    def compute_energy_array_with_tilts(self, *, positions, tilts) -> float:
        """Tilt-dependent energy for fixed positions and a given tilt field (``evaluation_manager.py:303-384``, the
        accept / reject energy of the single-field tilt relaxation, ``runtime/minimizer.py:756-768``): only the
        modules with ``USES_TILT`` take part; ``compute_energy_array`` is preferred, then the array gradient API
        into a scratch gradient, then the legacy dict API -- the reference's order."""
        index_map = self.mesh.vertex_index_to_row
        names = self.energy_module_names
        if len(names) != len(self.energy_modules):
            names = [getattr(m, "__name__", m.__class__.__name__) for m in self.energy_modules]
        total = 0.0
        for name, mod in zip(names, self.energy_modules):
            if not getattr(mod, "USES_TILT", False):
                continue
            scale = float(self.experimental_energy_scale_fn(str(name)))
            if hasattr(mod, "compute_energy_array"):
                e = self._call_fn(mod.compute_energy_array, positions=positions, index_map=index_map, tilts=tilts)
            elif hasattr(mod, "compute_energy_and_gradient_array"):
                dummy = None if hasattr(mod, "B200_MODULE") else np.zeros_like(np.asarray(positions, dtype=np.float64))
                e = self._call_fn(mod.compute_energy_and_gradient_array, positions=positions, index_map=index_map,
                                  grad_arr=dummy, tilts=tilts, tilt_grad_arr=None)
            else:
                try:
                    e, _ = mod.compute_energy_and_gradient(self.mesh, self.global_params, self.param_resolver,
                                                           compute_gradient=False)
                except TypeError:
                    e, _ = mod.compute_energy_and_gradient(self.mesh, self.global_params, self.param_resolver)
            total += scale * float(e)
        return float(total)

    def compute_energy_and_tilt_gradient_array(self, *, positions, tilts, tilt_grad_arr) -> float:
        """Tilt-dependent energy and dense tilt gradient (``evaluation_manager.py:386-462``): only the
        modules with ``USES_TILT`` take part; ``tilt_grad_arr`` is overwritten."""
        index_map = self.mesh.vertex_index_to_row
        tilt_grad_arr.fill(0.0)
        total = 0.0
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            if not getattr(mod, "USES_TILT", False):
                continue
            scale = float(self.experimental_energy_scale_fn(str(name)))
            part = np.zeros_like(tilt_grad_arr)
            dummy = None if hasattr(mod, "B200_MODULE") else np.zeros_like(np.asarray(positions, dtype=np.float64))
            e = self._call_fn(mod.compute_energy_and_gradient_array, positions=positions, index_map=index_map,
                              grad_arr=dummy, tilts=tilts, tilt_grad_arr=part)
            tilt_grad_arr += scale * part
            total += scale * float(e)
        return float(total)

    # -- leaflet tilt entry points (evaluation_manager.py:464-742) ---------------------------------
    def _leaflet_pass(self, positions, tilts_in, tilts_out, *, tilt_in_grad_arr=None, tilt_out_grad_arr=None):
        """Evaluate every loaded leaflet twin (``B200_LEAFLET``) on the device at frozen positions: ONE upload
        of the positions and of each tilt field, ONE sweep per leaflet covering both of its modules, tilt
        gradients only (the shape gradient is never needed by these entry points: the reference lets the
        modules write it into a scratch array it discards).  Returns {module name: scaled energy}."""
        from ..modules.energy import _leaflet as LF

        pos = positions_array(positions)
        st = get_state(self.mesh, pos)
        groups: dict[str, list] = {}
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            tag = getattr(mod, "B200_LEAFLET", None)
            if tag is not None:
                groups.setdefault(tag[0], []).append((name, tag[1], float(self.experimental_energy_scale_fn(str(name)))))
        energies: dict[str, float] = {}
        if not groups:
            return energies
        st.dm.set_positions(pos)
        if self._leaflet_pair_pass(st, LF, groups, tilts_in, tilts_out, tilt_in_grad_arr, tilt_out_grad_arr, energies):
            return energies
        for leaflet, members in groups.items():
            tilts = tilts_in if leaflet == "in" else tilts_out
            out = tilt_in_grad_arr if leaflet == "in" else tilt_out_grad_arr
            t = LF._leaflet_tilts(self.mesh, leaflet, tilts)
            st.dm.upload(LF.ARR_TILTS[leaflet], t)
            spec = LF.selection(self.mesh, self.global_params, self.param_resolver, leaflet)
            a, b = spec.get("keep_bt"), spec.get("keep_tilt")
            same_keep = (a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b))
            together = same_keep and len(members) > 1 and all(scale == 1.0 for _, _, scale in members)
            batches = [members] if together else [[m] for m in members]
            for batch in batches:
                bits = 0
                for name, bit, _ in batch:
                    if LF.configure(st, self.mesh, self.global_params, self.param_resolver, leaflet, bit) is None:
                        energies[name] = 0.0
                    else:
                        bits |= bit
                if not bits:
                    continue
                if len(batch) > 1:
                    st.set_leaflet(leaflet, bits, spec, LF.SIGN[leaflet])
                got = st.dm.eval_leaflet(LF.WHICH[leaflet], bits, want_grad=False, want_tilt_grad=out is not None)
                scale = batch[0][2] if len(batch) == 1 else 1.0
                for name, bit, sc in batch:
                    if bit & bits:
                        energies[name] = sc * got[LF.ENERGY_SLOT[bit]]
                if out is not None:
                    out += scale * st.dm.download(LF.ARR_TILT_GRAD[leaflet])
        return energies

    def _leaflet_pair_pass(self, st, LF, groups, tilts_in, tilts_out, grad_in, grad_out, energies) -> bool:
        """Both leaflets carry the same modules with unit scales and one facet selection each: evaluate them with
        ONE call (one cooperative launch on small meshes) and read all energies back with one copy."""
        if set(groups) != {"in", "out"}:
            return False
        bits = {}
        specs = {}
        for leaflet, members in groups.items():
            if any(scale != 1.0 for _, _, scale in members):
                return False
            spec = LF.selection(self.mesh, self.global_params, self.param_resolver, leaflet)
            a, b = spec.get("keep_bt"), spec.get("keep_tilt")
            if not ((a is None and b is None) or (a is not None and b is not None and np.array_equal(a, b))):
                return False
            specs[leaflet] = spec
            bits[leaflet] = 0
            for _, bit, _ in members:
                bits[leaflet] |= bit
        if bits["in"] != bits["out"]:
            return False
        for leaflet, members in groups.items():
            for name, bit, _ in members:          # a module that contributes nothing (zero modulus) breaks the symmetry
                if LF.configure(st, self.mesh, self.global_params, self.param_resolver, leaflet, bit) is None:
                    return False
            st.set_leaflet(leaflet, bits[leaflet], specs[leaflet], LF.SIGN[leaflet])
            tilts = tilts_in if leaflet == "in" else tilts_out
            st.dm.upload(LF.ARR_TILTS[leaflet], LF._leaflet_tilts(self.mesh, leaflet, tilts))
        want_tg = grad_in is not None or grad_out is not None
        st.dm.eval_leaflet_pair(bits["in"], want_grad=False, want_tilt_grad=want_tg)
        res = st.dm.leaflet_results()
        for leaflet, members in groups.items():
            for name, bit, _ in members:
                energies[name] = float(res[LF.WHICH[leaflet], LF.ENERGY_SLOT[bit]])
        if grad_in is not None:
            grad_in += st.dm.download(LF.ARR_TILT_GRAD["in"])
        if grad_out is not None:
            grad_out += st.dm.download(LF.ARR_TILT_GRAD["out"])
        return True

    def _other_leaflet_energy(self, name, mod, *, positions, tilts_in, tilts_out, grad_arr, tilt_in_grad_arr=None,
                              tilt_out_grad_arr=None) -> float:
        """A module without a leaflet twin, through the reference's array contract (keywords the signature does
        not list are dropped by ``_call_fn``, as ``evaluation_manager.py:88-124`` does)."""
        index_map = self.mesh.vertex_index_to_row
        scale = float(self.experimental_energy_scale_fn(str(name)))
        before = None
        if abs(scale - 1.0) > 1.0e-15 and tilt_in_grad_arr is not None:
            before = (tilt_in_grad_arr.copy(), tilt_out_grad_arr.copy())
        e = self._call_fn(mod.compute_energy_and_gradient_array, positions=positions, index_map=index_map,
                          grad_arr=grad_arr, tilts_in=tilts_in, tilts_out=tilts_out,
                          tilt_in_grad_arr=tilt_in_grad_arr, tilt_out_grad_arr=tilt_out_grad_arr)
        if before is not None:
            tilt_in_grad_arr[:] = before[0] + scale * (tilt_in_grad_arr - before[0])
            tilt_out_grad_arr[:] = before[1] + scale * (tilt_out_grad_arr - before[1])
        return scale * self._coerce(e)

    def compute_energy_and_leaflet_tilt_gradients_array(self, *, positions, tilts_in, tilts_out, tilt_in_grad_arr,
                                                        tilt_out_grad_arr, tilt_vertex_areas_in=None,
                                                        tilt_vertex_areas_out=None, grad_dummy=None,
                                                        tilt_only: bool = False) -> float:
        """Total energy and leaflet tilt gradients at frozen positions (``evaluation_manager.py:630-742``).
        ``tilt_vertex_areas_*`` (the reference's closed form for lumped ``tilt_in`` / ``tilt_out``) are accepted
        and not needed: the device evaluates the modules themselves, which gives the same value."""
        _ = (tilt_vertex_areas_in, tilt_vertex_areas_out)
        if grad_dummy is None:
            grad_dummy = np.zeros_like(np.asarray(positions, dtype=np.float64))
        else:
            grad_dummy.fill(0.0)
        tilt_in_grad_arr.fill(0.0)
        tilt_out_grad_arr.fill(0.0)
        done = self._leaflet_pass(positions, tilts_in, tilts_out, tilt_in_grad_arr=tilt_in_grad_arr,
                                  tilt_out_grad_arr=tilt_out_grad_arr)
        total = float(sum(done.values()))
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            if name in done:
                continue
            leafy = getattr(mod, "USES_TILT_LEAFLETS", False)
            total += self._other_leaflet_energy(name, mod, positions=positions, tilts_in=tilts_in, tilts_out=tilts_out,
                                                grad_arr=None if (tilt_only and leafy) else grad_dummy,
                                                tilt_in_grad_arr=tilt_in_grad_arr, tilt_out_grad_arr=tilt_out_grad_arr)
        return float(total)

    def compute_tilt_dependent_energy_with_leaflet_tilts(self, *, positions, tilts_in, tilts_out, grad_dummy=None,
                                                         tilt_vertex_areas_in=None, tilt_vertex_areas_out=None) -> float:
        """Energy of the leaflet-tilt modules only (``evaluation_manager.py:537-628``)."""
        _ = (grad_dummy, tilt_vertex_areas_in, tilt_vertex_areas_out)
        done = self._leaflet_pass(positions, tilts_in, tilts_out)
        total = float(sum(done.values()))
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            if name in done or not getattr(mod, "USES_TILT_LEAFLETS", False):
                continue
            total += self._other_leaflet_energy(name, mod, positions=positions, tilts_in=tilts_in, tilts_out=tilts_out,
                                                grad_arr=None)
        return float(total)

    def compute_energy_array_with_leaflet_tilts(self, *, positions, tilts_in, tilts_out, grad_dummy=None) -> float:
        """Total energy for fixed positions and leaflet tilt arrays (``evaluation_manager.py:464-535``)."""
        done = self._leaflet_pass(positions, tilts_in, tilts_out)
        total = float(sum(done.values()))
        fused = [(n, m) for n, m in zip(self.energy_module_names, self.energy_modules)
                 if hasattr(m, "B200_MODULE")]
        part = self._fused_eval(positions, fused, want_grad=False) if fused else None
        if part is not None:
            for name, e in part[0].items():
                total += float(self.experimental_energy_scale_fn(str(name))) * e
                done[name] = e
        for name, mod in zip(self.energy_module_names, self.energy_modules):
            if name in done:
                continue
            dummy = np.zeros_like(np.asarray(positions, dtype=np.float64))
            total += self._other_leaflet_energy(name, mod, positions=positions, tilts_in=tilts_in, tilts_out=tilts_out,
                                                grad_arr=dummy)
        return float(total)

    def compute_energy_array_total(self, *, positions) -> float:
        """Total energy for fixed positions (``evaluation_manager.py:184-225``)."""
        return float(sum(self.compute_energy_breakdown(positions=positions).values()))
