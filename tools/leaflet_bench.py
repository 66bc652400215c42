"""Leaflet tilt modules (BASELINE configs[3] module set: bending_tilt_in/out + tilt_in/out): device timings.

1. the caveolin free-disk fixture (2 208 facets, tests/golden/leaflet.npz) -- the size the reference runs:
   per-call latency resident on the device, through the plugin API with host arrays, and the CPU port (oracle);
2. the same module set on the 10 M-facet icosphere (throughput of the generic per-facet sweeps).
Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from membrane_solver_b200 import _lib as L  # noqa: E402
from membrane_solver_b200.context import DeviceMesh  # noqa: E402
from membrane_solver_b200.synthetic import icosphere  # noqa: E402


def out(**kw):
    print(json.dumps(kw), flush=True)


def small():
    from oracle import ref_leaflet as rl

    g = np.load(os.path.join(ROOT, "tests", "golden", "leaflet.npz"))
    pos, tri, isb = g["r1_pos"], g["r1_tri"], g["r1_is_boundary"].astype(np.uint8)
    nv, nf = pos.shape[0], tri.shape[0]
    dm = DeviceMesh(0)
    dm.set_topology(nv, tri, is_boundary=isb)
    dm.set_positions(pos)
    leaf = {}
    for name, which, arr in (("in", L.LEAFLET_IN, L.ARR_TILTS_IN), ("out", L.LEAFLET_OUT, L.ARR_TILTS_OUT)):
        pre = f"r1_{name}_"
        leaf[name] = dict(tilts=g[pre + "tilts"], keep=g[pre + "keep"], interior=g[pre + "interior"],
                          base_zero=g[pre + "base_zero"], kappa=g[pre + "kappa"], c0=g[pre + "c0"],
                          sign=float(g[pre + "sign"]), k_tilt=float(g[pre + "k_tilt"]))
        d = leaf[name]
        dm.set_leaflet(which, div_sign=d["sign"], kappa=1.0, c0=0.0, k_tilt=d["k_tilt"], facet_keep=d["keep"],
                       interior=d["interior"], base_zero=d["base_zero"])
        dm.upload(arr, d["tilts"])
    both = L.MOD_TILT | L.MOD_BENDING_TILT
    for label, kw in (("tilt_only", dict(want_grad=False)), ("full", dict(want_grad=True))):
        for _ in range(20):
            dm.eval_leaflet(L.LEAFLET_IN, both, **kw)
            dm.eval_leaflet(L.LEAFLET_OUT, both, **kw)
        n = 300
        t0 = time.perf_counter()
        for _ in range(n):
            dm.eval_leaflet(L.LEAFLET_IN, both, **kw)       # returns the two energies: one sync per call
            dm.eval_leaflet(L.LEAFLET_OUT, both, **kw)
        dt = (time.perf_counter() - t0) / n
        # the CPU port of the same four modules
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            for name in ("in", "out"):
                d = leaf[name]
                gr = np.zeros_like(pos) if kw["want_grad"] else None
                tg = np.zeros_like(pos)
                rl.leaflet_bending_tilt_energy_and_gradient(pos, tri, d["tilts"], d["kappa"], d["c0"], sign=d["sign"],
                                                            keep=d["keep"], interior=d["interior"],
                                                            base_zero=d["base_zero"], is_boundary=isb.astype(bool),
                                                            grad=gr, tilt_grad=tg)
                rl.leaflet_tilt_energy_and_gradient(pos, tri, d["tilts"], d["k_tilt"], keep=d["keep"], grad=gr, tilt_grad=tg)
        cpu = (time.perf_counter() - t0) / reps
        out(case="caveolin_r1", facets=nf, evaluation=label, modules="bending_tilt_in+out, tilt_in+out",
            device_ms=dt * 1e3, cpu_port_ms=cpu * 1e3, note="device: resident inputs, energies read back every call")
    dm.close()


def large():
    pos, tri = icosphere(708)
    nv, nf = pos.shape[0], tri.shape[0]
    rng = np.random.default_rng(2)
    t = 0.1 * rng.standard_normal((nv, 3))
    t -= (t * pos).sum(axis=1, keepdims=True) * pos / (pos * pos).sum(axis=1, keepdims=True)
    dm = DeviceMesh(0)
    dm.set_topology(nv, tri)
    dm.set_positions(pos)
    dm.set_leaflet(L.LEAFLET_IN, div_sign=-1.0, kappa=1.0, c0=0.0, k_tilt=225.0)
    dm.upload(L.ARR_TILTS_IN, t)
    both = L.MOD_TILT | L.MOD_BENDING_TILT
    lib, h = dm._lib, dm._h
    for label, mods, wg in (("bt+tilt tilt_only", both, 0), ("bt+tilt full", both, 1), ("tilt full", L.MOD_TILT, 1)):
        for _ in range(2):
            L.check(lib.ms_ctx_eval_leaflet(h, L.LEAFLET_IN, mods, wg, 1, 0, 0, None))
        dm.sync()
        dm.timer_start()
        n = 5
        for _ in range(n):
            L.check(lib.ms_ctx_eval_leaflet(h, L.LEAFLET_IN, mods, wg, 1, 0, 0, None))
        ms = dm.timer_stop() / n
        out(case="icosphere_708", facets=nf, evaluation=label, device_ms=ms, gfacet_evals_per_s=nf / ms / 1e6)
    dm.close()


def relax():
    """Row f3: the tilt relaxations of tests/golden/tilt_relaxation.npz (gd5: 5 accepted GD steps, 25 halvings; cg6: 6
    Jacobi-preconditioned CG steps) resident on the device vs. the CPU port of the same loops."""
    from membrane_solver_b200.runtime.device_tilt_relaxer import DeviceTiltRelaxer
    from oracle import ref_leaflet as rl

    g = np.load(os.path.join(ROOT, "tests", "golden", "tilt_relaxation.npz"))
    for case in ("gd5", "cg6"):
        p = case + "_"
        pos, tri = g[p + "pos"], g[p + "tri"]
        leaflets = {k: dict(keep=g[p + f"{k}_keep"], interior=g[p + f"{k}_interior"], base_zero=g[p + f"{k}_base_zero"],
                            kappa=g[p + f"{k}_kappa"], c0=g[p + f"{k}_c0"], k_tilt=float(g[p + f"{k}_k_tilt"]))
                    for k in ("in", "out")}
        dm = DeviceMesh(0)
        dm.set_topology(pos.shape[0], tri, is_boundary=g[p + "is_boundary"].astype(np.uint8))
        dm.set_positions(pos)
        for leaf, d in leaflets.items():
            which = L.LEAFLET_IN if leaf == "in" else L.LEAFLET_OUT
            dm.set_leaflet(which, div_sign=-1.0 if leaf == "in" else 1.0, kappa=1.0, c0=0.0, k_tilt=d["k_tilt"],
                           facet_keep=d["keep"].astype(np.uint8), interior=d["interior"].astype(np.uint8),
                           base_zero=d["base_zero"].astype(np.uint8))
            dm.set_leaflet_fixed(which, g[p + f"fixed_{leaf}"].astype(np.uint8))
        solver = str(g[p + "solver"])
        ks = {"in": float(g[p + "k_smooth_in"]), "out": float(g[p + "k_smooth_out"])}
        kw = dict(max_iters=int(g[p + "steps"]), step_size=float(g[p + "step_size"]))
        times = []
        for rep in range(6):
            dm.upload(L.ARR_TILTS_IN, g[p + "tilts_in0"])
            dm.upload(L.ARR_TILTS_OUT, g[p + "tilts_out0"])
            t0 = time.perf_counter()
            st = DeviceTiltRelaxer(dm).relax(solver=solver, k_smooth=ks, area_kept_only={"out": True}, **kw)
            times.append(time.perf_counter() - t0)
        dev = min(times[1:])
        start = {"in": g[p + "tilts_in0"], "out": g[p + "tilts_out0"]}
        fixed = {"in": g[p + "fixed_in"], "out": g[p + "fixed_out"]}
        t0 = time.perf_counter()
        if solver == "cg":
            _, st_cpu = rl.relax_leaflet_tilts_cg(pos, tri, start, leaflets, is_boundary=g[p + "is_boundary"], fixed=fixed,
                                                  k_smooth=ks, **kw)
        else:
            _, st_cpu = rl.relax_leaflet_tilts_gd(pos, tri, start, leaflets, is_boundary=g[p + "is_boundary"], fixed=fixed,
                                                  **kw)
        cpu = time.perf_counter() - t0
        out(case=f"caveolin_r1 tilt relaxation {case}", facets=int(tri.shape[0]), accepted=st["accepted_steps"],
            backtracking=st["backtracking_steps"], final_energy=st["final_energy"], device_ms=dev * 1e3,
            cpu_port_ms=cpu * 1e3, cpu_final_energy=st_cpu["final_energy"])
        dm.close()


if __name__ == "__main__":
    small()
    relax()
    large()
