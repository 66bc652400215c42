#!/usr/bin/env python
"""Race / bounds evidence for the patch kernels without compute-sanitizer (closed on the GPU pool).

Runs with the self-check build (``make -C membrane_solver_b200/csrc checked``): every read-modify-write of an owned
accumulator row takes a per-row lock with atomicCAS, every patch-local index is bounds-checked, the epilogue checks
that all locks are free (csrc/ms_kernels.cu, MS_SELF_CHECK).  Prints one JSON line per case; exit code 1 on any
violation.  ``--inject`` runs the negative control (a deliberately repeated record): the checker must fire.

    MS_B200_LIB=$PWD/membrane_solver_b200/libms_b200_checked.so python tools/self_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MS_B200_LIB", os.path.join(ROOT, "membrane_solver_b200", "libms_b200_checked.so"))
if "--inject" in sys.argv:
    os.environ["MS_SELF_CHECK_INJECT"] = "1"
from membrane_solver_b200 import _lib as L  # noqa: E402
from membrane_solver_b200.context import DeviceMesh  # noqa: E402
from membrane_solver_b200.synthetic import icosphere, open_sheet  # noqa: E402


def main():
    inject = "--inject" in sys.argv
    bad = 0
    mods = L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME
    cases = [(150, dict(), 20), (150, dict(threads=32, max_owned=64, max_local=200), 10),
             (150, dict(threads=160, max_owned=512, max_local=896), 10), (150, dict(threads=64, max_owned=256, max_local=512), 10),
             (708, dict(), 5)]
    if inject:
        cases = cases[:1]
    for n, pack, reps in cases:
        pos, tri = icosphere(n)
        nv, nf = pos.shape[0], tri.shape[0]
        dm = DeviceMesh(0, **pack)
        dm.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8))
        dm.set_surface_tension(1.0)
        dm.set_bending_params(1.0, 0.1)
        dm.set_positions(pos)
        for _ in range(reps):
            dm.eval_async(dm.options(mods, constraint_mode=0))
            dm.eval_async(dm.options(mods, want_grad=False))
            dm.eval_async(dm.options(L.MOD_SURFACE | L.MOD_VOLUME))
        dm.sync()
        c = dm.self_check()
        print(json.dumps({"mesh": f"icosphere({n})", "facets": int(nf), "pack": pack, "evaluations": 3 * reps,
                          "rows_locked_twice": c[0], "index_out_of_range": c[1], "lock_held_at_epilogue": c[2],
                          "negative_control": inject}))
        bad += sum(c)
        dm.close()
    if not inject:
        pos, tri = open_sheet(60, 40, jitter=0.05)
        from membrane_solver_b200.geometry.array_mesh import ArrayMesh

        nv, nf = pos.shape[0], tri.shape[0]
        boundary = np.zeros(nv, np.uint8)
        boundary[list(ArrayMesh(pos, tri).boundary_vertex_ids)] = 1
        rng = np.random.default_rng(3)
        dm = DeviceMesh(0)
        dm.set_topology(nv, tri, is_boundary=boundary, order_hint=pos)
        dm.set_surface_tension(rng.uniform(0.5, 1.5, size=nf))
        dm.set_bending_params(rng.uniform(0.5, 1.5, size=nv), 0.1)
        dm.set_positions(pos)
        for _ in range(5):  # (the tilt-magnitude variant needs every byte of shared memory: no room for the lock words)
            dm.eval_async(dm.options(L.MOD_SURFACE | L.MOD_BENDING, diagnostics=True))
            dm.eval_async(dm.options(L.MOD_BENDING, flags=L.FLAG_WILLMORE))
        dm.sync()
        c = dm.self_check()
        print(json.dumps({"mesh": "open sheet 60x40, boundary flags, per-entity parameters", "facets": int(nf), "evaluations": 10,
                          "rows_locked_twice": c[0], "index_out_of_range": c[1], "lock_held_at_epilogue": c[2],
                          "negative_control": False}))
        bad += sum(c)
        dm.close()
    if inject:
        print("negative control:", "the checker fired" if bad > 0 else "THE CHECKER DID NOT FIRE")
        return 0 if bad > 0 else 1
    print("self-check:", "clean" if bad == 0 else f"{bad} violations")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
