#!/usr/bin/env python
"""BASELINE.json configs[2] on the device: the cube of ``meshes/bending_cube.yaml`` (Helfrich bending,
kappa = 1, hard volume constraint V0 = 1), refined k times (24 * 4^k facets; k = 8 -> 1 572 864), minimised
with the device-resident gradient descent.  The reference's ``benchmarks/benchmark_bending.py`` reports the
average time of a minimiser step; so does this.

    python tools/config3_bending_cube.py [levels=8] [steps=20]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from membrane_solver_b200 import _lib as L
from membrane_solver_b200.context import DeviceMesh
from membrane_solver_b200.geometry.refine import cube_mesh, refine_triangles
from membrane_solver_b200.runtime.device_minimizer import DeviceMinimizer

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t0 = time.perf_counter()
pos, tri = cube_mesh()
for _ in range(levels):
    pos, tri, _, _ = refine_triangles(pos, tri)
t_refine = time.perf_counter() - t0
nv, nf = pos.shape[0], tri.shape[0]
dm = DeviceMesh(0)
t0 = time.perf_counter()
dm.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8), order_hint=pos)   # refinement order -> Morton order inside
t_pack = time.perf_counter() - t0
dm.set_surface_tension(0.0)
dm.set_bending_params(1.0, 0.0)
dm.set_positions(pos)
mini = DeviceMinimizer(dm, L.MOD_BENDING, volume_mode="lagrange", v_target=1.0, enforce_volume=True,
                       projection_during_minimization=True)
e0 = mini.energy()
mini.minimize(n_steps=2)  # warm-up (kernel load, array allocation)
dm.sync()
t0 = time.perf_counter()
res = mini.minimize(n_steps=steps)
dm.sync()
dt = time.perf_counter() - t0
info = dm.pack_info()
vol = dm.eval(dm.options(L.MOD_VOLUME, want_grad=False)).volume
print(f"bending_cube r{levels}: {nf} facets / {nv} vertices; refine {t_refine:.2f} s, pack {t_pack:.2f} s "
      f"(listed {info['n_listed'] / nf:.3f} x nf)")
print(f"  energy {e0:.6f} -> {res['energy']:.6f} after {steps + 2} GD steps, volume {vol:.12f}, "
      f"accepted {sum(h[3] for h in mini.history)}/{len(mini.history)}")
print(f"  average minimiser step: {dt / steps * 1e3:.3f} ms")
dm.close()
