#!/usr/bin/env python
"""Small, fast exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_target.py

The persistent patch kernels hand shared-memory buffers between a producer warp, consumer groups (token ring of
named barriers) and epilogue warps (mbarriers), and finalise through a last-CTA ticket; the leaflet kernels have a
cooperative single-launch form; the peer-memory halo pulls wait on flags.  Every result is also checked against the
oracle, so a run under the sanitizer is a correctness run as well."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from membrane_solver_b200 import _lib as L  # noqa: E402
from membrane_solver_b200.context import DeviceMesh  # noqa: E402
from membrane_solver_b200.synthetic import icosphere, open_sheet  # noqa: E402
from oracle import ref_modules as ref  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


def closed_meshes():
    for n, pack in ((12, dict()), (12, dict(threads=32, max_owned=24, max_local=150)),
                    (20, dict(threads=64, max_owned=128, max_local=300))):
        pos, tri = icosphere(n)
        nv, nf = pos.shape[0], tri.shape[0]
        want = ref.fused_surface_bending_volume(pos, tri, np.ones(nf), 1.0, 0.05, np.zeros(nv, bool))
        dm = DeviceMesh(0, **pack)
        dm.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8))
        dm.set_surface_tension(1.0)
        dm.set_bending_params(1.0, 0.05)
        dm.set_positions(pos)
        mods = L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME
        for _ in range(3):  # fused finalisation (last-CTA ticket) back to back
            r = dm.eval(dm.options(mods, constraint_mode=0))
        lam = float(np.vdot(want["grad"], want["vol_grad"]) / np.vdot(want["vol_grad"], want["vol_grad"]))
        assert rel(dm.download(L.ARR_GRAD), want["grad"] - lam * want["vol_grad"]) <= 1e-11
        assert abs(r.e_bending - want["E_bending"]) <= 1e-12 * want["E_bending"]
        dm.eval(dm.options(mods, want_grad=False))                  # energy-only: pass A alone finalises
        dm.eval(dm.options(L.MOD_SURFACE | L.MOD_VOLUME))           # pass B alone (no bending)
        grad = np.empty_like(pos)
        dm.eval_host(dm.options(mods, constraint_mode=0), pos, grad=grad)   # host path (unfused reduce / coefficient)
        assert rel(grad, want["grad"] - lam * want["vol_grad"]) <= 1e-11
        dm.direction_from_gradient(-1.0)                            # projection on the fly
        dm.line_search_stats()
        dm.close()
        print("closed", n, pack, "ok")


def open_mesh_with_tilts():
    pos, tri = open_sheet(9, 7, jitter=0.08)
    nv, nf = pos.shape[0], tri.shape[0]
    from membrane_solver_b200.geometry.array_mesh import ArrayMesh

    boundary = np.zeros(nv, np.uint8)
    boundary[list(ArrayMesh(pos, tri).boundary_vertex_ids)] = 1
    rng = np.random.default_rng(3)
    tilts = 0.1 * rng.normal(size=pos.shape)
    dm = DeviceMesh(0)
    dm.set_topology(nv, tri, is_boundary=boundary, order_hint=pos)
    dm.set_surface_tension(rng.uniform(0.5, 1.5, size=nf))
    dm.set_bending_params(rng.uniform(0.5, 1.5, size=nv), 0.1)
    dm.set_tilt_rigidity(2.0)
    dm.set_positions(pos)
    dm.set_tilts(tilts)
    dm.eval(dm.options(L.MOD_SURFACE | L.MOD_BENDING | L.MOD_TILT, diagnostics=True))
    dm.eval(dm.options(L.MOD_BENDING, flags=L.FLAG_WILLMORE))
    dm.eval(dm.options(L.MOD_BENDING_TILT | L.MOD_TILT))
    dm.close()
    print("open sheet with tilts ok")


def main():
    if L.device_count() < 1:
        raise SystemExit("needs a CUDA device")
    closed_meshes()
    open_mesh_with_tilts()
    print("sanitize_target: all kernels exercised")


if __name__ == "__main__":
    main()
