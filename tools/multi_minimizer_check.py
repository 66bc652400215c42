"""torchrun -n N tools/multi_minimizer_check.py [facets] [steps]: the device-resident minimiser over a mesh
partitioned across N GPUs (halo exchange + all-reduced line-search scalars, runtime/partitioned_minimizer.py)
against the same minimiser on ONE GPU holding the whole mesh (rank 0): energies per step, accepted step sizes and
final positions must agree (1e-9, BASELINE.json north_star trajectories), for gradient descent and for the
Polak-Ribiere stepper, with the lagrange volume constraint (KKT-projected gradient) and with the hard Newton volume
projection inside the line search."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from membrane_solver_b200 import _lib as L  # noqa: E402
from membrane_solver_b200.context import DeviceMesh  # noqa: E402
from membrane_solver_b200.partition import PartitionedMesh, split_mesh  # noqa: E402
from membrane_solver_b200.runtime.device_minimizer import DeviceMinimizer  # noqa: E402
from membrane_solver_b200.runtime.partitioned_minimizer import partitioned_minimizer  # noqa: E402
from membrane_solver_b200.synthetic import frequency_for_facets, icosphere  # noqa: E402

CASES = {
    "gd_lagrange": dict(stepper="gd", volume_mode="lagrange", enforce_volume=False),
    "cg_lagrange": dict(stepper="cg", volume_mode="lagrange", enforce_volume=False),
    "gd_hard_volume": dict(stepper="gd", volume_mode="lagrange", enforce_volume=True),
}


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    facets = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pos, tri = icosphere(frequency_for_facets(facets))
    rng = np.random.default_rng(4)
    edge = float(np.linalg.norm(pos[tri[:, 0]] - pos[tri[:, 1]], axis=1).mean())
    pos = pos * (1.0 + 0.05 * edge * rng.standard_normal((pos.shape[0], 1)))   # 5 % of an edge: steps get accepted
    nv, nf = pos.shape[0], tri.shape[0]
    local = split_mesh(nv, tri, world, rank)
    rows = local.global_rows()
    mods = L.MOD_SURFACE | L.MOD_BENDING
    common = dict(modules=mods, v_target=4.15, step_size=1e-3 * edge * edge, k_vol=0.0)   # sphere volume 4.18
    out = {"n_gpus": world, "facets": int(nf), "steps": steps, "cases": {}}
    for name, kw in CASES.items():
        pm = PartitionedMesh(local, local_rank, body_mask=np.ones(local.tri.shape[0], np.uint8))
        pm.dm.set_surface_tension(1.0)
        pm.dm.set_bending_params(1.0, 0.0)
        pm.dm.set_positions(pos[rows])
        mini = partitioned_minimizer(pm, **common, **kw)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = mini.minimize(steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert not pm.dm.halo_error()
        owned = pm.dm.download(L.ARR_POSITIONS)[: local.n_owned]
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((local.lo, owned), gathered, dst=0)
        hist = list(mini.history)
        pm.close()
        if rank == 0:
            one = DeviceMesh(local_rank)
            one.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8))
            one.set_surface_tension(1.0)
            one.set_bending_params(1.0, 0.0)
            one.set_positions(pos)
            ref = DeviceMinimizer(dm=one, **common, **kw)
            t0 = time.perf_counter()
            r1 = ref.minimize(steps)
            dt1 = time.perf_counter() - t0
            p1 = one.download(L.ARR_POSITIONS)
            one.close()
            p = np.zeros_like(p1)
            for lo, blk in gathered:
                p[lo:lo + len(blk)] = blk
            h, h1 = np.array(hist, dtype=float), np.array(ref.history, dtype=float)
            case = {
                "energy": res["energy"], "energy_one_gpu": r1["energy"],
                "energy_rel_err": abs(res["energy"] - r1["energy"]) / abs(r1["energy"]),
                "max_position_err": float(np.abs(p - p1).max()),
                "history_equal_1e-9": bool(h.shape == h1.shape and np.allclose(h, h1, rtol=1e-9, atol=0)),
                "accepted_steps": int(sum(1 for x in hist if x[3])),
                "seconds": dt, "seconds_one_gpu": dt1,
            }
            # parity with one GPU is the criterion; "moved" says whether the line search accepted anything at all
            case["ok"] = bool(case["energy_rel_err"] <= 1e-9 and case["max_position_err"] <= 1e-9
                              and case["history_equal_1e-9"])
            case["moved"] = case["accepted_steps"] > 0
            out["cases"][name] = case
    if rank == 0:
        out["ok"] = all(c["ok"] for c in out["cases"].values())
        print(json.dumps(out), flush=True)
    ok = torch.tensor([1 if (rank != 0 or out["ok"]) else 0], device=f"cuda:{local_rank}")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
