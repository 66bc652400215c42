"""torchrun -n N tools/multi_halo_check.py [facets_per_gpu]: the peer-memory halo against the NCCL send/recv halo on
the same partitioned mesh -- scalars and owned gradients must be BITWISE equal (only the transport differs) -- and
the time per distributed evaluation with either."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from membrane_solver_b200 import _lib as L  # noqa: E402
from membrane_solver_b200.partition import PartitionedMesh, split_mesh  # noqa: E402
from membrane_solver_b200.synthetic import frequency_for_facets, icosphere  # noqa: E402


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pos, tri = icosphere(frequency_for_facets(per_gpu * world))
    rng = np.random.default_rng(4)
    pos = pos * (1.0 + 0.01 * rng.standard_normal((pos.shape[0], 1)))
    local = split_mesh(pos.shape[0], tri, world, rank)
    pm = PartitionedMesh(local, local_rank, body_mask=np.ones(local.tri.shape[0], np.uint8))
    assert pm.transport == "peer", pm.transport
    dm = pm.dm
    dm.set_surface_tension(1.0)
    dm.set_bending_params(1.0, 0.0)
    rows = local.global_rows()
    # ghost rows start from garbage: the halo has to bring them
    start = pos[rows].copy()
    start[local.n_owned:] = 7.0
    opts = dm.options(L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME, constraint_mode=0)
    out = {}
    energy_opts = dm.options(L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME, want_grad=False)
    sv_opts = dm.options(L.MOD_SURFACE | L.MOD_VOLUME, constraint_mode=0)
    # "fused" = ms_ctx_eval_partition: with push targets set (default) the owners store into the ghost slots
    for transport in ("nccl", "peer", "fused", "inkernel", "nccl", "peer", "fused", "inkernel"):
        pm.transport = "nccl" if transport == "nccl" else "peer"
        pm.fused = transport in ("fused", "inkernel")
        pm.in_kernel = transport == "inkernel"
        dm.set_positions(start)
        res = pm.eval(opts)
        grad = dm.download(L.ARR_GRAD)[: local.n_owned]
        got_pos = dm.download(L.ARR_POSITIONS)
        assert np.array_equal(got_pos, pos[rows]), f"rank {rank}: ghost positions differ with {transport}"
        e_only = pm.eval(energy_opts)          # pass A alone finalises (and publishes, when fused)
        sv = pm.eval(sv_opts)                  # no pass A, no seed exchange
        sv_grad = dm.download(L.ARR_GRAD)[: local.n_owned]
        assert (e_only.e_surface, e_only.e_bending, e_only.volume) == (res.e_surface, res.e_bending, res.volume) \
            or transport == "nccl"
        key = (res.e_surface, res.e_bending, res.volume, res.kkt_lambda, sv.e_surface, sv.kkt_lambda,
               float(np.abs(sv_grad).sum()))
        if transport in out:
            assert out[transport][0] == key and np.array_equal(out[transport][1], grad), f"{transport} not repeatable"
        out[transport] = (key, grad)
        for _ in range(3):
            pm.eval_async(opts)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pm.eval_async(opts)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=pm.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[transport + "_ms"] = float(ms.item())
        assert not dm.halo_error()
    # the peer all-reduce adds in rank order; NCCL's order is its own: identical for two ranks, rounding beyond
    a, b = np.array(out["nccl"][0]), np.array(out["peer"][0])
    bitwise = bool(np.array_equal(a, b) and np.array_equal(out["nccl"][1], out["peer"][1]))
    assert np.all(np.abs(a - b) <= 1e-13 * np.maximum(1.0, np.abs(a))), (a, b)
    scale = np.abs(out["nccl"][1]).max()
    assert np.abs(out["nccl"][1] - out["peer"][1]).max() <= 1e-12 * scale
    assert bitwise or world > 2
    # folding the transport into the compute launches: the local sums are added by the last CTA (32 row groups)
    # instead of the reduce kernel (64 row groups), so scalars agree to rounding, not bitwise; each variant is
    # run-to-run repeatable (checked above)
    c = np.array(out["fused"][0])
    fused_scalar_err = float(np.max(np.abs(c - b) / np.maximum(1e-300, np.abs(b))))
    fused_grad_err = float(np.abs(out["fused"][1] - out["peer"][1]).max() / scale)
    assert fused_scalar_err <= 1e-12 and fused_grad_err <= 1e-13, (fused_scalar_err, fused_grad_err)
    # the exchange inside the patch kernels changes the order in which a CTA walks its patches, hence the order of the
    # per-CTA scalar sums: rounding-level differences in the scalars (and through lambda in the projected gradient)
    d = np.array(out["inkernel"][0])
    ik_scalar_err = float(np.max(np.abs(d - b) / np.maximum(1e-300, np.abs(b))))
    ik_grad_err = float(np.abs(out["inkernel"][1] - out["peer"][1]).max() / scale)
    assert ik_scalar_err <= 1e-12 and ik_grad_err <= 1e-13, (ik_scalar_err, ik_grad_err)
    # energy-only evaluation at trial positions (the line-search call) through the peer halo
    pm.transport, pm.fused, pm.in_kernel = "peer", True, True
    if rank == 0:
        print(json.dumps({"n_gpus": world, "facets": int(tri.shape[0]), "ghost_rows": int(local.ghost_ids.size),
                          "bitwise_equal": bitwise, "nccl_ms": out["nccl_ms"], "peer_ms": out["peer_ms"], "fused_ms": out["fused_ms"], "inkernel_ms": out["inkernel_ms"], "push": bool(pm.push),
                          "inkernel_vs_peer_scalar_rel_err": ik_scalar_err, "inkernel_vs_peer_grad_rel_err": ik_grad_err,
                          "fused_vs_peer_scalar_rel_err": fused_scalar_err, "fused_vs_peer_grad_rel_err": fused_grad_err,
                          "E_surface": out["peer"][0][0], "E_bending": out["peer"][0][1]}), flush=True)
    dist.barrier()
    dm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
