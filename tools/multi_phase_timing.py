"""Per-phase device timing of one distributed evaluation (diagnostic; torchrun --nproc-per-node N)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from membrane_solver_b200 import _lib as L
from membrane_solver_b200.partition import PartitionedMesh, split_mesh
from membrane_solver_b200.synthetic import frequency_for_facets, icosphere

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = frequency_for_facets(int(sys.argv[1]) * world if len(sys.argv) > 1 else 5_000_000 * world)
pos, tri = icosphere(n)
lm = split_mesh(pos.shape[0], tri, world, rank)
pm = PartitionedMesh(lm, local, body_mask=np.ones(lm.tri.shape[0], np.uint8))
dm = pm.dm
dm.set_surface_tension(1.0); dm.set_bending_params(1.0, 0.0); dm.set_positions(pos[lm.global_rows()])
opts = dm.options(L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME, constraint_mode=0)
for _ in range(5): pm.eval_async(opts)
torch.cuda.synchronize(); dist.barrier()
names = ["halo(pos)", "pass A", "halo(seeds)", "pass B", "reduce", "all-reduce", "project"]
acc = np.zeros(len(names)); reps = 20
for _ in range(reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record(); pm.exchange(L.ARR_POSITIONS)
    ev[1].record(); dm.eval_pass_a(opts)
    ev[2].record(); pm.exchange(L.ARR_SEEDS)
    ev[3].record(); dm.eval_pass_b(opts)
    ev[4].record(); dm.eval_reduce(opts)
    ev[5].record(); dist.all_reduce(pm.view(L.ARR_SCALARS)[:12])
    ev[6].record(); dm.eval_project(opts)
    ev[7].record(); torch.cuda.synchronize()
    acc += [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
if rank == 0:
    print({k: round(v / reps, 4) for k, v in zip(names, acc)}, "sum", round(acc.sum() / reps, 4), "facets/rank", lm.tri.shape[0])
dist.destroy_process_group()
