"""Host-side logic of the multi-GPU path, on the CPU.

* the partition plan (owned ranges, ghosts, send lists) is consistent;
* a world_size-2 run over ``gloo`` that evaluates each partition with the TEST-ONLY host
  emulator (same packer and per-facet code as the kernels), exchanges the seed halo
  between pass A and pass B and all-reduces the scalars reproduces the single-domain oracle.
"""

import os
import socket
import sys

import numpy as np
import pytest

from ms_test_helpers import rel_err
from membrane_solver_b200 import partition as part
from membrane_solver_b200.synthetic import icosphere


@pytest.mark.parametrize("world", [2, 3, 8])
def test_split_covers_mesh(world):
    pos, tri = icosphere(12)
    nv, nf = pos.shape[0], tri.shape[0]
    locals_ = [part.split_mesh(nv, tri, world, r) for r in range(world)]
    assert sum(m.n_owned for m in locals_) == nv
    primary = np.zeros(nf, int)
    for m in locals_:
        rows = m.global_rows()
        assert np.array_equal(rows[m.tri], tri[m.facet_ids])      # local numbering maps back
        assert np.all(np.diff(m.ghost_ids) > 0)
        assert not np.any((m.ghost_ids >= m.lo) & (m.ghost_ids < m.lo + m.n_owned))
        # every facet around an owned vertex is listed
        touching = ((tri >= m.lo) & (tri < m.lo + m.n_owned)).any(axis=1)
        assert np.array_equal(np.nonzero(touching)[0], m.facet_ids)
        primary[m.facet_ids[m.tri[:, 0] < m.n_owned]] += 1          # owner of the first vertex
        covered = sum(c for _, _, c in m.recv_blocks)
        assert covered == m.ghost_ids.size
    assert np.all(primary == 1)
    ghosts = [m.ghost_ids for m in locals_]
    for m in locals_:
        for dst, rows in part.send_lists(m, ghosts):
            blk = [b for b in locals_[dst].recv_blocks if b[0] == m.rank]
            assert len(blk) == 1 and blk[0][2] == rows.size
            a = blk[0][1]
            assert np.array_equal(locals_[dst].ghost_ids[a:a + rows.size], rows.astype(np.int64) + m.lo)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ms_test_helpers as H

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pos, tri = icosphere(14)
    nv = pos.shape[0]
    local = part.split_mesh(nv, tri, world, rank)
    rows = local.global_rows()
    lpos = np.ascontiguousarray(pos[rows])
    gathered = [None] * world
    dist.all_gather_object(gathered, local.ghost_ids)
    sends = part.send_lists(local, gathered)
    halo = part.HaloExchange(local, sends, dist, torch, torch.device("cpu"))
    kw = dict(modules=H.MOD_SURFACE | H.MOD_BENDING | H.MOD_VOLUME, body_mask=np.ones(len(local.tri), np.uint8),
              kappa_u=1.0, c0_u=0.05, n_owned=local.n_owned, threads=32, max_owned=40, max_local=160)
    # halo(positions): ghost rows arrive from their owners
    tpos = torch.from_numpy(lpos.copy())
    tpos[local.n_owned:] = 0.0
    halo.exchange(tpos, lambda buf: buf.copy_(tpos[torch.from_numpy(halo.send_rows.astype(np.int64))]))
    assert np.array_equal(tpos.numpy(), lpos)
    a = H.emulate(tpos.numpy(), local.tri, phase=1, **kw)
    seeds = torch.from_numpy(a["seeds"])
    halo.exchange(seeds, lambda buf: buf.copy_(seeds[torch.from_numpy(halo.send_rows.astype(np.int64))]))
    b = H.emulate(tpos.numpy(), local.tri, phase=2, seeds=seeds.numpy(), **kw)
    sc = torch.tensor([a["E_surface"], a["area"], a["volume"], a["E_bending"]], dtype=torch.float64)
    dist.all_reduce(sc)
    grads = [None] * world
    dist.all_gather_object(grads, (local.lo, b["grad"][:local.n_owned], b["volgrad"][:local.n_owned]))
    if rank == 0:
        g = np.zeros((nv, 3))
        vg = np.zeros((nv, 3))
        for lo, gg, vv in grads:
            g[lo:lo + len(gg)] = gg
            vg[lo:lo + len(vv)] = vv
        np.savez(out_path, sc=sc.numpy(), grad=g, volgrad=vg)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_gloo_matches_oracle(tmp_path, world):
    import torch.multiprocessing as mp

    from oracle import ref_modules as ref

    out = str(tmp_path / "out.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    pos, tri = icosphere(14)
    nv, nf = pos.shape[0], tri.shape[0]
    want = ref.fused_surface_bending_volume(pos, tri, np.ones(nf), 1.0, 0.05, np.zeros(nv, bool))
    for k, name in enumerate(("E_surface", "area", "volume", "E_bending")):
        assert abs(got["sc"][k] - want[name]) <= 1e-12 * abs(want[name]), name
    assert rel_err(got["grad"], want["grad"]) <= 1e-12
    assert rel_err(got["volgrad"], want["vol_grad"]) <= 1e-12


def _minimizer_worker(rank, world, port, out_path):
    """Partitioned minimiser facade over gloo: every rank holds an emulated device with the whole mesh but reports
    the line-search statistics of ITS row range only; the facade's reductions must rebuild the global values and
    keep the ranks on the same branch of the loop."""
    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.runtime.partitioned_minimizer import partitioned_minimizer

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pos, tri = icosphere(6)
    rng = np.random.default_rng(11)
    pos = pos * (1.0 + 0.03 * rng.standard_normal((pos.shape[0], 1)))
    nv = pos.shape[0]
    cuts = part.vertex_cuts(nv, world)
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])

    class Sliced(FakeDeviceMesh):
        def line_search_stats(self):
            t, p = self.tri, self.pos
            mine = (t[:, 0] >= lo) & (t[:, 0] < hi)
            tm = t[mine]
            e = np.concatenate([p[tm[:, 2]] - p[tm[:, 1]], p[tm[:, 0]] - p[tm[:, 2]], p[tm[:, 1]] - p[tm[:, 0]]])
            g, d = self.arrays[L.ARR_GRAD][lo:hi], self.dir[lo:hi]
            return (float(np.sqrt((e * e).sum(axis=1).min())), float(np.sqrt((d**2).sum(axis=1).max())),
                    float((g * d).sum()), float((g * g).sum()))

    class FakePartition:
        def __init__(self):
            self.dm = Sliced(threads=32, max_owned=40, max_local=160)
            self.dist, self.torch, self.device = dist, torch, torch.device("cpu")
            self.exchanges = 0

        def eval(self, opts):
            return self.dm.eval(opts)

        def exchange(self, which):
            self.exchanges += 1

    pm = FakePartition()
    pm.dm.set_topology(nv, tri, body_mask=np.ones(len(tri), np.uint8))
    pm.dm.set_surface_tension(1.0)
    pm.dm.set_positions(pos)
    mini = partitioned_minimizer(pm, modules=L.MOD_SURFACE, volume_mode="lagrange", v_target=4.0, step_size=1e-2)
    res = mini.minimize(4)
    if rank == 0:
        np.savez(out_path, energy=res["energy"], pos=pm.dm.pos, history=np.array(mini.history, dtype=float))
    dist.destroy_process_group()


def test_partitioned_minimizer_facade_matches_single_domain(tmp_path):
    import torch.multiprocessing as mp

    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)
    from fake_device import FakeDeviceMesh

    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.runtime.device_minimizer import DeviceMinimizer

    out = str(tmp_path / "mini.npz")
    mp.spawn(_minimizer_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    pos, tri = icosphere(6)
    rng = np.random.default_rng(11)
    pos = pos * (1.0 + 0.03 * rng.standard_normal((pos.shape[0], 1)))
    dm = FakeDeviceMesh(threads=32, max_owned=40, max_local=160)
    dm.set_topology(pos.shape[0], tri, body_mask=np.ones(len(tri), np.uint8))
    dm.set_surface_tension(1.0)
    dm.set_positions(pos)
    mini = DeviceMinimizer(dm=dm, modules=L.MOD_SURFACE, volume_mode="lagrange", v_target=4.0, step_size=1e-2)
    res = mini.minimize(4)
    assert abs(float(got["energy"]) - res["energy"]) <= 1e-12 * abs(res["energy"])
    assert np.max(np.abs(got["pos"] - dm.pos)) <= 1e-12
    assert np.allclose(got["history"], np.array(mini.history, dtype=float), rtol=1e-12, atol=0)
    assert any(h[3] for h in mini.history)      # at least one accepted step: the loop really moved


def _push_targets_worker(rank, world, port, out_dir):
    """``PartitionedMesh._push_targets`` without a device: every rank computes, for the rows its neighbours list as
    ghosts, (neighbour, row here, row there); applying all of them to per-rank arrays must fill every ghost row with
    the owner's data -- the same result as the gloo halo exchange."""
    import torch
    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.dirname(here)):
        if p not in sys.path:
            sys.path.insert(0, p)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pos, tri = icosphere(10)
    nv = pos.shape[0]
    local = part.split_mesh(nv, tri, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, local.ghost_ids)
    sends = part.send_lists(local, gathered)
    layout = [None] * world
    dist.all_gather_object(layout, (int(local.n_owned), [tuple(int(x) for x in b) for b in local.recv_blocks]))

    class Shell:   # the two attributes _push_targets reads, no device behind them
        pass

    pm = Shell()
    pm.local = local
    pm.halo = part.HaloExchange(local, sends, dist, torch, torch.device("cpu"))
    pm.L = __import__("membrane_solver_b200._lib", fromlist=["_lib"])
    slots, src, dst = part.PartitionedMesh._push_targets(pm, layout)
    assert slots.shape == src.shape == dst.shape and slots.size == pm.halo.send_rows.size
    assert np.all(src < local.n_owned) and np.all(slots != rank)
    rows = local.global_rows()
    np.savez(os.path.join(out_dir, f"push_{rank}.npz"), slots=slots, src_global=rows[src], dst=dst, rows=rows,
             n_owned=local.n_owned)
    dist.destroy_process_group()


def test_push_targets_fill_every_ghost_row(tmp_path):
    import torch.multiprocessing as mp

    world = 3
    mp.spawn(_push_targets_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    data = [np.load(str(tmp_path / f"push_{r}.npz")) for r in range(world)]
    # simulate the stores: the value is the global id of the owner's row
    arrays = [np.where(np.arange(d["rows"].size) < int(d["n_owned"]), d["rows"], -1) for d in data]
    for r, d in enumerate(data):
        for slot, g, dst in zip(d["slots"], d["src_global"], d["dst"]):
            assert dst >= int(data[slot]["n_owned"])          # lands in the ghost region of the neighbour
            assert arrays[slot][dst] == -1                    # no ghost row is written twice
            arrays[slot][dst] = g
    for r, d in enumerate(data):
        assert np.array_equal(arrays[r], d["rows"])           # every ghost row holds its owner's row
