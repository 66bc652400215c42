"""The peer-memory transport of the multi-GPU path on ONE GPU: three partitions of a mesh as
three contexts of this process, each on its own stream, wired to each other with device pointers
(``ms_ctx_peer_set_pointer``) instead of CUDA IPC handles.  The kernels are the ones the multi-process path runs:
epoch flags, wait + pull of the ghost rows, rank-order all-reduce of the scalars."""

import numpy as np
import pytest

from ms_test_helpers import rel_err


@pytest.mark.gpu
def test_three_partitions_exchange_over_peer_pointers():
    import torch

    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh
    from membrane_solver_b200.partition import ghost_sources, split_mesh
    from membrane_solver_b200.synthetic import icosphere

    pos, tri = icosphere(40)
    rng = np.random.default_rng(9)
    pos = pos * (1.0 + 0.02 * rng.standard_normal((pos.shape[0], 1)))
    nv, nf = pos.shape[0], tri.shape[0]
    mods = L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME

    whole = DeviceMesh(0)
    whole.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8))
    whole.set_surface_tension(1.0)
    whole.set_bending_params(1.2, 0.05)
    whole.set_positions(pos)
    want = whole.eval(whole.options(mods, constraint_mode=0))
    want_g = whole.download(L.ARR_GRAD)
    want_e = whole.eval(whole.options(mods, want_grad=False))
    whole.close()

    world = 3
    lib = L.lib()
    parts = [split_mesh(nv, tri, world, r) for r in range(world)]
    dms, streams = [], []
    for loc in parts:
        dm = DeviceMesh(0)
        dm.set_topology(loc.nv_local, loc.tri, n_owned=loc.n_owned, body_mask=np.ones(loc.tri.shape[0], np.uint8))
        dm.set_surface_tension(1.0)
        dm.set_bending_params(1.2, 0.05)
        start = pos[loc.global_rows()].copy()
        start[loc.n_owned:] = 7.0                     # ghost rows must come through the halo
        dm.set_positions(start)
        stream = torch.cuda.Stream()
        L.check(lib.ms_ctx_set_stream(dm._h, stream.cuda_stream))
        dms.append(dm)
        streams.append(stream)
    torch.cuda.synchronize()
    for r, (loc, dm) in enumerate(zip(parts, dms)):
        owners, rows = ghost_sources(loc)
        for o in np.unique(owners):
            for which in (L.ARR_POSITIONS, L.ARR_TRIAL, L.ARR_SEEDS):
                dm.peer_set_pointer(int(o), which, dms[int(o)].device_ptr(which))
        for o in range(world):
            if o != r:
                dm.peer_set_pointer(o, L.IPC_FLAGS, dms[o].flag_words_ptr())
        dm.set_rank_slot(r, world)
        dm.set_ghost_sources(world, owners, rows)
    for dm in dms:
        dm.halo_prepare()      # allocations / NULL-stream copies would serialise the three streams of this process
    torch.cuda.synchronize()

    def evaluate(opts):
        # every context issues the same sequence; the waits resolve on the device across the three streams
        for dm in dms:
            dm.halo_signal(L.FLAG_POSITIONS)
            dm.halo_pull(L.ARR_POSITIONS, L.FLAG_POSITIONS)
            dm.eval_pass_a(opts)
            if opts.want_grad:
                dm.halo_signal(L.FLAG_SEEDS)
                dm.halo_pull(L.ARR_SEEDS, L.FLAG_SEEDS)
            dm.eval_pass_b(opts)
            dm.eval_reduce(opts)
            dm.allreduce_scalars(12)
            dm.eval_project(opts)
        out = [dm.read_scalars() for dm in dms]
        assert not any(dm.halo_error() for dm in dms)
        return out

    opts = dms[0].options(mods, constraint_mode=0)
    for step in range(3):                              # repeated: epochs advance, slots alternate
        res = evaluate(opts)
        for r in res:
            assert r.scalars.tobytes() == res[0].scalars.tobytes()       # rank-order sum: bitwise the same everywhere
        for name in ("e_surface", "e_bending", "volume", "area", "kkt_lambda"):
            a, b = getattr(res[0], name), getattr(want, name)
            assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (name, a, b)
        grad = np.concatenate([dm.download(L.ARR_GRAD)[: loc.n_owned] for dm, loc in zip(dms, parts)])
        assert rel_err(grad, want_g) <= 2e-12
        for dm, loc in zip(dms, parts):
            assert np.array_equal(dm.download(L.ARR_POSITIONS), pos[loc.global_rows()])
    res = evaluate(dms[0].options(mods, want_grad=False))      # energy-only: positions halo + all-reduce alone
    assert abs(res[0].e_bending - want_e.e_bending) <= 1e-12 * abs(want_e.e_bending)
    for dm in dms:
        L.check(lib.ms_ctx_set_stream(dm._h, None))
        dm.close()


def test_ghost_sources_point_at_the_owners_rows():
    """CPU tier: the (owner rank, owner-local row) table the peer pulls read through."""
    from membrane_solver_b200.partition import ghost_sources, split_mesh
    from membrane_solver_b200.synthetic import icosphere

    pos, tri = icosphere(12)
    nv = pos.shape[0]
    for world in (2, 3, 5):
        parts = [split_mesh(nv, tri, world, r) for r in range(world)]
        for loc in parts:
            owners, rows = ghost_sources(loc)
            assert owners.shape == rows.shape == loc.ghost_ids.shape
            assert not np.any(owners == loc.rank)
            for o in np.unique(owners):
                sel = owners == o
                owner = parts[int(o)]
                assert np.all(rows[sel] < owner.n_owned)
                assert np.array_equal(owner.global_rows()[rows[sel]], loc.ghost_ids[sel])


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [1, 3], ids=["separate-launches", "inside-the-patch-kernels"])
def test_partition_entry_point_with_one_rank(flags):
    """``ms_ctx_eval_partition`` on a "partition" that is the whole mesh (one rank, no ghost rows): the entry point
    the multi-process path calls per evaluation runs in the 1-GPU tier too -- the rank's flag raised by the kernel,
    interior-first patch order, the last CTA reducing, publishing AND gathering the scalars, the multiplier from
    the gathered sums -- and must give the plain evaluation's results (scalars to rounding: the per-CTA sums are
    taken over a different patch order; gradient rows bitwise)."""
    from membrane_solver_b200 import _lib as L
    from membrane_solver_b200.context import DeviceMesh
    from membrane_solver_b200.synthetic import icosphere

    pos, tri = icosphere(48)
    rng = np.random.default_rng(5)
    pos = pos * (1.0 + 0.02 * rng.standard_normal((pos.shape[0], 1)))
    nv, nf = pos.shape[0], tri.shape[0]
    mods = L.MOD_SURFACE | L.MOD_BENDING | L.MOD_VOLUME
    dm = DeviceMesh(0)
    dm.set_topology(nv, tri, body_mask=np.ones(nf, np.uint8))
    dm.set_surface_tension(1.0)
    dm.set_bending_params(1.2, 0.05)
    dm.set_positions(pos)
    opts = dm.options(mods, constraint_mode=0)
    want = dm.eval(opts)
    want_raw = dm.download(L.ARR_GRAD)            # projected gradient of the plain evaluation
    dm.ipc_export(L.IPC_FLAGS)                    # allocates the flag words
    dm.set_rank_slot(0, 1)
    dm.set_ghost_sources(1, np.zeros(0, np.int32), np.zeros(0, np.int32))
    dm.halo_prepare()
    for _ in range(3):                            # epochs advance, the counters of the in-kernel form too
        dm.eval_partition(opts, True, in_kernel=(flags == 3))
    got = dm.read_scalars()
    assert not dm.halo_error()
    for name in ("e_surface", "e_bending", "volume", "area", "kkt_lambda"):
        a, b = getattr(got, name), getattr(want, name)
        assert abs(a - b) <= 1e-12 * max(1.0, abs(b)), (name, a, b)
    assert rel_err(dm.download(L.ARR_GRAD), want_raw) <= 1e-13
    # energy-only evaluation: pass A alone publishes and gathers
    e_opts = dm.options(mods, want_grad=False)
    dm.eval_partition(e_opts, True, in_kernel=(flags == 3))
    e_only = dm.read_scalars()
    assert abs(e_only.e_bending - want.e_bending) <= 1e-12 * abs(want.e_bending)
    dm.close()
