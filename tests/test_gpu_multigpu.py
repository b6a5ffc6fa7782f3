"""GPU, needs >= 2 devices (gpurun --gpus 2): planet sharded over two processes, halo slabs pulled over NVLink P2P via CUDA IPC
(tsom_ipc_export / tsom_ipc_attach / tsom_halo_pull); mesh(2 GPUs) must equal the single-process golden byte for byte in
ids / face counts / indices."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _worker(rank, world, port, name, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    import thisspaceofmine_b200 as T
    from thisspaceofmine_b200.sharding import ShardedPlanet

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        count = tuple(int(v) for v in g["count"])
        ctx = T.Context(rank)
        sp = ShardedPlanet(ctx, count, rank, world)
        ids = sp.plan.local_ids(rank)
        ok = True
        for rep in range(2):  # twice: the second round re-pulls into existing ghost slabs
            sp.GenerateChunks(ctx.blockLibrary, int(g["seed"]))
            sp.PullHalos()
            batch = sp.BuildMeshes(render=True)
            blocks = sp.planet.GetContent(sp.local_grid)
            ok &= np.array_equal(np.array([_crc(b) for b in blocks], np.uint32), g["block_crc"][ids])
            ok &= np.array_equal(batch.face_counts, g["face_count"][ids])
            for i in range(len(sp.local_grid)):
                m = batch.chunk(i)
                ok &= (0 if m is None else _crc(m.indices)) == g["index_crc"][ids[i]]
            dist.barrier()
        q.put((rank, bool(ok), int(batch.total_faces)))
        dist.barrier()
        sp.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("name", ["planet_7_3x3x3", "planet_42_5x5x5"])
def test_two_gpu_sharded_planet_matches_golden(name):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert sum(r[2] for r in res) == int(g["face_count"].sum())


def _edit_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle as O
    import thisspaceofmine_b200 as T
    from helpers import assert_mesh_equal
    from test_sharding_gloo import make_shard_edits
    from thisspaceofmine_b200.sharding import ShardedPlanet

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        count, seed = (3, 3, 3), 7
        ctx = T.Context(rank)
        sp = ShardedPlanet(ctx, count, rank, world)
        sp.GenerateChunks(ctx.blockLibrary, seed)
        sp.PullHalos()
        sp.BuildMeshes(render=True, collider=True)
        sp.CollectInvalidated(clear=True)

        whole = O.Planet(1.0, 16.0)  # the reference run: one process, whole planet
        assert whole.generate(seed, count, nthreads=2)
        whole.clear_invalidated()
        ok = True
        rebuilt = 0
        for tick in range(3):
            edits = make_shard_edits(sp.plan, 200, 100 + tick)
            sp.UpdateBlocks(edits)
            dirty, batch = sp.Update(render=True, collider=True)
            whole.apply_edits(edits)
            want = {tuple(k) for k in whole.collect_dirty().tolist()}
            whole.clear_invalidated()
            mine = {tuple(k) for k in sp.local_grid.tolist()}
            ok &= {tuple(k) for k in dirty.tolist()} == (want & mine)
            for i, k in enumerate(dirty.tolist()):
                assert_mesh_equal(batch.chunk(i), whole.mesh(k), f"rank {rank} tick {tick} chunk {k}")
            got = sp.planet.GetContent(sp.local_grid)
            for i, k in enumerate(sp.local_grid.tolist()):
                ok &= np.array_equal(got[i], whole.get_blocks(k))
            rebuilt += len(dirty)
        ctx.check_guards()
        q.put((rank, bool(ok), rebuilt))
        dist.barrier()
        sp.close()
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_edit_ticks_match_single_process():
    """Edits on a planet sharded over two GPUs: owners apply, both sides of the rank boundary rebuild what the reference's neighbour
    mask says (ghost slabs re-pulled over P2P first), blocks and meshes equal the single-process oracle replay."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_edit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert all(r[2] > 0 for r in res)
