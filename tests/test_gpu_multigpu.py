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
        s, e = sp.plan.ranges[rank]
        ok = True
        for rep in range(2):  # twice: the second round re-pulls into existing ghost slabs
            sp.GenerateChunks(ctx.blockLibrary, int(g["seed"]))
            sp.PullHalos()
            batch = sp.BuildMeshes(render=True)
            blocks = sp.planet.GetContent(sp.local_grid)
            ok &= np.array_equal(np.array([_crc(b) for b in blocks], np.uint32), g["block_crc"][s:e])
            ok &= np.array_equal(batch.face_counts, g["face_count"][s:e])
            for i in range(len(sp.local_grid)):
                m = batch.chunk(i)
                ok &= (0 if m is None else _crc(m.indices)) == g["index_crc"][s + i]
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
