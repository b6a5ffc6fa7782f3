"""GPU (one device is enough): the single-process multi-GPU planet, tsom_planet_create_sharded.

The parts of a sharded planet may share a device, so the whole multi-part path — box plan, per-part generation with per-owner
terrain tables, peer-mapped halo slabs, the device-side epoch flags of tsom_halo_pull, routed edits, cross-part dirty
propagation — runs here on cuda:0 with 2, 3 and 8 parts and must reproduce the single-process golden fixtures / the oracle
byte for byte in ids, face counts and index buffers.  With >= 2 devices the same tests also spread the parts over them."""
import os
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _devices(parts):
    import torch

    n = max(1, torch.cuda.device_count())
    return [i % n for i in range(parts)]


@pytest.mark.parametrize("name,parts", [("planet_7_3x3x3", 2), ("planet_7_3x3x3", 3), ("planet_42_5x5x5", 2), ("planet_42_5x5x5", 8)])
def test_sharded_planet_in_one_process_matches_golden(name, parts):
    import thisspaceofmine_b200 as T

    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    count = tuple(int(v) for v in g["count"])
    mp = T.MultiGpuPlanet(_devices(parts), count)
    try:
        assert mp.parts == parts
        owners = mp.owner(mp.grid)
        assert set(owners.tolist()) == set(range(parts))
        for rep in range(2):  # twice: the second round regenerates in place and re-pulls into existing ghost slabs
            mp.GenerateChunks(int(g["seed"]))
            views, where, flags = mp.BuildMeshes(render=True)
            assert np.array_equal(where, owners)
            blocks = mp.GetContent(mp.grid)
            assert np.array_equal(np.array([_crc(b) for b in blocks], np.uint32), g["block_crc"]), "block ids differ from the single-process golden"
            assert np.array_equal(views["face_count"], g["face_count"]), "face counts differ from the single-process golden"
            for i in range(len(mp.grid)):
                m = mp.chunk_mesh(views[i], where[i], flags)
                assert (0 if m is None else _crc(m.indices)) == g["index_crc"][i], f"index buffer of chunk {mp.grid[i].tolist()}"
    finally:
        mp.close()


def test_sharded_planet_vertices_and_collider_equal_single_gpu(ctx):
    """mesh(3 parts) == mesh(1 GPU): vertices, indices and collider positions of every chunk, flat and deformed."""
    import thisspaceofmine_b200 as T

    count = (3, 3, 3)
    grid = T.Planet.chunk_grid(count)
    one = T.Planet(ctx, capacity=27)
    one.GenerateChunks(ctx.blockLibrary, 11, count)
    mp = T.MultiGpuPlanet(_devices(3), count)
    try:
        mp.GenerateChunks(11)
        for kw in (dict(), dict(deformed=True, deformRadius=16.0)):
            ref = one.BuildMeshes(grid, render=True, collider=True, **kw)
            views, where, flags = mp.BuildMeshes(render=True, collider=True, **kw)
            assert np.array_equal(views["face_count"], ref.face_counts)
            for i in range(len(grid)):
                a, b = mp.chunk_mesh(views[i], where[i], flags), ref.chunk(i)
                assert (a is None) == (b is None)
                if a is not None:
                    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.indices, b.indices) and np.array_equal(a.collider_positions, b.collider_positions)
    finally:
        mp.close()
        one.close()


def test_sharded_edit_ticks_match_the_oracle():
    """Edits routed to their owners, dirty neighbours across part boundaries, a fresh exchange round before the rebuild: blocks,
    dirty sets and rebuilt meshes equal the single-process oracle replay (Planet.cpp:39-58, ChunkEntities.cpp:108-111)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O
    import thisspaceofmine_b200 as T
    from helpers import assert_mesh_equal
    from test_sharding_gloo import make_shard_edits
    from thisspaceofmine_b200.sharding import ShardPlan

    count, seed = (3, 3, 3), 7
    mp = T.MultiGpuPlanet(_devices(4), count)
    try:
        mp.GenerateChunks(seed)
        mp.BuildMeshes(render=True, views=False)
        assert len(mp.CollectInvalidated(clear=True)) == 27  # OnReset invalidates every generated chunk
        whole = O.Planet(1.0, 16.0)
        assert whole.generate(seed, count, nthreads=2)
        whole.clear_invalidated()
        plan = ShardPlan(count, 4)
        rebuilt = 0
        for tick in range(3):
            edits = make_shard_edits(plan, 200, 100 + tick)
            edits[5, 0:3] = (9, 9, 9)  # outside the planet: skipped
            mp.UpdateBlocks(edits)
            whole.apply_edits(edits[np.arange(len(edits)) != 5])
            dirty = mp.CollectInvalidated(clear=True)
            want = whole.collect_dirty()
            whole.clear_invalidated()
            assert sorted(map(tuple, dirty.tolist())) == sorted(map(tuple, want.tolist())), f"dirty set of tick {tick}"
            assert dirty.tolist() == sorted(dirty.tolist()), "sorted by (x, y, z)"
            views, where, flags = mp.BuildMeshes(dirty, render=True, collider=True)
            for i, k in enumerate(dirty.tolist()):
                assert_mesh_equal(mp.chunk_mesh(views[i], where[i], flags), whole.mesh(k), f"tick {tick} chunk {k}")
            got = mp.GetContent(mp.grid)
            for i, k in enumerate(mp.grid.tolist()):
                assert np.array_equal(got[i], whole.get_blocks(k)), f"blocks of {k} after tick {tick}"
            rebuilt += len(dirty)
        assert rebuilt > 0
    finally:
        mp.close()


def test_mesh_refuses_unpulled_ghosts_and_exchange_times_out(ctx):
    """tsom_mesh must not cull against a remote slab that was never pulled, and a peer that never publishes shows up as
    TSOM_ERR_INTERNAL after the device-side time-out (200 ms here, 5 s by default) instead of hanging the GPU."""
    import ctypes as C

    import thisspaceofmine_b200 as T
    from thisspaceofmine_b200 import _lib as L

    a, b = T.Planet(ctx, capacity=1), T.Planet(ctx, capacity=1)
    try:
        ia, ib = np.array([[0, 0, 0]], np.int32), np.array([[1, 0, 0]], np.int32)
        blocks = np.full((1, 32768), 7, np.uint16)
        a.ResetChunks(ia, blocks)
        b.ResetChunks(ib, blocks)
        L.check(L.lib().tsom_peer_attach_local(a._h, 1, b._h))
        slot = np.zeros(1, np.uint32)
        rank = np.ones(1, np.int32)
        L.check(L.lib().tsom_container_add_remote_chunks(a._h, ib.ctypes.data_as(C.POINTER(C.c_int32)), slot.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                         rank.ctypes.data_as(C.POINTER(C.c_int32)), 1))
        with pytest.raises(T.TsomError) as e:
            a.BuildMeshes(ia)
        assert e.value.code == L.ERR_INVALID and "tsom_halo_pull" in str(e.value)
        # b publishes, a pulls: the +X face of a is now culled against b's stone
        b.PublishHalo()
        a.PullHalos()
        assert a.BuildMeshes(ia).face_counts.tolist() == [5 * 1024]
        # next round: b never publishes -> a's pull polls b's flag, gives up after the time-out and the next blocking call reports it
        L.check(L.lib().tsom_halo_set_timeout(a._h, 200))
        a.PullHalos()
        with pytest.raises(T.TsomError) as e:
            a.BuildMeshes(ia)
        assert e.value.code == L.ERR_INTERNAL and "timed out" in str(e.value)
        ctx.synchronize()
    finally:
        a.close()
        b.close()
