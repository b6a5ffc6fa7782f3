"""CPU, world_size 2 over gloo: the host-side sharding logic of thisspaceofmine_b200.sharding.

Each rank generates its chunk range with the oracle, fetches ONLY the boundary slabs ShardPlan says it needs from the other rank
(the CPU stand-in for tsom_halo_pull's P2P copy), meshes its range and must reproduce the single-process golden fixture
(tests/golden/planet_7_3x3x3.npz): mesh(N ranks) == mesh(1 rank)."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

from thisspaceofmine_b200.sharding import CHUNK_DIR_OFFSET, ShardPlan, chunk_grid, plan_boxes  # noqa: E402


def test_plan_boxes_tile_the_planet_and_balance():
    """tsom_shard_plan (host code of the C library): the boxes are disjoint, cover every chunk, and split a cubic planet into
    near-equal parts; an explicit weight array moves the cut to where the weight is."""
    for cc, w in [((5, 5, 5), 1), ((5, 5, 5), 2), ((5, 5, 5), 8), ((3, 3, 3), 4), ((47, 47, 47), 8), ((31, 31, 31), 3), ((7, 3, 5), 5)]:
        plan = ShardPlan(cc, w)
        counts = np.bincount(plan.owner, minlength=w)
        assert counts.sum() == plan.n and np.all(counts > 0)
        vol = [int(np.prod(b[3:6] - b[0:3] + 1)) for b in plan.boxes]
        assert vol == counts.tolist(), "every box holds exactly the chunks it owns (no overlap)"
        for r in range(w):
            assert np.array_equal(plan.slot_of(plan.local_ids(r)), np.arange(counts[r])), "slot == position in the owner's box order"
    b = plan_boxes((47, 47, 47), 8)
    sizes = np.prod(b[:, 3:6] - b[:, 0:3] + 1, axis=1)
    assert sizes.max() / sizes.min() < 1.15, "2 x 2 x 2 octants of a cubic planet"
    # all the weight in the top two z-layers: the first cut isolates them
    wts = np.ones((9, 9, 9), np.float32)
    wts[7:, :, :] = 100.0
    b = plan_boxes((9, 9, 9), 2, wts)
    assert b[0].tolist() == [-4, -4, -4, 4, 4, 3] and b[1].tolist() == [-4, -4, 4, 4, 4, 4]
    with pytest.raises(Exception):
        plan_boxes((1, 1, 2), 3)


def test_chunk_grid_is_generatechunks_order():
    g = chunk_grid((5, 5, 5))
    assert g[0].tolist() == [-2, -2, -2] and g[1].tolist() == [-1, -2, -2] and g[5].tolist() == [-2, -1, -2] and g[25].tolist() == [-2, -2, -1]
    plan = ShardPlan((5, 3, 7), 3)
    assert np.array_equal(plan.linear_id(plan.grid), np.arange(plan.n))
    assert plan.linear_id([[3, 0, 0], [0, 2, 0], [0, 0, -4]]).tolist() == [-1, -1, -1]


def test_remote_neighbours_are_symmetric_and_face_adjacent_only():
    plan = ShardPlan((5, 5, 5), 4)
    rem = [plan.remote_neighbours(r) for r in range(4)]
    for r in range(4):
        mine = set(map(tuple, plan.local(r).tolist()))
        for peer, lids in rem[r].items():
            assert peer != r and np.all(plan.owner[lids] == peer)
            for idx in plan.grid[lids]:
                assert any(tuple((idx - CHUNK_DIR_OFFSET[d]).tolist()) in mine for d in range(6))
            # symmetry: if I need chunks from peer, peer needs chunks from me
            assert r in rem[peer]
    # boxes: a rank's halo is the surface of its box that lies inside the planet, one slab per boundary chunk face
    b = plan.boxes[0]
    ext = b[3:6] - b[0:3] + 1
    assert plan.halo_bytes(0) <= 2 * int(ext[0] * ext[1] + ext[0] * ext[2] + ext[1] * ext[2]) * 2048
    ids = plan.local_ids(2)
    assert plan.slot_of(ids).tolist() == list(range(len(ids)))
    assert ShardPlan((3, 3, 3), 1).remote_neighbours(0) == {}


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _facing_layer(blocks, d):
    """The layer of a neighbour chunk (towards direction d from me) that faces me, as a full ghost chunk (zeros elsewhere)."""
    b = blocks.reshape(32, 32, 32)  # [z(up), y(world Z), x]
    g = np.zeros_like(b)
    # Direction order Back(+Z: local y+1), Down(local z-1), Front(local y-1), Left(x-1), Right(x+1), Up(local z+1)
    if d == 0:
        g[:, 0, :] = b[:, 0, :]
    elif d == 1:
        g[31, :, :] = b[31, :, :]
    elif d == 2:
        g[:, 31, :] = b[:, 31, :]
    elif d == 3:
        g[:, :, 31] = b[:, :, 31]
    elif d == 4:
        g[:, :, 0] = b[:, :, 0]
    else:
        g[0, :, :] = b[0, :, :]
    return g.reshape(-1)


def _worker(rank, world, port, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(GOLDEN, "planet_7_3x3x3.npz"))
        count = tuple(int(v) for v in g["count"])
        seed = int(g["seed"])
        plan = ShardPlan(count, world)
        ids = plan.local_ids(rank)
        mine = plan.local(rank)
        blocks = {}
        for idx in mine:
            b, ok = O.generate_chunk(seed, count, idx)
            assert ok
            blocks[tuple(idx.tolist())] = b

        # "halo pull": publish which (linear id, direction) slabs I need, owners answer with exactly those slabs
        need = []  # (owner, linear id of the neighbour, direction from my chunk)
        for idx in mine:
            for d in range(6):
                lid = int(plan.linear_id(idx + CHUNK_DIR_OFFSET[d])[0])
                if lid >= 0 and plan.owner[lid] != rank:
                    need.append((int(plan.owner[lid]), lid, d))
        # the plan's remote list must cover exactly the owners/chunks found by the brute-force walk
        rem = plan.remote_neighbours(rank)
        assert {(o, l) for o, l, _ in need} == {(p, int(l)) for p, ls in rem.items() for l in ls}
        all_need = [None] * world
        dist.all_gather_object(all_need, need)
        answers = []
        for r in range(world):
            answers.append([(lid, d, _facing_layer(blocks[tuple(plan.grid[lid].tolist())], d)) for o, lid, d in all_need[r] if o == rank])
        got = [None] * world
        dist.all_gather_object(got, answers)  # got[p][rank] = slabs rank p owns that I asked for
        assert sum(len(got[p][rank]) for p in range(world)) * 2048 == plan.halo_bytes(rank)

        planet = O.Planet(1.0, 16.0)
        for k, b in blocks.items():
            planet.add_chunk(k, b)
        ghosts = {}
        for p in range(world):
            for lid, d, layer in got[p][rank]:
                k = tuple(plan.grid[lid].tolist())
                ghosts[k] = layer if k not in ghosts else np.maximum(ghosts[k], layer)  # a ghost can face two of my chunks
        for k, b in ghosts.items():
            planet.add_chunk(k, b)

        faces, icrc = [], []
        for idx in mine:
            m = planet.mesh(idx)
            faces.append(0 if m is None else m.face_count)
            icrc.append(0 if m is None else _crc(m.indices))
        ok = np.array_equal(np.array(faces, np.uint32), g["face_count"][ids]) and np.array_equal(np.array(icrc, np.uint32), g["index_crc"][ids])
        bok = all(_crc(blocks[tuple(idx.tolist())]) == g["block_crc"][ids[i]] for i, idx in enumerate(mine))
        result_q.put((rank, bool(ok), bool(bok), int(sum(faces))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_two_rank_sharded_planet_matches_single_process_golden():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res), "sharded meshes differ from the single-process golden"
    assert all(r[2] for r in res), "sharded block ids differ from the single-process golden"
    g = np.load(os.path.join(GOLDEN, "planet_7_3x3x3.npz"))
    assert sum(r[3] for r in res) == int(g["face_count"].sum())


# ------------------------------------------------------------------------------------------------ edits on a sharded planet
def make_shard_edits(plan, n, seed):
    """Random edits with many boundary coordinates (0 / 31), so that a good part of them invalidates chunks of the other rank."""
    rng = np.random.default_rng(seed)
    e = np.zeros((n, 7), np.int64)
    e[:, 0:3] = plan.grid[rng.integers(0, plan.n, n)]
    coords = rng.integers(0, 32, (n, 3))
    snap = rng.random((n, 3))
    coords[snap < 0.2] = 0
    coords[snap > 0.8] = 31
    e[:, 3:6] = coords
    e[:, 6] = rng.choice([0, 2, 7, 13], n)
    return e


def test_route_edits_partitions_the_tick():
    plan = ShardPlan((3, 3, 3), 3)
    e = make_shard_edits(plan, 500, 3)
    e[7, 0:3] = (9, 9, 9)  # outside the planet: nobody's
    owns = [plan.route_edits(r, e)[0] for r in range(3)]
    total = np.sum(owns, axis=0)
    assert total[7] == 0 and np.all(np.delete(total, 7) == 1), "every edit inside the planet has exactly one owner"
    for r in range(3):
        own, extra = plan.route_edits(r, e)
        assert np.all(plan.owner[extra] == r)
        # brute force: my chunks that are the masked neighbour of an edit on a chunk I do not own
        want = set()
        for row in e[~own]:
            lid = int(plan.linear_id(row[0:3])[0])
            if lid < 0:
                continue
            x, y, z = row[3:6]
            for cond, d in ((x == 0, 3), (x == 31, 4), (y == 0, 2), (y == 31, 0), (z == 0, 1), (z == 31, 5)):
                if cond:
                    nb = int(plan.linear_id(row[0:3] + CHUNK_DIR_OFFSET[d])[0])
                    if nb >= 0 and plan.owner[nb] == r:
                        want.add(nb)
        assert set(extra.tolist()) == want


def _edit_worker(rank, world, port, result_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        count, seed = (3, 3, 3), 7
        plan = ShardPlan(count, world)
        mine = plan.local(rank)
        mine_set = set(map(tuple, mine.tolist()))
        rem = plan.remote_neighbours(rank)

        planet = O.Planet(1.0, 16.0)
        for idx in mine:
            b, ok = O.generate_chunk(seed, count, idx)
            assert ok
            planet.add_chunk(idx, b)

        def exchange_ghosts(first):
            """every rank publishes the blocks of the chunks its peers mirror (the CPU stand-in for tsom_halo_pull)"""
            wanted = [None] * world
            dist.all_gather_object(wanted, {p: ls.tolist() for p, ls in rem.items()})
            answers = {}
            for r in range(world):
                for lid in wanted[r].get(rank, []):
                    answers[(r, lid)] = planet.get_blocks(plan.grid[lid])
            got = [None] * world
            dist.all_gather_object(got, answers)
            for p in range(world):
                for (r, lid), blocks in got[p].items():
                    if r == rank:
                        (planet.add_chunk if first else planet.reset_chunk)(plan.grid[lid], blocks)

        exchange_ghosts(True)
        planet.clear_invalidated()

        # the reference run: one process, whole planet
        whole = O.Planet(1.0, 16.0)
        assert whole.generate(seed, count, nthreads=2)
        whole.clear_invalidated()

        ok_dirty = ok_mesh = True
        nd = 0
        for tick in range(3):
            edits = make_shard_edits(plan, 200, 100 + tick)  # the same list on every rank
            own, extra = plan.route_edits(rank, edits)
            planet.apply_edits(edits[own])
            local = planet.collect_dirty()  # my edited chunks + masked neighbours in my container (ghosts included)
            planet.clear_invalidated()
            dirty = {tuple(k) for k in local.tolist() if tuple(k) in mine_set} | {tuple(plan.grid[l].tolist()) for l in extra}
            exchange_ghosts(False)  # refresh the mirrored neighbours before meshing
            planet.clear_invalidated()

            whole.apply_edits(edits)
            want = {tuple(k) for k in whole.collect_dirty().tolist()}
            whole.clear_invalidated()
            all_dirty = [None] * world
            dist.all_gather_object(all_dirty, sorted(dirty))
            union = set().union(*[set(map(tuple, d)) for d in all_dirty])
            ok_dirty &= union == want and sum(len(d) for d in all_dirty) == len(want)  # complete, and nobody rebuilds another rank's chunk
            for k in sorted(dirty):
                a, b = planet.mesh(k), whole.mesh(k)
                same = (a is None) == (b is None)
                if same and a is not None:
                    same = np.array_equal(a.indices, b.indices) and np.array_equal(a.position, b.position) and np.array_equal(a.uvw, b.uvw)
                ok_mesh &= bool(same)
            nd += len(dirty)
        result_q.put((rank, bool(ok_dirty), bool(ok_mesh), nd))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_edit_ticks_match_single_process():
    """Three ticks of 200 edits on a planet split over two ranks: the owner applies an edit, both sides of a rank boundary rebuild
    what the reference's neighbour mask says, and every rebuilt mesh equals the single-process one."""
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
    assert all(r[1] for r in res), "sharded dirty sets differ from the single-process dirty set"
    assert all(r[2] for r in res), "a mesh rebuilt on a shard differs from the single-process mesh"
    assert all(r[3] > 0 for r in res)
