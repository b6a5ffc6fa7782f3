"""GPU: the Ship container (Ship.cpp) and the batched FlatChunk box colliders (FlatChunk.cpp:13-30,56-153) against the reference's own
code (oracle/_ref, falling back to the restatement where that library is absent)."""
import numpy as np
import pytest

import oracle as O
import thisspaceofmine_b200 as T

pytestmark = pytest.mark.gpu

DIR_OFFSET = np.array([[0, 0, 1], [0, -1, 0], [0, 0, -1], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.int32)


def _ref():
    return O.reference() if O.reference_available() else O


@pytest.mark.parametrize("small", [True, False])
def test_ship_generate_blocks_and_invalidation(ctx, small):
    R = _ref()
    ship, want = T.Ship(ctx, 2.0), R.Ship(2.0)
    try:
        ship.Generate(ctx.blockLibrary, small)
        want.generate(small)
        assert np.array_equal(ship.GetContent([(0, 0, 0)])[0], want.get_blocks((0, 0, 0)))
        # Chunk::Reset() fires no OnReset: the chunk is in the set through its UpdateBlock calls, with an empty neighbour mask
        assert want.take_invalidated() == {(0, 0, 0): 0x80}
        assert ship.CollectInvalidated(clear=True).tolist() == [[0, 0, 0]]
        assert ship.CollectInvalidated(clear=True).tolist() == []
    finally:
        ship.close()


def test_ship_edit_masks_swap_up_and_down(ctx):
    """The same edits on a Ship and on a Planet invalidate different vertical neighbours (Ship.cpp:48-51 vs Planet.cpp:50-53)."""
    R = _ref()
    rng = np.random.default_rng(9)
    chunks = [(0, 0, 0), (0, 1, 0), (0, -1, 0), (1, 0, 0), (0, 0, 1)]
    blocks = (rng.random((len(chunks), 32768)) < 0.4).astype(np.uint16) * 4
    ship, want = T.Ship(ctx, 1.0), R.Ship(1.0)
    planet = T.Planet(ctx, capacity=8)
    try:
        for k, b in zip(chunks, blocks):
            ship.AddChunk(ctx.blockLibrary, k, b)
            want.add_chunk(k, b)
        planet.ResetChunks(np.array(chunks, np.int32), blocks)
        ship.CollectInvalidated(clear=True)
        planet.CollectInvalidated(clear=True)
        want.take_invalidated()
        for tick in range(20):
            n = int(rng.integers(1, 6))
            edits = np.zeros((n, 7), np.int64)
            edits[:, 0:3] = (0, 0, 0)
            edits[:, 3:6] = rng.choice([0, 5, 31], (n, 3))
            edits[:, 6] = rng.choice([0, 4, 13], n)
            ship.UpdateBlocks(edits)
            planet.UpdateBlocks(edits)
            for e in edits:
                want.update_block(e[0:3], e[3:6], int(e[6]))
            masks = want.take_invalidated()
            expect = set()
            for k, m in masks.items():
                expect.add(k)
                for d in range(6):
                    nb = tuple((np.array(k) + DIR_OFFSET[d]).tolist())
                    if (m >> d) & 1 and nb in chunks:
                        expect.add(nb)
            got = {tuple(k) for k in ship.CollectInvalidated(clear=True).tolist()}
            assert got == expect, (tick, edits.tolist())
            pgot = {tuple(k) for k in planet.CollectInvalidated(clear=True).tolist()}
            z = set(edits[:, 5].tolist())
            if 0 in z and 31 not in z:
                assert (0, 1, 0) in got and (0, -1, 0) not in got and (0, -1, 0) in pgot and (0, 1, 0) not in pgot
        assert np.array_equal(ship.GetContent([(0, 0, 0)])[0], want.get_blocks((0, 0, 0)))
    finally:
        ship.close()
        planet.close()


def _hull_rows(ship_py, hull, tile):
    """(offset, boxes) children or a bare box array -> rows like orc_ship_hull_collider: child offset, centre * tile, lengths * tile."""
    if hull is None:
        return np.zeros((0, 9), np.float32)
    children = hull if isinstance(hull, list) else [((0.0, 0.0, 0.0), hull)]
    rows = []
    for off, boxes in children:
        for b in boxes:
            rows.append([*off, *((b[0:3] + b[3:6] / np.float32(2)) * np.float32(tile)), *(b[3:6] * np.float32(tile))])
    return np.array(rows, np.float32).reshape(-1, 9)


def test_ship_hull_collider_matches_the_reference(ctx):
    R = _ref()
    rng = np.random.default_rng(21)
    key = lambda h: h[np.lexsort(h.T[::-1])]
    ship, want = T.Ship(ctx, 2.0), R.Ship(2.0)
    try:
        assert ship.BuildHullCollider() is None and len(want.hull_collider()) == 0
        ship.Generate(ctx.blockLibrary, False)
        want.generate(False)
        got = _hull_rows(ship, ship.BuildHullCollider(), 2.0)
        assert len(got) > 0 and np.array_equal(got, want.hull_collider()), "one chunk: its own compound, boxes in the greedy order"
        for k in [(1, 0, -1), (0, 2, 0)]:
            b = (rng.random(32768) < 0.35).astype(np.uint16) * 4
            ship.AddChunk(ctx.blockLibrary, k, b)
            want.add_chunk(k, b)
        empty = np.zeros(32768, np.uint16)  # a chunk without a colliding block contributes no child (Ship.cpp:72-74)
        ship.AddChunk(ctx.blockLibrary, (3, 3, 3), empty)
        want.add_chunk((3, 3, 3), empty)
        got = _hull_rows(ship, ship.BuildHullCollider(), 2.0)
        ref = want.hull_collider()
        assert len(got) == len(ref) > 0 and np.array_equal(key(got), key(ref))
    finally:
        ship.close()


def test_batched_box_colliders_equal_the_reference_for_every_chunk(ctx):
    """tsom_box_colliders on the whole default planet + its four platforms, one call: same boxes in the same order as
    FlatChunk::BuildCollider chunk by chunk; chunks without a colliding block get no box."""
    R = _ref()
    count = (5, 5, 5)
    grid = T.Planet.chunk_grid(count)
    planet = T.Planet(ctx, capacity=125)
    want = R.Planet()
    try:
        planet.GenerateChunks(ctx.blockLibrary, 42, count)
        assert want.generate(42, count, nthreads=4)
        for d, c in ((T.Direction.Up, (0, 71, 0)), (T.Direction.Back, (3, 5, 70)), (T.Direction.Left, (-70, 1, 2)), (T.Direction.Down, (2, -71, 1))):
            planet.GeneratePlatform(ctx.blockLibrary, d, c)
            want.generate_platform(d, c)
        first, cnt, boxes = planet.BuildBoxColliders(grid)
        assert int(cnt.sum()) == len(boxes) and np.array_equal(first, np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.uint64))
        total = 0
        for i, k in enumerate(grid):
            ref = want.box_collider(k)
            got = boxes[int(first[i]):int(first[i]) + int(cnt[i])]
            assert got.shape == ref.shape and np.array_equal(got, ref), tuple(k)
            total += len(ref)
        assert total > 1000
        # a checkerboard: the worst case of 16,384 single-cell boxes
        z, y, x = np.meshgrid(np.arange(32), np.arange(32), np.arange(32), indexing="ij")
        cb = (((x + y + z) & 1) * 7).astype(np.uint16).reshape(-1)
        one = T.Planet(ctx, capacity=2)
        one.ResetChunks([(0, 0, 0), (5, 5, 5)], np.stack([cb, np.zeros(32768, np.uint16)]))
        f, c, b = one.BuildBoxColliders([(5, 5, 5), (0, 0, 0)])
        assert c.tolist() == [0, 16384] and np.all(b[:, 3:6] == 1.0)
        one.close()
    finally:
        planet.close()
