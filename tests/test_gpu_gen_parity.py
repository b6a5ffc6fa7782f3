"""GPU parity: Planet::GenerateChunk(s) / GeneratePlatform / edits on sm_100a against the CPU oracle — block ids bit-exact."""
import numpy as np
import pytest

import oracle as O
import thisspaceofmine_b200 as T
from helpers import assert_mesh_equal

pytestmark = pytest.mark.gpu

# ServerPlanetEnvironment.cpp:60-63: the four start platforms of the default planet
PLATFORMS = [(T.Direction.Right, (65, -18, -39)), (T.Direction.Back, (-34, 2, 53)), (T.Direction.Front, (22, -35, -59)), (T.Direction.Down, (23, -62, 26))]


def _planets(ctx, seed, count, platforms=False):
    lib = ctx.blockLibrary
    n = int(np.prod(count))
    gp = T.Planet(ctx, 1.0, 16.0, 9.81, capacity=n)
    gp.GenerateChunks(lib, seed, count)
    op = O.Planet(1.0, 16.0)
    assert op.generate(seed, count, nthreads=8)
    if platforms:
        for d, c in PLATFORMS:
            gp.GeneratePlatform(lib, d, c)
            op.generate_platform(d, c)
    return gp, op


def _assert_blocks_equal(gp, op, grid):
    got = gp.GetContent(grid)
    for i, idx in enumerate(grid):
        want = op.get_blocks(idx)
        if not np.array_equal(got[i], want):
            bad = np.flatnonzero(got[i] != want)
            k = int(bad[0])
            raise AssertionError(f"chunk {tuple(idx)}: {len(bad)} ids differ, first at local {(k % 32, (k // 32) % 32, k // 1024)}: gpu {got[i][k]} oracle {want[k]}")


@pytest.mark.parametrize("seed,count", [(42, (5, 5, 5)), (42, (3, 3, 3)), (7, (1, 1, 1)), (123456789, (3, 5, 5)), (0xFFFFFFFE, (7, 3, 3))])
def test_generate_chunks_bit_exact(ctx, seed, count):
    gp, op = _planets(ctx, seed, count)
    _assert_blocks_equal(gp, op, T.Planet.chunk_grid(count))


@pytest.mark.parametrize("seed,count", [(42, (9, 9, 9)), (2024, (11, 9, 9))])
def test_template_and_masked_template_chunks_bit_exact(ctx, seed, count):
    """Planets large enough that a chunk in which EVERY block draws also sits in the terrain band: the generator's masked template copy
    (template chunk AND NOT "above a column start"), which the 5^3 fixtures never reach (there the drawing chunks and the carved chunks are
    disjoint).  Every chunk's ids against the oracle; the classes the generator saw are checked to be the ones this test is about."""
    gp, op = _planets(ctx, seed, count)
    grid = T.Planet.chunk_grid(count)
    _assert_blocks_equal(gp, op, grid)
    # some full-draw chunk must have been carved (stone / stone_mossy chunk with Empty blocks above the terrain) and some must not
    got = gp.GetContent(grid)
    lib = ctx.blockLibrary
    stone, mossy = lib.GetBlockIndex("stone"), lib.GetBlockIndex("stone_mossy")
    only_draws = np.array([np.isin(b, (0, stone, mossy)).all() and (b != 0).any() for b in got])
    carved = only_draws & np.array([(b == 0).any() for b in got])
    not_shaft = ~((grid[:, 0] == 0) & (grid[:, 2] == 0))
    assert (carved & not_shaft).sum() > 10, "no masked template chunk in this planet"
    assert (only_draws & ~carved & not_shaft).sum() > 10, "no plain template chunk in this planet"
    # and every mesh: the summaries the generator wrote let the mesher skip chunks (all opaque with all-opaque neighbours, all Empty) and
    # know the row masks of all-opaque chunks up front — each of those shortcuts has to give what the oracle gives
    batch = gp.BuildMeshes(grid, render=True, collider=True)
    skipped = with_faces = 0
    for i, idx in enumerate(grid):
        st = assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"chunk {tuple(idx)}")
        with_faces += st["faces"] > 0
        skipped += st["faces"] == 0 and only_draws[i] and not carved[i]
    assert skipped > 20 and with_faces > 300, (skipped, with_faces)
    gp.close()


def test_bernoulli_tie_where_the_first_engine_output_decides(ctx):
    """std::bernoulli_distribution(0.9) over std::minstd_rand burns two engine outputs u1, u2 per draw and is decided by u2 alone
    (stone iff u2 - 1 < 1932735281) except when u2 - 1 == 1932735281, where u1 decides (about once in 2^31 draws; the kernel's slow
    path).  Since u2 = x0 * 48271^(2n+2) mod (2^31 - 1), a seed can be solved for that puts the tie on a chosen draw n: here draw
    5367 of every chunk whose indices sum to 0 — (0,0,0) (every block draws: the full-draw loop), (1,0,-1) (31 of 32 rows draw, whole
    8-block groups) and (-1,0,1) (31 of 32 columns draw: the mixed group).  With u1 = u2 / 48271 the tie resolves to stone_mossy, the
    opposite of what the sign of the u2 comparison alone would give."""
    M, A, K = 2147483647, 48271, 1932735282
    n = 5 * 1024 + 7 * 32 + 20
    seed = K * pow(pow(A, -1, M), 2 * n + 2, M) % M
    assert seed == 786579212
    x = seed
    for _ in range(2 * n + 2):
        x = x * A % M
    assert x == K  # the tie is really there
    count = (5, 5, 5)
    gp, op = _planets(ctx, seed, count)
    lib = ctx.blockLibrary
    assert op.get_blocks((0, 0, 0))[n] == lib.GetBlockIndex("stone_mossy")
    _assert_blocks_equal(gp, op, T.Planet.chunk_grid(count))


def test_default_planet_with_platforms_and_meshes(ctx):
    """BASELINE config 1/3: seed 42, 5x5x5, the 4 GeneratePlatform scripts, then every chunk meshed with real neighbours."""
    count = (5, 5, 5)
    gp, op = _planets(ctx, 42, count, platforms=True)
    grid = T.Planet.chunk_grid(count)
    _assert_blocks_equal(gp, op, grid)
    batch = gp.BuildMeshes(grid, render=True, collider=True)
    total = inexact = 0
    for i, idx in enumerate(grid):
        s = assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"chunk {tuple(idx)}")
        total += s["faces"]
        inexact += s["inexact"]
    assert total == batch.total_faces and total > 50000
    # dirty set after generation + platforms = every chunk (OnReset -> DirectionMask_All)
    assert np.array_equal(gp.CollectInvalidated(), op.collect_dirty())


def test_default_planet_against_the_references_own_code(ctx):
    """The same as above, but the CPU side is oracle/_ref: the reference's own Planet.cpp / Chunk.cpp / FlatChunk.cpp compiled from its
    sources (shipped prebuilt to the GPU box) instead of the restatement — CUDA vs the reference directly, every chunk of the default
    planet with its four platforms: block ids, dirty set, and every mesh (indices bit-exact, vertices within 1e-5)."""
    if not O.reference_available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    R = O.reference()
    count = (5, 5, 5)
    lib = ctx.blockLibrary
    gp = T.Planet(ctx, 1.0, 16.0, 9.81, capacity=125)
    gp.GenerateChunks(lib, 42, count)
    rp = R.Planet(1.0, 16.0)
    assert rp.generate(42, count, nthreads=8)
    for d, c in PLATFORMS:
        gp.GeneratePlatform(lib, d, c)
        rp.generate_platform(d, c)
    grid = T.Planet.chunk_grid(count)
    _assert_blocks_equal(gp, rp, grid)
    batch = gp.BuildMeshes(grid, render=True, collider=True)
    total = 0
    for i, idx in enumerate(grid):
        total += assert_mesh_equal(batch.chunk(i), rp.mesh(idx), f"chunk {tuple(idx)}")["faces"]
    assert total == batch.total_faces and total > 50000
    assert np.array_equal(gp.CollectInvalidated(), rp.collect_dirty())
    # a tick of edits on both sides
    rng = np.random.default_rng(4)
    edits = np.zeros((2000, 7), np.int64)
    edits[:, 0:3] = grid[rng.integers(0, len(grid), 2000)]
    edits[:, 3:6] = rng.integers(0, 32, (2000, 3))
    edits[:, 6] = rng.choice([0, 2, 7, 13], 2000)
    rp.clear_invalidated()
    gp.UpdateBlocks(edits)
    rp.apply_edits(edits)
    dirty = gp.CollectInvalidated()
    assert np.array_equal(dirty, rp.collect_dirty())
    batch = gp.BuildMeshes(dirty, render=True)
    for i, idx in enumerate(dirty):
        assert_mesh_equal(batch.chunk(i), rp.mesh(idx), f"after edits, chunk {tuple(idx)}")
    gp.close()


def test_even_chunk_count_is_rejected(ctx):
    """chunkCount 4 puts blocks beyond +-maxHeight: the reference asserts (Nz::SafeCaster, Planet.cpp:199); we refuse."""
    gp = T.Planet(ctx, capacity=64)
    with pytest.raises(T.TsomError) as e:
        gp.GenerateChunks(ctx.blockLibrary, 42, (4, 4, 4))
    assert e.value.code == 3
    _, ok = O.generate_chunk(42, (4, 4, 4), (-2, 0, 0))
    assert not ok


def test_crossed_axes_reject_unequal_y_z(ctx):
    """depth crosses Y and Z (Planet.cpp:199-203), so chunkCount.y != chunkCount.z leaves blocks beyond maxHeight: refused on
    both sides."""
    gp = T.Planet(ctx, capacity=15)
    with pytest.raises(T.TsomError) as e:
        gp.GenerateChunks(ctx.blockLibrary, 1, (3, 5, 1))
    assert e.value.code == 3
    assert not O.Planet().generate(1, (3, 5, 1))


def test_regen_single_chunk(ctx):
    """/regenchunk: Planet::GenerateChunk on one chunk of an existing planet (PlayerSessionHandler.cpp:290)."""
    gp, op = _planets(ctx, 42, (3, 3, 3))
    gp.UpdateBlocks(np.array([[1, 1, 0, 3, 4, 5, 13]]))
    gp.GenerateChunk(ctx.blockLibrary, [(1, 1, 0)], 42, (3, 3, 3))
    _assert_blocks_equal(gp, op, np.array([[1, 1, 0]], np.int32))


def test_edit_stream_and_dirty_set(ctx):
    """BASELINE config 4 in small: random UpdateBlock ops per tick, dirty set = edited chunks + masked neighbours, remesh them."""
    count = (3, 3, 3)
    gp, op = _planets(ctx, 42, count)
    grid = T.Planet.chunk_grid(count)
    gp.CollectInvalidated()
    op.collect_dirty()
    rng = np.random.default_rng(2025)
    for tick in range(4):
        n = 2000
        edits = np.zeros((n, 7), np.int64)
        edits[:, 0:3] = grid[rng.integers(0, len(grid), n)]
        # bias towards boundaries so the neighbour-mask rule (Planet.cpp:39-58) is exercised, and repeat blocks so order matters
        edits[:, 3:6] = rng.choice([0, 1, 15, 30, 31], size=(n, 3))
        edits[:, 6] = rng.choice([0, 2, 7, 13], size=n)
        if tick == 3:
            edits[:, 0:3] = (9, 9, 9)  # edits to a chunk that does not exist are skipped
        gp.UpdateBlocks(edits)
        op.apply_edits(edits)
        dg, do = gp.CollectInvalidated(), op.collect_dirty()
        assert np.array_equal(dg, do), f"tick {tick}: dirty sets differ"
        if len(dg):
            batch = gp.BuildMeshes(dg)
            for i, idx in enumerate(dg):
                assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"tick {tick} chunk {tuple(idx)}")
    _assert_blocks_equal(gp, op, grid)


def test_interior_edit_invalidates_only_its_chunk(ctx):
    gp, op = _planets(ctx, 42, (3, 3, 3))
    gp.CollectInvalidated()
    gp.UpdateBlock((0, 0, 0), (5, 5, 5), 7)
    assert gp.CollectInvalidated().tolist() == [[0, 0, 0]]
    gp.UpdateBlock((0, 0, 0), (0, 5, 31), 7)  # x == 0 -> Left, z == 31 -> Up
    assert gp.CollectInvalidated().tolist() == [[-1, 0, 0], [0, 0, 0], [0, 1, 0]]


def test_box_collider(ctx):
    """FlatChunk::BuildCollider greedy merge (FlatChunk.cpp:56-153): same boxes, same order."""
    gp, op = _planets(ctx, 42, (3, 3, 3))
    for idx in [(0, 1, 0), (1, 1, 1), (0, 0, 0), (-1, 0, 1)]:
        got, want = gp.BuildBoxCollider(idx), op.box_collider(idx)
        assert got.shape == want.shape and np.array_equal(got, want), f"{idx}: {got.shape} vs {want.shape}"


def test_halo_pull_between_two_shards(ctx):
    """Multi-GPU path on one GPU: the planet split into two containers ("ranks"); each pulls the other's boundary slabs
    (the P2P kernel, here over local pointers) and must produce the meshes of the unsplit planet."""
    import ctypes as C
    from thisspaceofmine_b200 import _lib as L

    count = (3, 3, 3)
    whole, op = _planets(ctx, 42, count)
    grid = T.Planet.chunk_grid(count)
    content = whole.GetContent(grid)
    half = len(grid) // 2 + 1
    parts = [(grid[:half], content[:half]), (grid[half:], content[half:])]
    shards = []
    for g, b in parts:
        s = T.ChunkContainer(ctx, len(g))
        s.ResetChunks(g, b)
        shards.append(s)
    owner = {tuple(idx): (r, i) for r, (g, _) in enumerate(parts) for i, idx in enumerate(g)}
    for r, s in enumerate(shards):
        other = 1 - r
        L.check(L.lib().tsom_peer_attach_local(s._h, other, shards[other]._h))
        rem = []
        for idx in parts[r][0]:
            for d in range(6):
                n = tuple(int(v) for v in idx + T.host.s_chunkDirOffset[d])
                if n in owner and owner[n][0] == other:
                    rem.append(n)
        rem = sorted(set(rem))
        if rem:
            ri = np.array(rem, np.int32)
            slots = np.zeros(len(rem), np.uint32)
            for k, n in enumerate(rem):
                sl = C.c_uint32()
                L.check(L.lib().tsom_chunk_slot(shards[other]._h, ri[k].ctypes.data_as(C.POINTER(C.c_int32)), C.byref(sl)))
                slots[k] = sl.value
            ranks = np.full(len(rem), other, np.int32)
            L.check(L.lib().tsom_container_add_remote_chunks(s._h, ri.ctypes.data_as(C.POINTER(C.c_int32)), slots.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                             ranks.ctypes.data_as(C.POINTER(C.c_int32)), len(rem)))
    # both containers share one context (one stream): publish all, then pull all (include/tsom_b200.h, tsom_halo_pull)
    for s in shards:
        s.PublishHalo()
    for s in shards:
        s.PullHalos()
    for r, s in enumerate(shards):
        batch = s.BuildMeshes(parts[r][0])
        for i, idx in enumerate(parts[r][0]):
            assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"rank {r} chunk {tuple(idx)}")
