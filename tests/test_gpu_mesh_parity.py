"""GPU parity: Chunk::BuildMesh on sm_100a (through the C ABI) against the CPU oracle, same seeded inputs."""
import numpy as np
import pytest

import oracle as O
import thisspaceofmine_b200 as T
from helpers import NBLK, assert_mesh_equal, hash_fill, lidx, random_chunk

pytestmark = pytest.mark.gpu


def _pair(ctx, chunks, tile=1.0):
    """Build the same container on both sides: chunks = {idx: blocks or None}."""
    gc = T.ChunkContainer(ctx, max(len(chunks), 1), tile)
    op = O.Planet(tile)
    for idx, b in chunks.items():
        op.add_chunk(idx, b)
        gc.AddChunk(idx, None if b is None else b[None, :])
    return gc, op


def _compare_all(gc, op, indices, **kw):
    batch = gc.BuildMeshes(indices, **{k: v for k, v in kw.items() if k in ("render", "collider", "deformed", "deformRadius", "gravityCenters")})
    inexact = faces = 0
    for i, idx in enumerate(indices):
        g = None if kw.get("gravityCenters") is None else kw["gravityCenters"][i]
        om = op.mesh(idx, deformed=kw.get("deformed", False), deform_radius=kw.get("deformRadius", 0.0), gravity_center=g)
        s = assert_mesh_equal(batch.chunk(i), om, what=f"chunk {tuple(idx)}", check_attr=kw.get("render", True))
        inexact += s["inexact"]
        faces += s["faces"]
    return faces, inexact


def test_single_block(ctx):
    b = np.zeros(NBLK, np.uint16)
    b[lidx(7, 6, 5)] = 7
    gc, op = _pair(ctx, {(0, 0, 0): b})
    faces, inexact = _compare_all(gc, op, [(0, 0, 0)])
    assert faces == 6


def test_empty_and_full_chunk(ctx):
    gc, op = _pair(ctx, {(0, 0, 0): np.zeros(NBLK, np.uint16), (5, 0, 0): np.full(NBLK, 7, np.uint16)})
    batch = gc.BuildMeshes([(0, 0, 0), (5, 0, 0)])
    assert batch.chunk(0) is None  # null mesh, ClientChunkEntities.cpp:161-162
    assert batch.face_counts[1] == 6 * 32 * 32
    _compare_all(gc, op, [(0, 0, 0), (5, 0, 0)])


@pytest.mark.parametrize("p", [0.05, 0.5, 0.95])
def test_random_isolated(ctx, p):
    rng = np.random.default_rng(1234)
    chunks = {(i, 0, 2 * i): random_chunk(rng, p) for i in range(3)}
    chunks = {(2 * i, 0, 0): b for i, (k, b) in enumerate(chunks.items())}  # isolated: no face-adjacent chunk
    gc, op = _pair(ctx, chunks)
    faces, inexact = _compare_all(gc, op, list(chunks))
    assert faces > 0


def test_benchmark_fill_matches_oracle(ctx):
    """The hash fill bench.py uses (BASELINE config 2): same face counts as the oracle."""
    chunks = {(3 * i, 0, 0): hash_fill(1234, i) for i in range(4)}
    gc, op = _pair(ctx, chunks)
    _compare_all(gc, op, list(chunks))


def test_mixed_ids_transparency_and_double_sided(ctx):
    """ids {0,2,3,7,8,9,13}: same-type suppression, transparent neighbours, double-sided glass twins (Chunk.cpp:190-205)."""
    rng = np.random.default_rng(99)
    chunks = {(0, 0, 0): random_chunk(rng, 0.7, ids=(2, 3, 7, 8, 9, 13)), (2, 0, 0): random_chunk(rng, 0.9, ids=(9, 13)), (4, 0, 0): np.full(NBLK, 13, np.uint16)}
    gc, op = _pair(ctx, chunks)
    _compare_all(gc, op, list(chunks))


def test_glass_next_to_stone_and_air(ctx):
    b = np.zeros(NBLK, np.uint16)
    b[lidx(10, 10, 10)] = 13  # glass
    b[lidx(11, 10, 10)] = 7   # stone
    gc, op = _pair(ctx, {(0, 0, 0): b})
    batch = gc.BuildMeshes([(0, 0, 0)])
    # glass: 5 air sides x 2 (double sided), nothing towards stone; stone: 5 air sides + 1 towards glass (transparent)
    assert batch.face_counts[0] == 10 + 6
    _compare_all(gc, op, [(0, 0, 0)])


def test_neighbours_3x3x3(ctx):
    """Halo slabs from all six neighbours, including chunks without content and missing chunks (Chunk.cpp:104-172)."""
    rng = np.random.default_rng(7)
    chunks = {}
    for cx in range(-1, 2):
        for cy in range(-1, 2):
            for cz in range(-1, 2):
                chunks[(cx, cy, cz)] = random_chunk(rng, 0.6, ids=(2, 7, 9, 13))
    chunks[(1, 0, 0)] = None   # exists, no content -> faces towards it are drawn
    del chunks[(0, -1, 0)]     # missing -> faces drawn
    gc, op = _pair(ctx, chunks)
    todo = [k for k, v in chunks.items() if v is not None]
    faces, inexact = _compare_all(gc, op, todo)
    assert faces > 0


def test_tile_size_and_custom_gravity(ctx):
    rng = np.random.default_rng(5)
    chunks = {(1, -2, 3): random_chunk(rng, 0.3, ids=(3, 7))}
    gc, op = _pair(ctx, chunks, tile=0.5)
    _compare_all(gc, op, [(1, -2, 3)])
    g = np.array([[3.25, -100.0, 17.5]], np.float32)
    _compare_all(gc, op, [(1, -2, 3)], gravityCenters=g)


@pytest.mark.parametrize("tile", [1.0, 0.5, 4.0, 0.3])
def test_table_path_is_bit_identical_to_generic_math(ctx, tile):
    """Power-of-two block sizes take the table-driven flat path; it must reproduce the per-face float math bit for bit
    (0.3 is not a power of two: both calls run the generic math)."""
    rng = np.random.default_rng(17)
    chunks = {(-3, 2, 1): random_chunk(rng, 0.5, ids=(1, 3, 7, 13)), (0, -4, 0): random_chunk(rng, 0.2, ids=(1, 3))}
    gc, op = _pair(ctx, chunks, tile=tile)
    idx = list(chunks)
    fast = gc.BuildMeshes(idx)
    meshes_fast = [fast.chunk(i) for i in range(len(idx))]
    gen = gc.BuildMeshes(idx, genericMath=True)
    for i, k in enumerate(idx):
        g = gen.chunk(i)
        assert np.array_equal(meshes_fast[i].vertices.view(np.uint32), g.vertices.view(np.uint32)), f"tile {tile} chunk {k}"
        assert np.array_equal(meshes_fast[i].indices, g.indices)
        assert_mesh_equal(g, op.mesh(k), f"tile {tile} chunk {k}")


def test_collider_reuses_face_list(ctx):
    """DeformedChunk::BuildCollider path: positions only + the same indices, from the same compacted face list."""
    rng = np.random.default_rng(11)
    chunks = {(0, 0, 0): random_chunk(rng, 0.4, ids=(7, 9, 13))}
    gc, op = _pair(ctx, chunks)
    both = gc.BuildMeshes([(0, 0, 0)], render=True, collider=True).chunk(0)
    only = gc.BuildMeshes([(0, 0, 0)], render=False, collider=True).chunk(0)
    om = op.mesh((0, 0, 0))
    assert_mesh_equal(both, om, "render+collider")
    assert only.vertices is None
    assert np.array_equal(only.indices, om.indices)
    assert np.array_equal(only.collider_positions, both.collider_positions)
    np.testing.assert_allclose(only.collider_positions, om.position, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("radius", [0.0, 16.0, 40.0])
def test_deformed_corners(ctx, radius):
    """DeformedChunk::ComputeVoxelCorners + DeformPosition (DeformedChunk.cpp:115-154)."""
    rng = np.random.default_rng(3)
    chunks = {(1, 1, 0): random_chunk(rng, 0.3, ids=(3, 7, 13)), (0, 0, 0): random_chunk(rng, 0.2)}
    gc, op = _pair(ctx, chunks)
    _compare_all(gc, op, list(chunks), deformed=True, deformRadius=radius)
    _compare_all(gc, op, list(chunks), deformed=True, deformRadius=radius, render=False, collider=True)


@pytest.mark.parametrize("tile,radius", [(1.0, 16.0), (1.0, 7.5), (0.5, 3.0), (4.0, 0.0), (1.0, 200.0)])
def test_deformed_table_shortcut_is_bit_identical_to_generic_math(ctx, tile, radius):
    """DeformedChunk blocks whose 8 corners the deformation leaves in place take their normals / uvs from the flat path's table
    (power-of-two block size); the call with genericMath=True evaluates DrawFace for every face.  Same bits either way, near the
    planet's rounded edges (corner chunk), on its flat sides (face chunk) and inside the deformation radius (centre chunk)."""
    rng = np.random.default_rng(23)
    chunks = {(2, 2, 2): random_chunk(rng, 0.4, ids=(1, 3, 7, 13)), (0, 3, 0): random_chunk(rng, 0.3, ids=(1, 3)), (0, 0, 0): random_chunk(rng, 0.1)}
    gc, op = _pair(ctx, chunks, tile=tile)
    idx = list(chunks)
    fast = gc.BuildMeshes(idx, deformed=True, deformRadius=radius)
    meshes_fast = [fast.chunk(i) for i in range(len(idx))]
    gen = gc.BuildMeshes(idx, deformed=True, deformRadius=radius, genericMath=True)
    for i, k in enumerate(idx):
        g = gen.chunk(i)
        assert np.array_equal(meshes_fast[i].vertices.view(np.uint32), g.vertices.view(np.uint32)), f"tile {tile} radius {radius} chunk {k}"
        assert np.array_equal(meshes_fast[i].indices, g.indices)
        assert_mesh_equal(g, op.mesh(k, deformed=True, deform_radius=radius), f"tile {tile} radius {radius} chunk {k}")
    # own centres (gravity and deformation): far along one axis on the 2^-10 grid (whole chunk undeformed), far but off the grid
    # (per-corner check), and close by (really deformed)
    centers = np.array([[-300.25, 3.0, 8.5], [0.3, 411.7, -2.9], [10.0, 12.0, 40.0]], np.float32)
    fast = gc.BuildMeshes(idx, deformed=True, deformRadius=radius, gravityCenters=centers)
    meshes_fast = [fast.chunk(i) for i in range(len(idx))]
    gen = gc.BuildMeshes(idx, deformed=True, deformRadius=radius, gravityCenters=centers, genericMath=True)
    for i, k in enumerate(idx):
        g = gen.chunk(i)
        assert np.array_equal(meshes_fast[i].vertices.view(np.uint32), g.vertices.view(np.uint32)), f"own centre: tile {tile} radius {radius} chunk {k}"
        assert_mesh_equal(g, op.mesh(k, deformed=True, deform_radius=radius, gravity_center=centers[i]), f"own centre: tile {tile} radius {radius} chunk {k}")


def test_deformed_table_shortcut_random_centres_and_radii(ctx):
    """Forty random (radius, centre) pairs — on and off the 2^-10 grid, near and far, radii with r * (1 / r) != 1 among them — over chunks
    at the corner, on an edge, on a face and in the middle of a planet: the shortcut run and the generic-math run write the same bits."""
    rng = np.random.default_rng(41)
    keys = [(2, 2, 2), (2, -2, 0), (0, 3, 0), (0, 0, 0), (-3, 1, 0), (1, 0, -3)]
    chunks = {k: random_chunk(rng, 0.35, ids=(1, 3, 7, 13)) for k in keys}
    gc, op = _pair(ctx, chunks)
    idx = list(chunks)
    taken = 0
    for trial in range(40):
        radius = float(rng.choice([0.0, 1.0, 3.0, 7.5, 16.0, 41.0, 49.0, 0.3, 100.0 / 3.0, 12.125]))
        kw = dict(deformed=True, deformRadius=radius)
        if trial % 2:
            c = rng.normal(0, 150, (len(idx), 3))
            c = np.round(c * 4) / 4 if trial % 4 == 1 else c  # quarter-block grid or arbitrary
            kw["gravityCenters"] = c.astype(np.float32)
        fast = gc.BuildMeshes(idx, **kw)
        meshes_fast = [fast.chunk(i) for i in range(len(idx))]
        gen = gc.BuildMeshes(idx, genericMath=True, **kw)
        meshes_gen = [gen.chunk(i) for i in range(len(idx))]  # a batch's views die with the next BuildMeshes call
        for i, k in enumerate(idx):
            g = meshes_gen[i]
            assert np.array_equal(meshes_fast[i].vertices.view(np.uint32), g.vertices.view(np.uint32)), f"trial {trial} radius {radius} chunk {k}"
            assert np.array_equal(meshes_fast[i].indices, g.indices)
        # how often the geometry really was left in place (the shortcut is exercised, not just skipped)
        if "gravityCenters" not in kw:
            flat = gc.BuildMeshes(idx)
            taken += sum(int(np.array_equal(flat.chunk(i).normal, meshes_gen[i].normal)) for i in range(len(idx)))
    assert taken > 0


def test_batch_order_and_views(ctx):
    """One batched call: per-chunk views are dense, in job order, and each equals the single-chunk result."""
    rng = np.random.default_rng(21)
    chunks = {(2 * i, 0, 0): random_chunk(rng, 0.1 * (i + 1)) for i in range(6)}
    gc, op = _pair(ctx, chunks)
    order = list(chunks)[::-1]
    batch = gc.BuildMeshes(order, ordered=True)  # TSOM_MESH_ORDERED: ranges in job order
    assert batch.first_faces[0] == 0
    assert np.array_equal(batch.first_faces[1:], np.cumsum(batch.face_counts.astype(np.uint64))[:-1])
    for i, idx in enumerate(order):
        assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"{idx}")
    # default: first come, first served — the ranges still tile [0, total) without gaps or overlaps
    batch = gc.BuildMeshes(order)
    o = np.argsort(batch.first_faces, kind="stable")
    nz = [i for i in o if batch.face_counts[i] > 0]
    ends = 0
    for i in nz:
        assert batch.first_faces[i] == ends
        ends += int(batch.face_counts[i])
    assert ends == batch.total_faces == ctx.arenas()[3]
    for i, idx in enumerate(order):
        assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"{idx}")


def test_unknown_chunk_is_an_error(ctx):
    gc, _ = _pair(ctx, {(0, 0, 0): np.zeros(NBLK, np.uint16)})
    with pytest.raises(T.TsomError):
        gc.BuildMeshes([(9, 9, 9)])


def test_hidden_faces_appear_after_edits_resets_and_removals(ctx):
    """3x3x3 solid stone chunks (the centre has no face at all) and an all-Empty chunk far away; then an edit, a reset and a
    removal that each uncover hidden faces.  Guards any per-chunk shortcut ("this chunk cannot have a face") against going stale."""
    stone = np.full(NBLK, 7, np.uint16)
    chunks = {(x, y, z): stone.copy() for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)}
    chunks[(5, 5, 5)] = np.zeros(NBLK, np.uint16)
    gc, op = _pair(ctx, chunks)
    all_idx = list(chunks)

    def check():
        present = [i for i in all_idx if i in chunks]
        f, _ = _compare_all(gc, op, present)
        return f

    f0 = check()
    assert check() == f0  # meshing twice gives the same result
    centre = all_idx.index((0, 0, 0))
    assert gc.BuildMeshes(all_idx).face_counts[centre] == 0

    # 1. break a block of a neighbour where it touches the centre: the centre chunk gains exactly one face
    gc.UpdateBlock((1, 0, 0), (0, 5, 5), 0)
    op.update_block((1, 0, 0), (0, 5, 5), 0)
    check()
    assert gc.BuildMeshes(all_idx).face_counts[centre] == 1

    # 2. place a block in the all-Empty chunk
    gc.UpdateBlock((5, 5, 5), (3, 3, 3), 13)
    op.update_block((5, 5, 5), (3, 3, 3), 13)
    check()
    assert gc.BuildMeshes([(5, 5, 5)]).face_counts[0] == 12  # glass: double-sided

    # 3. reset a neighbour to glass (transparent): the centre's whole Down side becomes visible
    glass = np.full(NBLK, 13, np.uint16)
    gc.ResetChunks([(0, -1, 0)], glass[None, :])
    op.reset_chunk((0, -1, 0), glass)
    chunks[(0, -1, 0)] = glass
    check()
    assert gc.BuildMeshes(all_idx).face_counts[centre] == 1 + 32 * 32

    # 4. remove a neighbour: no neighbour chunk => faces drawn (Chunk.cpp:165-166)
    gc.RemoveChunk((0, 0, 1))
    del chunks[(0, 0, 1)]
    op2 = O.Planet()
    for idx in chunks:
        op2.add_chunk(idx, op.get_blocks(idx))
    op = op2
    present = [i for i in all_idx if i in chunks]
    batch = gc.BuildMeshes(present)
    for i, idx in enumerate(present):
        assert_mesh_equal(batch.chunk(i), op.mesh(idx), what=f"after removal {idx}")
    assert batch.face_counts[present.index((0, 0, 0))] == 1 + 2 * 32 * 32


@pytest.mark.parametrize("ordered", [False, True])
def test_views_are_valid_when_the_arenas_had_to_grow(ordered):
    """A fresh context has tiny arenas: the first call only counts, grows them and runs again.  With first-come placement the
    second run places the chunks elsewhere than the first, so the views must come from the second run."""
    c2 = T.Context(0)
    try:
        rng = np.random.default_rng(99)
        chunks = {(2 * i, 0, 0): random_chunk(rng, 0.5) for i in range(40)}
        gc, op = _pair(c2, chunks)
        keys = list(chunks)
        batch = gc.BuildMeshes(keys, ordered=ordered)
        assert c2.timings()["mesh_reemit"]
        for i, idx in enumerate(keys):
            assert_mesh_equal(batch.chunk(i), op.mesh(idx), f"{idx}")
        gc.close()
    finally:
        c2.close()


def test_concurrent_callers_on_one_context(ctx):
    """The reference calls BuildMesh from many TaskScheduler workers at once (ClientChunkEntities.cpp:207-241).  Every C entry point
    locks the context; a worker brackets "mesh + read the arenas" with tsom_ctx_lock.  6 threads, each with its own container, 12 rounds:
    every mesh equals the one computed alone."""
    import threading

    rng = np.random.default_rng(5)
    pairs = []
    for t in range(6):
        chunks = {(2 * i, t, 0): random_chunk(rng, 0.3 + 0.1 * t) for i in range(3)}
        pairs.append((*_pair(ctx, chunks), list(chunks)))
    want = [[op.mesh(k) for k in keys] for _, op, keys in pairs]
    errors = []

    def worker(t):
        gc, _, keys = pairs[t]
        try:
            for _ in range(12):
                with ctx.locked():
                    batch = gc.BuildMeshes(keys, collider=True)
                    got = [batch.chunk(i) for i in range(len(keys))]
                for i, k in enumerate(keys):
                    assert_mesh_equal(got[i], want[t][i], f"thread {t} chunk {k}")
        except Exception as e:  # noqa: BLE001 - reported below
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for gc, _, _ in pairs:
        gc.close()
    assert not errors, errors[0]


@pytest.mark.parametrize("mode", ["flat", "deformed", "collider_only"])
def test_aabb_is_the_box_of_the_emitted_positions(ctx, mode):
    """Nz::StaticMesh::GenerateAABB (ClientChunkEntities.cpp:168) on the device: min / max of exactly the positions in the arena, and
    within 1e-5 of the box of the oracle's positions; zeros for a chunk without faces."""
    rng = np.random.default_rng(11)
    chunks = {(i, 0, 0): random_chunk(rng, 0.02 + 0.2 * i) for i in range(4)}
    chunks[(9, 9, 9)] = np.zeros(32768, np.uint16)  # no faces
    gc, op = _pair(ctx, chunks)
    keys = list(chunks)
    kw = dict(render=mode != "collider_only", collider=mode == "collider_only", deformed=mode == "deformed", deformRadius=16.0 if mode == "deformed" else 0.0)
    batch = gc.BuildMeshes(keys, **kw)
    boxes = batch.aabbs()
    assert boxes.shape == (len(keys), 6)
    for i, k in enumerate(keys):
        m = batch.chunk(i)
        if m is None:
            assert not boxes[i].any()
            continue
        pos = m.vertices[:, :3] if mode != "collider_only" else m.collider_positions
        assert np.array_equal(boxes[i, :3], pos.min(0)) and np.array_equal(boxes[i, 3:], pos.max(0)), f"chunk {k}"
        om = op.mesh(k, deformed=mode == "deformed", deform_radius=kw["deformRadius"])
        want = np.concatenate([om.position.min(0), om.position.max(0)])
        assert np.allclose(boxes[i], want, rtol=1e-5, atol=1e-5), f"chunk {k}"
    gc.close()


@pytest.mark.parametrize("scenario", ["neighbours_mixed_ids", "deformed", "tile_and_gravity"])
def test_meshes_against_the_references_own_code(ctx, scenario):
    """CUDA vs oracle/_ref (the reference's own Chunk.cpp / FlatChunk.cpp / DeformedChunk.cpp compiled from its sources, shipped prebuilt
    to the GPU box) instead of the restatement: halos with missing / content-less neighbours and transparent, double-sided and
    non-colliding blocks; deformed corners; a custom block size and gravity centre."""
    if not O.reference_available():
        pytest.skip("oracle/_ref was not built (needs /root/reference at build time)")
    R = O.reference()
    rng = np.random.default_rng(21)
    tile = 0.5 if scenario == "tile_and_gravity" else 1.0
    chunks = {}
    for cx in range(-1, 2):
        for cy in range(0, 2):
            for cz in range(-1, 1):
                chunks[(cx, cy, cz)] = random_chunk(rng, 0.55, ids=(2, 3, 7, 8, 9, 13))
    chunks[(1, 0, 0)] = None
    del chunks[(0, 1, -1)]
    gc = T.ChunkContainer(ctx, len(chunks), tile)
    rp = R.Planet(tile)
    for idx, b in chunks.items():
        rp.add_chunk(idx, b)
        gc.AddChunk(idx, None if b is None else b[None, :])
    todo = [k for k, v in chunks.items() if v is not None]
    kw = {}
    if scenario == "deformed":
        kw = dict(deformed=True, deformRadius=16.0)
    if scenario == "tile_and_gravity":
        kw = dict(gravityCenters=np.tile(np.array([[3.25, -100.0, 17.5]], np.float32), (len(todo), 1)))
    faces, _ = _compare_all(gc, rp, todo, **kw)
    assert faces > 10000
    gc.close()


@pytest.mark.parametrize("tile", [1.0, 0.5])
def test_compute_voxel_corners_query(ctx, tile):
    """Chunk::ComputeVoxelCorners (the block-picking helper) as a batched device query: FlatChunk and DeformedChunk providers, default and
    caller-supplied deformation centres, against the restatement and — when built — the reference's own code; flat corners exactly,
    deformed ones within 1e-5."""
    rng = np.random.default_rng(8)
    chunk_keys = [(0, 0, 0), (1, -2, 3), (-2, 0, 1)]
    gc = T.ChunkContainer(ctx, len(chunk_keys), tile)
    backends = [O] + ([O.reference()] if O.reference_available() else [])
    planets = []
    for B in backends:
        p = B.Planet(tile)
        for k in chunk_keys:
            p.add_chunk(k, np.zeros(32768, np.uint16))
        planets.append(p)
    for k in chunk_keys:
        gc.AddChunk(k, None)
    n = 64
    ci = np.array([chunk_keys[i] for i in rng.integers(0, len(chunk_keys), n)], np.int32)
    loc = rng.integers(0, 32, (n, 3)).astype(np.uint32)
    loc[0], loc[1] = (0, 0, 0), (31, 31, 31)
    centers = rng.normal(0, 40, (n, 3)).astype(np.float32)
    for mode, kw, okw in (("flat", {}, {}),
                          ("deformed", dict(deformed=True, deformRadius=16.0), dict(deformed=True, deform_radius=16.0)),
                          ("deformed, own centres", dict(deformed=True, deformRadius=7.5, deformCenters=centers), dict(deformed=True, deform_radius=7.5))):
        got = gc.ComputeVoxelCorners(ci, loc, **kw)
        assert got.shape == (n, 8, 3)
        for p in planets:
            for i in range(n):
                extra = dict(okw, deform_center=centers[i]) if "own" in mode else okw
                want = p.voxel_corners(ci[i], loc[i], **extra)
                if mode == "flat":
                    assert np.array_equal(got[i], want), (mode, i)
                else:
                    assert np.allclose(got[i], want, rtol=1e-5, atol=1e-5), (mode, i)
    gc.close()
