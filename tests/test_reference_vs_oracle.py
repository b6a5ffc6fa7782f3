"""The restated CPU oracle against THE REFERENCE'S OWN CODE.

``oracle/_ref/libtsom_ref.so`` is src/CommonLib/{Chunk,FlatChunk,DeformedChunk,Planet,BlockLibrary,ChunkContainer}.cpp compiled
unmodified from /root/reference over the stand-in third-party headers of oracle/ref_shim (oracle/ref_shim/Makefile).  These tests
pin the oracle's structure — block table, index math, generation passes and RNG consumption, the platform script, neighbour /
visibility / double-sided rules, emission order, greedy box merge, the neighbour-mask rule — to the reference's real code.
Bit-exact everywhere, floats included (same arithmetic, -ffp-contract=off on both sides).

Skipped only when the library is neither prebuilt nor buildable (no /root/reference)."""
import numpy as np
import pytest

import oracle as O
from helpers import NBLK, lidx, random_chunk

pytestmark = pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref is not built and /root/reference is absent")


@pytest.fixture(scope="module")
def R():
    return O.reference()


def grid(cc):
    return [(x - cc[0] // 2, y - cc[1] // 2, z - cc[2] // 2) for z in range(cc[2]) for y in range(cc[1]) for x in range(cc[0])]


def assert_same_mesh(a, b, what):
    assert (a is None) == (b is None), what
    if a is None:
        return 0
    assert a.face_count == b.face_count, f"{what}: faces {a.face_count} != {b.face_count}"
    assert np.array_equal(a.indices, b.indices), f"{what}: indices"
    assert np.array_equal(a.position, b.position), f"{what}: positions"
    if a.normal is not None:
        assert np.array_equal(a.normal, b.normal), f"{what}: normals"
    if a.uvw is not None:
        assert np.array_equal(a.uvw, b.uvw), f"{what}: uvw"
    return a.face_count


def test_reference_kats_on_reference_code(R):
    """tests/UnitTests/Common/ChunkTests.cpp:17-43 executed by the reference's own ChunkContainer (sanity of the build)."""
    H = 16
    assert R.get_block_indices((0, 0, 0), (0, 0, 0)) == (-H, -H, -H)
    assert R.get_block_indices((-1, 0, 0), (0, 0, 0)) == (-H * 3, -H, -H)
    assert R.get_block_indices((0, 1, 0), (0, 0, 0)) == (-H, H, -H)
    for chunk, local in [((1, 0, 0), (1, 0, 0)), ((-1, 42, 2), (14, 10, 12)), ((3, 42, -2), (31, 31, 31))]:
        assert R.get_chunk_indices_by_block_indices(R.get_block_indices(chunk, local)) == (chunk, local)


def test_index_math_agrees_everywhere(R):
    rng = np.random.default_rng(0)
    for _ in range(300):
        chunk = tuple(int(v) for v in rng.integers(-50, 50, 3))
        local = tuple(int(v) for v in rng.integers(0, 32, 3))
        g = O.get_block_indices(chunk, local)
        assert g == R.get_block_indices(chunk, local)
        assert O.get_chunk_indices_by_block_indices(g) == R.get_chunk_indices_by_block_indices(g) == (chunk, local)


def test_block_library_is_the_references(R):
    a, b = O.block_library(), R.block_library()
    assert len(a) == len(b) == 14
    for x, y in zip(a, b):
        assert x[0] == y[0] and list(x[1]) == list(y[1]) and x[2:] == y[2:], (x, y)


def test_direction_from_normal(R):
    rng = np.random.default_rng(1)
    v = rng.normal(size=(500, 3)).astype(np.float32)
    v[:50] = np.round(v[:50])  # exact ties
    for n in v:
        assert O.direction_from_normal(n) == R.direction_from_normal(n)


@pytest.mark.parametrize("seed,cc", [(42, (5, 5, 5)), (7, (3, 3, 3)), (1234567, (7, 7, 7))])
def test_generation_is_bit_identical(R, seed, cc):
    po, pr = O.Planet(), R.Planet()
    assert po.generate(seed, cc, nthreads=4)
    assert pr.generate(seed, cc, nthreads=4)
    nonempty = 0
    for idx in grid(cc):
        a, b = po.get_blocks(idx), pr.get_blocks(idx)
        assert np.array_equal(a, b), f"seed {seed} chunk {idx}"
        nonempty += int((a != 0).sum())
    assert nonempty > 0


def test_generation_of_unsupported_planets_fails_on_both(R):
    """An even chunkCount makes `depth` negative for the outermost blocks: Nz::SafeCaster asserts in the reference (Planet.cpp:199)."""
    _, ok_o = O.generate_chunk(42, (4, 4, 4), (-2, -2, -2))
    _, ok_r = R.generate_chunk(42, (4, 4, 4), (-2, -2, -2))
    assert not ok_o and not ok_r
    # non-cubic planets: the y/z terms of `depth` are crossed (Planet.cpp:201-202), so the longer axis goes negative
    po, pr = O.Planet(), R.Planet()
    assert not po.generate(1, (3, 5, 7), nthreads=2)
    assert not pr.generate(1, (3, 5, 7), nthreads=2)


def make_default_pair(R, platforms=True):
    cc = (5, 5, 5)
    po, pr = O.Planet(), R.Planet()
    assert po.generate(42, cc, nthreads=4) and pr.generate(42, cc, nthreads=4)
    if platforms:
        for d, c in [(O.UP, (0, 70, 0)), (O.DOWN, (3, -68, -5)), (O.LEFT, (-67, 4, 2)), (O.BACK, (5, -3, 66))]:
            po.generate_platform(d, c)
            pr.generate_platform(d, c)
    return cc, po, pr


def test_platform_script_is_bit_identical(R):
    cc, po, pr = make_default_pair(R)
    for idx in grid(cc):
        assert np.array_equal(po.get_blocks(idx), pr.get_blocks(idx)), idx


def test_default_planet_meshes_are_bit_identical(R):
    cc, po, pr = make_default_pair(R)
    faces = 0
    for idx in grid(cc):
        faces += assert_same_mesh(po.mesh(idx), pr.mesh(idx), f"flat {idx}")
    assert faces > 50_000
    # positions-only call shape (collider path: null normal / uv SparsePtrs)
    for idx in [(0, 2, 0), (2, 0, 0), (-2, -2, -2)]:
        assert_same_mesh(po.mesh(idx, normal=False, uv=False), pr.mesh(idx, normal=False, uv=False), f"collider {idx}")


def test_deformed_meshes_are_bit_identical(R):
    cc, po, pr = make_default_pair(R, platforms=False)
    for idx in [(0, 2, 0), (2, 2, 2), (-2, 0, 1), (1, -2, -1), (0, 0, 0)]:
        for radius in (16.0, 1.0, 40.0):
            a = po.mesh(idx, deformed=True, deform_radius=radius)
            b = pr.mesh(idx, deformed=True, deform_radius=radius)
            assert_same_mesh(a, b, f"deformed {idx} r={radius}")


def test_transparent_double_sided_and_neighbour_rules(R):
    """Random chunks mixing every flag combination (glass 13: transparent + double-sided, forcefield 9: transparent), with
    neighbours that exist / have no content / are absent."""
    rng = np.random.default_rng(5)
    ids = (2, 3, 7, 9, 13)
    po, pr = O.Planet(), R.Planet()
    layout = {(0, 0, 0): 0.5, (1, 0, 0): 0.7, (-1, 0, 0): 0.2, (0, 1, 0): 0.9, (0, 0, 1): 0.5, (0, 0, -1): None}  # None: exists, no content; (0,-1,0) absent
    for idx, p in layout.items():
        blocks = None if p is None else random_chunk(rng, p, ids)
        assert po.add_chunk(idx, blocks) == 0 and pr.add_chunk(idx, blocks) == 0
    faces = 0
    for idx, p in layout.items():
        a, b = po.mesh(idx), pr.mesh(idx)
        if p is None:
            assert a is None and b is None
            continue
        faces += assert_same_mesh(a, b, f"mixed {idx}")
    assert faces > 100_000


def test_gravity_center_argument(R):
    rng = np.random.default_rng(6)
    blocks = random_chunk(rng, 0.3, (7, 13))
    po, pr = O.Planet(), R.Planet()
    po.add_chunk((0, 0, 0), blocks)
    pr.add_chunk((0, 0, 0), blocks)
    for g in [(0, 0, 0), (100, 3, -2), (-5, -200, 7), (4, 4, 300), (0.5, 0.5, 0.5)]:
        assert_same_mesh(po.mesh((0, 0, 0), gravity_center=g), pr.mesh((0, 0, 0), gravity_center=g), f"gravity {g}")


def test_greedy_box_collider_is_identical(R):
    cc, po, pr = make_default_pair(R)
    total = 0
    for idx in grid(cc):
        a, b = po.box_collider(idx), pr.box_collider(idx)
        assert a.shape == b.shape and np.array_equal(a, b), idx
        total += len(a)
    assert total > 100
    rng = np.random.default_rng(8)
    blocks = random_chunk(rng, 0.6, (7, 9, 13))  # forcefield has no collisions
    po.add_chunk((9, 9, 9), blocks)
    pr.add_chunk((9, 9, 9), blocks)
    assert np.array_equal(po.box_collider((9, 9, 9)), pr.box_collider((9, 9, 9)))


def test_edit_stream_and_dirty_sets(R):
    cc, po, pr = make_default_pair(R, platforms=False)
    po.clear_invalidated()
    pr.clear_invalidated()
    rng = np.random.default_rng(9)
    for tick in range(3):
        n = 400
        edits = np.zeros((n, 7), np.int32)
        edits[:, 0:3] = rng.integers(-3, 4, (n, 3))       # some chunks do not exist
        edits[:, 3:6] = rng.integers(0, 32, (n, 3))
        edits[: n // 4, 3:6] = rng.choice([0, 31], (n // 4, 3))  # boundary blocks raise neighbour masks
        edits[:, 6] = rng.choice([0, 2, 7, 13], n)
        po.apply_edits(edits)
        pr.apply_edits(edits)
        a, b = po.collect_dirty(), pr.collect_dirty()
        assert np.array_equal(a, b), f"tick {tick}"
        assert len(a) > 0
    for idx in grid(cc):
        assert np.array_equal(po.get_blocks(idx), pr.get_blocks(idx)), idx
    for idx in [(0, 0, 0), (2, 2, 2), (-1, 2, 0)]:
        assert_same_mesh(po.mesh(idx), pr.mesh(idx), f"after edits {idx}")


def test_ship_mask_rule_generate_and_hull_collider():
    """Ship.cpp compiled from the reference vs the restatement: Ship::Generate's blocks, the Up / Down-swapped neighbour mask of
    Ship::AddChunk (Ship.cpp:36-54) per edit, and Ship::BuildHullCollider for one and for several chunks (Ship.cpp:63-93)."""
    R = O.reference()
    rng = np.random.default_rng(5)
    for small in (True, False):
        a, b = O.Ship(2.0), R.Ship(2.0)
        a.generate(small)
        b.generate(small)
        assert np.array_equal(a.get_blocks((0, 0, 0)), b.get_blocks((0, 0, 0)))
        assert a.take_invalidated() == b.take_invalidated()
        for _ in range(200):
            loc = rng.choice([0, 1, 15, 30, 31], 3)
            blk = int(rng.choice([0, 4, 9, 13]))
            assert a.update_block((0, 0, 0), loc, blk) == 0 and b.update_block((0, 0, 0), loc, blk) == 0
            assert a.take_invalidated() == b.take_invalidated(), loc
        assert np.array_equal(a.get_blocks((0, 0, 0)), b.get_blocks((0, 0, 0)))
        ha, hb = a.hull_collider(), b.hull_collider()
        assert len(ha) > 0 and np.array_equal(ha, hb)
        # a second chunk turns the hull into a compound of per-chunk compounds with chunk offsets
        extra = (rng.random(32768) < 0.3).astype(np.uint16) * 4
        assert a.add_chunk((1, 0, -1), extra) == 0 and b.add_chunk((1, 0, -1), extra) == 0
        ha, hb = a.hull_collider(), b.hull_collider()
        key = lambda h: h[np.lexsort(h.T[::-1])]  # the reference iterates a hash map: compare as sets of rows
        assert len(ha) == len(hb) and np.array_equal(key(ha), key(hb))
        assert {tuple(r) for r in ha[:, 0:3].tolist()} == {(0.0, 0.0, 0.0), (64.0, 0.0, -64.0)}
    empty = R.Ship(1.0)
    assert len(empty.hull_collider()) == 0 and len(O.Ship(1.0).hull_collider()) == 0


def test_planet_sample_and_collider_bench_agree():
    """The CPU arm's workload (a planet-wide sample of generate + mesh) gives the same faces on the restatement and on the reference's
    own code, and equals meshing those chunks inside the whole planet."""
    R = O.reference()
    count = (5, 5, 5)
    a, b = O.PlanetSample(42, count, 7, 3, nthreads=2), R.PlanetSample(42, count, 7, 3, nthreads=2)
    ra, rb = a.step(2), b.step(2)
    assert ra["chunks"] == rb["chunks"] == len(range(3, 125, 7)) and ra["faces"] == rb["faces"] > 0
    whole = O.Planet()
    assert whole.generate(42, count, nthreads=2)
    grid = [(x - 2, y - 2, z - 2) for z in range(5) for y in range(5) for x in range(5)]
    want = sum((whole.mesh(grid[k]).face_count if whole.mesh(grid[k]) is not None else 0) for k in range(3, 125, 7))
    assert ra["faces"] == want
    rw = R.Planet()
    assert rw.generate(42, count, nthreads=2)
    ta, na, ca = O.bench_box_colliders(whole, count, 2)
    tb, nb, cb = R.bench_box_colliders(rw, count, 2)
    assert (na, ca) == (nb, cb) and ca == 125 and na > 0
