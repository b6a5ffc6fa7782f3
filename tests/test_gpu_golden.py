"""GPU parity against the COMMITTED golden fixtures (tests/golden/*.npz) — no oracle involved at run time."""
import os
import zlib

import numpy as np
import pytest

import thisspaceofmine_b200 as T
from helpers import ATOL, RTOL, hash_fill

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLATFORMS = [(T.Direction.Right, (65, -18, -39)), (T.Direction.Back, (-34, 2, 53)), (T.Direction.Front, (22, -35, -59)), (T.Direction.Down, (23, -62, 26))]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


@pytest.mark.parametrize("name", ["planet_42_5x5x5", "planet_42_5x5x5_platforms", "planet_7_3x3x3"])
def test_planet_fixture(ctx, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    count = tuple(int(v) for v in g["count"])
    grid = g["grid"]
    assert np.array_equal(grid, T.Planet.chunk_grid(count))
    planet = T.Planet(ctx, capacity=len(grid))
    planet.GenerateChunks(ctx.blockLibrary, int(g["seed"]), count)
    if int(g["platforms"]):
        for d, c in PLATFORMS:
            planet.GeneratePlatform(ctx.blockLibrary, d, c)
    blocks = planet.GetContent(grid)
    assert np.array_equal(np.array([crc(b) for b in blocks], np.uint32), g["block_crc"])  # block ids bit-exact
    batch = planet.BuildMeshes(grid, render=True, collider=True)
    assert np.array_equal(batch.face_counts, g["face_count"])
    for i in range(len(grid)):
        m = batch.chunk(i)
        assert (0 if m is None else crc(m.indices)) == g["index_crc"][i]  # index buffers bit-exact
    k = 0
    pos = {tuple(idx): i for i, idx in enumerate(grid)}
    while f"keep{k}_idx" in g:
        idx = tuple(int(v) for v in g[f"keep{k}_idx"])
        m = batch.chunk(pos[idx])
        np.testing.assert_allclose(m.vertices, g[f"keep{k}_vertices"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(m.collider_positions, g[f"keep{k}_vertices"][:, 0:3], rtol=RTOL, atol=ATOL)
        assert np.array_equal(m.indices, g[f"keep{k}_indices"])
        assert np.array_equal(planet.BuildBoxCollider(idx), g[f"keep{k}_boxes"])
        k += 1
    planet.close()


def test_isolated_fixture(ctx):
    g = np.load(os.path.join(GOLDEN, "isolated_chunks.npz"))
    chunks = np.stack([hash_fill(1234, 0), hash_fill(1234, 1), hash_fill(1234, 2, p=0.05), g["mixed_blocks"]])
    c = T.ChunkContainer(ctx, 4)
    c.ResetChunks(g["idx"], chunks)
    batch = c.BuildMeshes(g["idx"])
    assert np.array_equal(batch.face_counts, g["face_count"])
    for i in range(4):
        assert crc(batch.chunk(i).indices) == g["index_crc"][i]
    np.testing.assert_allclose(batch.chunk(2).vertices, g["sparse_vertices"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(batch.chunk(3).vertices[:8000], g["mixed_vertices_head"], rtol=RTOL, atol=ATOL)
    d = c.BuildMeshes(g["idx"][2:3], deformed=True, deformRadius=16.0)
    np.testing.assert_allclose(d.chunk(0).vertices, g["sparse_deformed_vertices"], rtol=RTOL, atol=ATOL)
    c.close()
