"""Shared helpers of the parity tests: seeded inputs and GPU-vs-oracle comparison.

Parity bar (BASELINE.json north_star): block ids, face counts and index buffers bit-exact; positions, normals and UVs within
1e-5 relative.  RTOL is that tolerance; ATOL only covers values that are mathematically 0 (u = 6e-8 vs 0) and is an order of magnitude
tighter, so that an absolute error of 1e-5 on a 0.5 uv cannot hide behind it.
"""
import numpy as np

RTOL = 1e-5
ATOL = 1e-6

N = 32
NBLK = N ** 3


def lidx(x, y, z):
    return N * (N * z + y) + x


def random_chunk(rng, p=0.5, ids=(7,)):
    """Random solid/air fill: each block is non-empty with probability p, id drawn uniformly from ids."""
    solid = rng.random(NBLK) < p
    pick = rng.integers(0, len(ids), NBLK)
    return np.where(solid, np.asarray(ids, np.uint16)[pick], 0).astype(np.uint16)


def hash_fill(seed, chunk, p=0.5, block=7):
    """Counter-based fill (splitmix64 of (seed, chunk, idx)) used by the benchmark workload: identical on every host."""
    i = np.arange(NBLK, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) ^ (np.uint64(chunk) * np.uint64(0xD1B54A32D192ED03)) ^ i
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return np.where(u < p, block, 0).astype(np.uint16)


def assert_mesh_equal(gpu, ora, what="", check_attr=True):
    """gpu: thisspaceofmine_b200.ChunkMesh or None; ora: oracle.Mesh or None."""
    of = 0 if ora is None else ora.face_count
    gf = 0 if gpu is None else gpu.face_count
    assert gf == of, f"{what}: face count {gf} != oracle {of}"
    if of == 0:
        return dict(faces=0, inexact=0)
    assert np.array_equal(gpu.indices, ora.indices), f"{what}: index buffer differs"
    stats = dict(faces=of, inexact=0)
    if gpu.vertices is not None:
        np.testing.assert_allclose(gpu.position, ora.position, rtol=RTOL, atol=ATOL, err_msg=f"{what}: positions")
        stats["inexact"] += int((gpu.position != ora.position).sum())
        if check_attr:
            np.testing.assert_allclose(gpu.normal, ora.normal, rtol=RTOL, atol=ATOL, err_msg=f"{what}: normals")
            np.testing.assert_allclose(gpu.uvw, ora.uvw, rtol=RTOL, atol=ATOL, err_msg=f"{what}: uvw")
            stats["inexact"] += int((gpu.normal != ora.normal).sum()) + int((gpu.uvw != ora.uvw).sum())
            assert not gpu.tangent.any(), f"{what}: tangent slot must stay zero (Chunk::BuildMesh never writes it)"
    if gpu.collider_positions is not None:
        np.testing.assert_allclose(gpu.collider_positions, ora.position, rtol=RTOL, atol=ATOL, err_msg=f"{what}: collider positions")
    return stats
