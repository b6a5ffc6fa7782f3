"""The reference's only known-answer vectors — tests/UnitTests/Common/ChunkTests.cpp:17-43 — checked against the CPU oracle,
the reference's own ChunkContainer (oracle/_ref) and the host mirror (thisspaceofmine_b200.ChunkContainer).  They pin ChunkContainer::GetBlockIndices
(ChunkContainer.inl:14-17) and GetChunkIndicesByBlockIndices (ChunkContainer.inl:19-42)."""
import pytest

import oracle as O
from thisspaceofmine_b200.host import ChunkContainer

ChunkSize = 32
H = ChunkSize // 2

# ChunkTests.cpp:17-23
BLOCK_INDICES_KAT = [
    ((0, 0, 0), (-H, -H, -H)),
    ((-1, 0, 0), (-H * 3, -H, -H)),
    ((1, 0, 0), (H, -H, -H)),
    ((0, -1, 0), (-H, -H * 3, -H)),
    ((0, 1, 0), (-H, H, -H)),
    ((0, 0, -1), (-H, -H, -H * 3)),
    ((0, 0, 1), (-H, -H, H)),
]

# ChunkTests.cpp:37-43
ROUND_TRIP_KAT = [
    ((1, 0, 0), (1, 0, 0)),
    ((-1, 0, 0), (0, 0, 0)),
    ((-1, 0, 0), (0, 0, 1)),
    ((-1, 0, 0), (0, 1, 0)),
    ((-1, 42, 2), (14, 10, 12)),
    ((3, 42, -2), (17, 10, ChunkSize - 1)),
    ((3, 42, -2), (ChunkSize - 1, ChunkSize - 1, ChunkSize - 1)),
]

def _ref(name):
    def call(*a):
        return getattr(O.reference(), name)(*a)
    return call


IMPLS = [
    pytest.param((O.get_block_indices, O.get_chunk_indices_by_block_indices), id="oracle"),
    pytest.param((_ref("get_block_indices"), _ref("get_chunk_indices_by_block_indices")), id="reference-sources",
                 marks=pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref not built")),
    pytest.param((ChunkContainer.GetBlockIndices, ChunkContainer.GetChunkIndicesByBlockIndices), id="host-mirror"),
]


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("chunk,want", BLOCK_INDICES_KAT)
def test_block_indices(impl, chunk, want):
    assert tuple(impl[0](chunk, (0, 0, 0))) == want


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("chunk,local", ROUND_TRIP_KAT)
def test_converting_back_and_forth(impl, chunk, local):
    g = impl[0](chunk, local)
    c, l = impl[1](g)
    assert tuple(c) == chunk
    assert tuple(l) == local


@pytest.mark.parametrize("impl", IMPLS)
def test_round_trip_dense(impl):
    """Beyond the reference's 7 samples: every local coordinate of a few chunks, negative indices included."""
    for chunk in [(0, 0, 0), (-2, 1, -3), (5, -7, 2)]:
        for local in [(0, 0, 0), (31, 0, 0), (0, 31, 0), (0, 0, 31), (31, 31, 31), (15, 16, 17), (1, 30, 2)]:
            g = impl[0](chunk, local)
            assert g == (chunk[0] * 32 + local[0] - 16, chunk[1] * 32 + local[2] - 16, chunk[2] * 32 + local[1] - 16)  # y/z swap
            c, l = impl[1](g)
            assert (tuple(c), tuple(l)) == (chunk, local)


# ------------------------------------------------------------------------------------------------ Perlin noise core
# siv::PerlinNoise (third party, absent from /root/reference) implements Ken Perlin's "improved noise" (SIGGRAPH 2002); the reference
# implementation's permutation table and its value at (3.14, 42, 7) are published with it and repeated in the test suites of most ports.
# The restated noise core (oracle/tsom_oracle.hpp PerlinNoise::noise3D, which the device terrain_kernel mirrors operation for operation)
# must reproduce that value when it is handed that table.  This pins fade / lerp / grad / the lattice hashing; the seed -> permutation
# shuffle of siv::PerlinNoise::reseed stays unpinned.
KEN_PERLIN_PERMUTATION = [
    151, 160, 137, 91, 90, 15, 131, 13, 201, 95, 96, 53, 194, 233, 7, 225, 140, 36, 103, 30, 69, 142, 8, 99, 37, 240, 21, 10, 23, 190, 6, 148,
    247, 120, 234, 75, 0, 26, 197, 62, 94, 252, 219, 203, 117, 35, 11, 32, 57, 177, 33, 88, 237, 149, 56, 87, 174, 20, 125, 136, 171, 168, 68, 175,
    74, 165, 71, 134, 139, 48, 27, 166, 77, 146, 158, 231, 83, 111, 229, 122, 60, 211, 133, 230, 220, 105, 92, 41, 55, 46, 245, 40, 244, 102, 143, 54,
    65, 25, 63, 161, 1, 216, 80, 73, 209, 76, 132, 187, 208, 89, 18, 169, 200, 196, 135, 130, 116, 188, 159, 86, 164, 100, 109, 198, 173, 186, 3, 64,
    52, 217, 226, 250, 124, 123, 5, 202, 38, 147, 118, 126, 255, 82, 85, 212, 207, 206, 59, 227, 47, 16, 58, 17, 182, 189, 28, 42, 223, 183, 170, 213,
    119, 248, 152, 2, 44, 154, 163, 70, 221, 153, 101, 155, 167, 43, 172, 9, 129, 22, 39, 253, 19, 98, 108, 110, 79, 113, 224, 232, 178, 185, 112, 104,
    218, 246, 97, 228, 251, 34, 242, 193, 238, 210, 144, 12, 191, 179, 162, 241, 81, 51, 145, 235, 249, 14, 239, 107, 49, 192, 214, 31, 181, 199, 106, 157,
    184, 84, 204, 176, 115, 121, 50, 45, 127, 4, 150, 254, 138, 236, 205, 93, 222, 114, 67, 29, 24, 72, 243, 141, 128, 195, 78, 66, 215, 61, 156, 180,
]


def test_noise_core_reproduces_ken_perlins_reference_value():
    import oracle as O

    assert sorted(KEN_PERLIN_PERMUTATION) == list(range(256))
    assert O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 3.14, 42.0, 7.0) == 0.13691995878400012
    # gradient noise vanishes on the integer lattice, for any table
    for p in [(0.0, 0.0, 0.0), (5.0, -3.0, 17.0), (255.0, 256.0, 257.0)]:
        assert O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, *p) == 0.0
    # the table wraps at 256 in every axis
    a = O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 1.3, 2.7, 0.34567)
    assert O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 257.3, 2.7, 0.34567) == pytest.approx(a, abs=1e-12)  # 257.3 - 257 != 1.3 - 1 in binary
    assert O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 1.3, 258.7, 0.34567) == pytest.approx(a, abs=1e-12)
    assert O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 1.25, 2.5, 0.75) == O.perlin_noise3d_perm(KEN_PERLIN_PERMUTATION, 257.25, 258.5, 256.75)  # exact fractions
