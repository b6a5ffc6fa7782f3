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
