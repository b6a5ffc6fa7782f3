"""Row f3 — the chunk save format, Chunk::Serialize / Chunk::Deserialize (src/CommonLib/Chunk.cpp:252-346).

CPU: the restated oracle writes byte-for-byte what the reference's own Chunk::Serialize writes (oracle/_ref, over the stand-in
ByteStream whose conventions — big-endian integers, UInt32-prefixed strings — are third-party behaviour restated in oracle/ref_shim),
both read each other's files back, and both reject malformed files with the reference's messages.
GPU: tsom_chunks_palettize + the host-side header == the oracle's bytes; tsom_chunks_reset_palettized restores the exact ids."""
import numpy as np
import pytest

import oracle as O
from helpers import NBLK, random_chunk

MANY = (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13)  # > 8 distinct types with Empty: 16-bit palette indices


def sample_chunks():
    rng = np.random.default_rng(31)
    return {
        (0, 0, 0): random_chunk(rng, 0.5, (7,)),                 # 2 types
        (1, 0, 0): random_chunk(rng, 0.7, (2, 3, 7, 8, 13)),     # 6 types
        (2, 0, 0): random_chunk(rng, 0.9, MANY),                 # 14 types -> u16 payload
        (3, 0, 0): np.zeros(NBLK, np.uint16),                    # 1 type (Empty only)
        (4, 0, 0): np.full(NBLK, 13, np.uint16),                 # 1 type, not Empty
        (5, 0, 0): random_chunk(rng, 1.0, (1, 2, 3, 4, 5, 6, 7, 8)),  # exactly 8 types, no Empty -> still u8
        (6, 0, 0): random_chunk(rng, 0.99, (1, 2, 3, 4, 5, 6, 7, 8)),  # 9 types -> u16
    }


def test_header_layout_is_what_chunk_cpp_writes():
    p = O.Planet()
    p.add_chunk((0, 0, 0), sample_chunks()[(1, 0, 0)])
    blob = p.serialize((0, 0, 0))
    assert blob[:4] == b"\x00\x00\x00\x01"                       # Constants::ChunkBinaryVersion
    assert blob[4:16] == b"\x00\x00\x00\x20" * 3                 # m_size
    assert blob[16:18] == b"\x00\x06"                            # 6 distinct types
    assert blob[18:22] == b"\x00\x00\x00\x05" and blob[22:27] == b"empty"
    names_len = sum(4 + len(n) for n in ("empty", "dirt", "grass", "stone", "stone_mossy", "glass"))
    assert len(blob) == 18 + names_len + NBLK                    # one byte per block


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref is not built and /root/reference is absent")
def test_oracle_bytes_equal_the_references_own_serialize():
    R = O.reference()
    po, pr = O.Planet(), R.Planet()
    chunks = sample_chunks()
    for k, v in chunks.items():
        po.add_chunk(k, v)
        pr.add_chunk(k, v)
    for k, v in chunks.items():
        a, b = po.serialize(k), pr.serialize(k)
        assert a == b, k
        # each side reads the other's file into a fresh chunk
        fresh = (k[0], 9, 9)
        po.add_chunk(fresh, None)
        pr.add_chunk(fresh, None)
        po.deserialize(fresh, b)
        pr.deserialize(fresh, a)
        assert np.array_equal(po.get_blocks(fresh), v) and np.array_equal(pr.get_blocks(fresh), v), k
    # generated planet chunks too
    po, pr = O.Planet(), R.Planet()
    assert po.generate(42, (3, 3, 3), 2) and pr.generate(42, (3, 3, 3), 2)
    for idx in [(-1, -1, -1), (0, 1, 0), (1, 1, 1)]:
        assert po.serialize(idx) == pr.serialize(idx)


def _malformed(blob):
    yield "incompatible chunk version", b"\x00\x00\x00\x02" + blob[4:]
    yield "incompatible chunk size", blob[:4] + b"\x00\x00\x00\x10" + blob[8:]
    yield "unknown block", blob[:22] + b"emptx" + blob[27:]


@pytest.mark.parametrize("backend", ["oracle", "reference"])
def test_malformed_files_are_rejected_like_the_reference(backend):
    if backend == "reference" and not O.reference_available():
        pytest.skip("oracle/_ref not built")
    B = O if backend == "oracle" else O.reference()
    p = B.Planet()
    p.add_chunk((0, 0, 0), sample_chunks()[(1, 0, 0)])
    blob = p.serialize((0, 0, 0))
    for msg, bad in _malformed(blob):
        with pytest.raises(RuntimeError, match=msg):
            p.deserialize((0, 0, 0), bad)
    with pytest.raises(RuntimeError):
        p.deserialize((0, 0, 0), blob[:-5])  # truncated payload


@pytest.mark.gpu
def test_gpu_serialize_matches_oracle_and_round_trips(ctx):
    import thisspaceofmine_b200 as T

    chunks = sample_chunks()
    keys = list(chunks)
    op = O.Planet()
    gc = T.ChunkContainer(ctx, 2 * len(keys), 1.0)
    for k, v in chunks.items():
        op.add_chunk(k, v)
        gc.AddChunk(k, v[None, :])
    lib = ctx.blockLibrary
    blobs = gc.SerializeChunks(keys, lib)
    for k, blob in zip(keys, blobs):
        assert blob == op.serialize(k), k
    # read them back into NEW chunks of the same container, ids must be identical; the chunks are invalidated like Chunk::Reset
    gc.CollectInvalidated(clear=True)
    fresh = [(k[0], 7, 7) for k in keys]
    gc.DeserializeChunks(fresh, blobs, lib)
    got = gc.GetContent(fresh)
    for i, k in enumerate(keys):
        assert np.array_equal(got[i], chunks[k]), k
    dirty = {tuple(d) for d in gc.CollectInvalidated(clear=True)}
    assert set(fresh) <= dirty
    # host-side header checks raise the reference's messages, device-side palette overflow is an error
    for msg, bad in _malformed(blobs[1]):
        with pytest.raises(RuntimeError, match=msg):
            gc.DeserializeChunks([fresh[1]], [bad], lib)
    bad = bytearray(blobs[0])
    bad[-1] = 200  # palette index 200 of 2
    with pytest.raises(T.TsomError):
        gc.DeserializeChunks([fresh[0]], [bytes(bad)], lib)
    gc.close()


@pytest.mark.gpu
def test_gpu_serialize_generated_planet(ctx):
    import thisspaceofmine_b200 as T

    cc = (3, 3, 3)
    planet = T.Planet(ctx, capacity=27)
    planet.GenerateChunks(ctx.blockLibrary, 42, cc)
    op = O.Planet()
    assert op.generate(42, cc, 2)
    grid = T.Planet.chunk_grid(cc)
    blobs = planet.SerializeChunks(grid, ctx.blockLibrary)
    for idx, blob in zip(grid, blobs):
        assert blob == op.serialize(tuple(int(v) for v in idx)), tuple(idx)
    planet.close()
