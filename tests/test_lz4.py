"""Packets::ChunkReset wire payload (src/CommonLib/Protocol/Packets.cpp:172-212, src/CommonLib/Utility/BinaryCompressor.cpp:17-47):
one LZ4 block per chunk.

lz4 is a third-party dependency the reference does not vendor, but a real LZ4 library is on this image inside pyarrow (codec
"lz4_raw" = LZ4_compress_default / LZ4_decompress_safe on raw blocks), on the GPU box as well.  CPU tests pin the restatement
(oracle/lz4_block.hpp) to that library byte for byte; GPU tests compare the device kernels, through the C ABI, with both."""
import numpy as np
import pytest

import oracle as O

pa = pytest.importorskip("pyarrow")
N = 65536


def lib_compress(data: bytes) -> bytes:
    return pa.Codec("lz4_raw").compress(data, asbytes=True)


def lib_decompress(stream: bytes, size: int) -> bytes:
    return pa.Codec("lz4_raw").decompress(stream, decompressed_size=size, asbytes=True)


def chunk_shaped_inputs():
    rng = np.random.default_rng(20251019)
    yield "all air", np.zeros(32768, np.uint16)
    yield "all stone", np.full(32768, 7, np.uint16)
    yield "stone / stone_mossy bernoulli(0.9)", np.where(rng.random(32768) < 0.9, 7, 8).astype(np.uint16)
    yield "14 block types, uniform", rng.integers(0, 14, 32768).astype(np.uint16)
    yield "runs of 16", np.repeat(rng.integers(0, 14, 2048), 16).astype(np.uint16)
    yield "random 16-bit ids (incompressible)", rng.integers(0, 65536, 32768).astype(np.uint16)
    yield "half solid", (np.arange(32768) < 16384).astype(np.uint16) * 7
    a = np.zeros(32768, np.uint16)
    a[::1024] = 3  # sparse markers: long matches with a skip-strided search in between
    yield "sparse", a
    a = rng.integers(0, 2, 32768).astype(np.uint16)
    a[-8:] = [9, 10, 11, 12, 13, 1, 2, 3]  # literals in the tail region where no match may start
    yield "binary + tail", a
    planet = O.Planet(1.0, 16.0)
    assert planet.generate(42, (5, 5, 5), nthreads=8)
    for idx in [(0, 0, 0), (2, 0, 0), (1, 1, 1), (-2, -2, -2), (0, 2, 0), (1, -1, 2), (0, 0, 2)]:
        yield f"default planet chunk {idx}", planet.get_blocks(idx).astype(np.uint16)


def test_restated_compressor_equals_the_lz4_library():
    for name, ids in chunk_shaped_inputs():
        data = ids.tobytes()
        assert len(data) == N
        want = lib_compress(data)
        got = O.lz4_compress(data)
        assert got == want, f"{name}: {len(got)} vs {len(want)} bytes"
        assert O.lz4_decompress(want, N) == data, name


@pytest.mark.parametrize("n", [0, 1, 12, 13, 14, 100, 4097, 65535, 65537, 65546])
def test_restated_compressor_other_lengths(n):
    """Lengths around the format's limits (13 = shortest input that can hold a match, 65,546 = longest that still uses the 16-bit table)."""
    rng = np.random.default_rng(n)
    data = rng.integers(0, 4, n, dtype=np.uint8).tobytes()
    want = lib_compress(data)
    assert O.lz4_compress(data) == want
    assert O.lz4_decompress(want, n) == data


def _len_bytes(n):
    out = b""
    while n >= 255:
        out += b"\xff"
        n -= 255
    return out + bytes([n])


MALFORMED = {
    "match runs to the very end (no last literals)": bytes([0x1F]) + b"a" + b"\x01\x00" + _len_bytes(N - 1 - 4 - 15),
    "only 4 last literals": bytes([0x1F]) + b"a" + b"\x01\x00" + _len_bytes(N - 1 - 4 - 4 - 15) + bytes([0x40]) + b"bbbb",
    "literals into the last 12 bytes, then a match": bytes([0xF0]) + _len_bytes(N - 10 - 15) + b"a" * (N - 10) + b"\x01\x00" + bytes([0x60]) + b"bbbbbb",
    "offset before the start of the block": bytes([0x1F]) + b"a" + b"\x02\x00" + _len_bytes(N - 1 - 5 - 4 - 15) + bytes([0x50]) + b"bbbbb",
    "truncated length bytes": bytes([0xF0, 0xFF, 0xFF]),
    "literal run longer than the input": bytes([0x80]) + b"abc",
    "output overflow": bytes([0x1F]) + b"a" + b"\x01\x00" + _len_bytes(N) + bytes([0x50]) + b"bbbbb",
}
WELL_FORMED = {
    "match, then exactly 5 last literals": bytes([0x1F]) + b"a" + b"\x01\x00" + _len_bytes(N - 1 - 5 - 4 - 15) + bytes([0x50]) + b"bbbbb",
}


def test_restated_decoder_rejects_what_the_library_rejects():
    for name, s in MALFORMED.items():
        with pytest.raises(Exception):
            lib_decompress(s, N)
        with pytest.raises(ValueError):
            O.lz4_decompress(s, N)
    for name, s in WELL_FORMED.items():
        assert O.lz4_decompress(s, N) == lib_decompress(s, N), name


def test_restated_decoder_on_mutated_streams():
    """Bit flips in valid streams: whenever the restatement decodes all 65,536 bytes the library must accept the stream and produce the
    same bytes; whenever the library rejects, so must the restatement."""
    rng = np.random.default_rng(7)
    base = lib_compress(np.where(rng.random(32768) < 0.9, 7, 8).astype(np.uint16).tobytes())
    agree = 0
    for _ in range(300):
        s = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        s = bytes(s)
        try:
            want = lib_decompress(s, N)
        except Exception:
            want = None
        try:
            got = O.lz4_decompress(s, N)
        except ValueError:
            got = None
        if want is None:
            assert got is None
        if got is not None and len(got) == N:
            assert want == got
            agree += 1
    assert agree > 20


# ------------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_compress_is_byte_identical(ctx):
    import thisspaceofmine_b200 as T

    cases = list(chunk_shaped_inputs())
    n = len(cases)
    cont = T.ChunkContainer(ctx, n, 1.0)
    idx = np.zeros((n, 3), np.int32)
    idx[:, 0] = np.arange(n)
    cont.ResetChunks(idx, np.stack([ids for _, ids in cases]))
    streams = cont.CompressChunks(idx)
    for (name, ids), s in zip(cases, streams):
        want = lib_compress(ids.tobytes())
        assert s == want, f"{name}: {len(s)} vs {len(want)} bytes"
        assert s == O.lz4_compress(ids.tobytes())
    cont.close()


@pytest.mark.gpu
def test_gpu_planet_round_trip_through_packets(ctx):
    """Server side: generate on the device, compress every chunk; client side: a second container fed only the LZ4 blocks ends up with the
    same blocks and the same meshes."""
    import thisspaceofmine_b200 as T

    cc = (5, 5, 5)
    server = T.Planet(ctx, capacity=125)
    server.GenerateChunks(ctx.blockLibrary, 42, cc)
    grid = T.Planet.chunk_grid(cc)
    streams = server.CompressChunks(grid)
    blocks = server.GetContent(grid)
    for i in range(0, 125, 7):
        assert streams[i] == lib_compress(blocks[i].astype(np.uint16).tobytes())
        assert lib_decompress(streams[i], N) == blocks[i].astype(np.uint16).tobytes()
    client = T.Planet(ctx, capacity=125)
    client.ResetChunksCompressed(grid, streams)
    assert np.array_equal(client.GetContent(grid), blocks)
    a, b = server.BuildMeshes(grid), client.BuildMeshes(grid)
    assert np.array_equal(a.face_counts, b.face_counts) and a.total_faces == b.total_faces
    server.close()
    client.close()


@pytest.mark.gpu
def test_gpu_decoder_takes_library_streams_and_rejects_malformed_ones(ctx):
    import thisspaceofmine_b200 as T

    cases = list(chunk_shaped_inputs())
    n = len(cases)
    cont = T.ChunkContainer(ctx, n + 1, 1.0)
    idx = np.zeros((n, 3), np.int32)
    idx[:, 0] = np.arange(n)
    cont.ResetChunksCompressed(idx, [lib_compress(ids.tobytes()) for _, ids in cases])
    got = cont.GetContent(idx)
    for i, (name, ids) in enumerate(cases):
        assert np.array_equal(got[i], ids), name
    for name, s in WELL_FORMED.items():
        cont.ResetChunksCompressed(idx[:1], [s])
        assert cont.GetContent(idx[:1])[0].astype(np.uint16).tobytes() == lib_decompress(s, N), name
    before = cont.GetContent(idx)
    for name, s in MALFORMED.items():
        good = lib_compress(cases[1][1].tobytes())
        with pytest.raises(T.TsomError, match="failed to decompress chunk"):
            cont.ResetChunksCompressed(idx[:3], [good, s, good])
    short = lib_compress(bytes(1000))
    with pytest.raises(T.TsomError, match="malformed packet"):
        cont.ResetChunksCompressed(idx[:1], [short])
    assert np.array_equal(cont.GetContent(idx), before), "a rejected batch must not touch any chunk"
    cont.close()


@pytest.mark.gpu
def test_gpu_decoder_on_mutated_streams(ctx):
    """Bit flips, truncations and extensions of valid blocks: the device decoder accepts exactly the streams the restatement decodes to
    65,536 bytes (and yields the same bytes); everything else is rejected and leaves the chunk untouched."""
    import thisspaceofmine_b200 as T

    rng = np.random.default_rng(17)
    ids = np.where(rng.random(32768) < 0.9, 7, 8).astype(np.uint16)
    ids[rng.integers(0, 32768, 600)] = 0
    base = lib_compress(ids.tobytes())
    cont = T.ChunkContainer(ctx, 1, 1.0)
    idx = np.zeros((1, 3), np.int32)
    cont.ResetChunks(idx, ids[None, :])
    current = ids.copy()
    accepted = rejected = 0
    for trial in range(250):
        s = bytearray(base)
        kind = trial % 5
        if kind == 0:
            s = s[: int(rng.integers(1, len(s)))]                      # truncated
        elif kind == 1:
            s += bytes(rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8))  # trailing garbage
        else:
            for _ in range(int(rng.integers(1, 4))):
                s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
        s = bytes(s)
        try:
            want = O.lz4_decompress(s, N)
        except ValueError:
            want = None
        if want is not None and len(want) == N:
            cont.ResetChunksCompressed(idx, [s])
            current = np.frombuffer(want, np.uint16).copy()
            accepted += 1
        else:
            with pytest.raises(T.TsomError):
                cont.ResetChunksCompressed(idx, [s])
            rejected += 1
        assert np.array_equal(cont.GetContent(idx)[0], current), f"trial {trial}"
    assert accepted > 10 and rejected > 10
    cont.close()
