#!/usr/bin/env python
"""Packets::ChunkReset payloads for a whole planet: python tools/bench_lz4.py [edge]
Device: tsom_chunks_lz4_compress (blocks stay in HBM, only the LZ4 blocks cross PCIe) and tsom_chunks_reset_lz4, wall clock around the
public calls.  CPU beside it: the real LZ4 library (pyarrow's lz4_raw codec = LZ4_compress_default) on the same chunks after a
tsom_chunks_get_content, one task per chunk on all host threads — what the reference's server does per joining player."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thisspaceofmine_b200 as T

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 15
ctx = T.Context(0)
cc = (edge,) * 3
n = edge ** 3
planet = T.Planet(ctx, capacity=n)
planet.GenerateChunks(ctx.blockLibrary, 42, cc)
grid = T.Planet.chunk_grid(cc)
planet.CompressChunks(grid[: min(n, 64)])  # warm-up (allocations)
t0 = time.perf_counter()
streams = planet.CompressChunks(grid)
t1 = time.perf_counter()
compress_kernels_ms = ctx.timings()["io_ms"]
client = T.Planet(ctx, capacity=n)
client.ResetChunksCompressed(grid[: min(n, 64)], streams[: min(n, 64)])
t2 = time.perf_counter()
client.ResetChunksCompressed(grid, streams)
t3 = time.perf_counter()
decode_kernel_ms = ctx.timings()["io_ms"]
out = {"chunks": n, "compressed_bytes": sum(map(len, streams)), "raw_bytes": n * 65536,
       "gpu_compress_ms": (t1 - t0) * 1e3, "gpu_compress_kernels_ms": compress_kernels_ms, "gpu_decode_kernel_ms": decode_kernel_ms, "gpu_compress_chunks_per_s": n / (t1 - t0),
       "gpu_reset_lz4_ms": (t3 - t2) * 1e3, "gpu_reset_lz4_chunks_per_s": n / (t3 - t2)}
try:
    import pyarrow as pa

    codec = pa.Codec("lz4_raw")
    sample = grid[:: max(1, n // 4096)]
    t4 = time.perf_counter()
    blocks = planet.GetContent(sample)
    raw = [blocks[i].tobytes() for i in range(len(sample))]
    threads = len(os.sched_getaffinity(0))
    with ThreadPoolExecutor(threads) as pool:
        ref = list(pool.map(lambda b: codec.compress(b, asbytes=True), raw))
    t5 = time.perf_counter()
    pos = {tuple(g): i for i, g in enumerate(map(tuple, grid))}
    assert all(streams[pos[tuple(s)]] == r for s, r in zip(sample, ref)), "device streams differ from the LZ4 library"
    out["cpu_lz4_library"] = {"chunks": len(sample), "threads": threads, "chunks_per_s": len(sample) / (t5 - t4), "note": "D2H of the raw blocks + LZ4_compress_default per chunk; byte-identical to the device streams"}
except ImportError:
    pass
print(json.dumps(out))
