#!/usr/bin/env python
"""Development aid: generate_kernel cost by chunk class on the 47^3 planet (core = template copies, shell = everything else)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import thisspaceofmine_b200 as T  # noqa: E402

ctx = T.Context(0)
edge = int(sys.argv[1]) if len(sys.argv) > 1 else 47
cc = (edge,) * 3
grid = T.Planet.chunk_grid(cc)
planet = T.Planet(ctx, capacity=len(grid))
lib = ctx.blockLibrary
m = np.abs(grid).max(axis=1)
H = ((edge + 1) // 2) * 32
core_r = (H - H // 2 - 32 - 16) // 32  # chunks guaranteed inactive: blockDepth - maxH/2 >= 32
sets = {"all": grid, "core(|c|<=%d)" % core_r: grid[m <= core_r], "rest": grid[m > core_r], "outer layer": grid[m == edge // 2], "mid": grid[(m > core_r) & (m < edge // 2)]}
for name, idx in sets.items():
    idx = np.ascontiguousarray(idx)
    ts = []
    for it in range(4):
        planet.GenerateChunk(lib, idx, 42, cc)
        t = ctx.timings()
        ts.append((t["terrain_ms"], t["generate_ms"]))
    a = np.array(ts[1:]).mean(axis=0)
    print(f"{name}: {len(idx)} chunks terrain {a[0]:.3f} generate {a[1]:.3f} ms -> {a[1] * 1e6 / len(idx):.1f} ns/chunk, {len(idx) * 65536 / a[1] / 1e6:.0f} GB/s", flush=True)
planet.GenerateChunks(lib, 42, cc)
b = planet.BuildMeshes(grid, views=True)
fc = b.face_counts
print("chunks with faces", int((fc > 0).sum()), "of", len(fc), "faces", int(fc.sum()), "median faces of those", int(np.median(fc[fc > 0])), "p90", int(np.percentile(fc[fc > 0], 90)))
hist = np.histogram(fc[fc > 0], bins=[1, 256, 1025, 2048, 4096, 8192, 16384, 1 << 20])
print("hist", hist[0].tolist(), hist[1].tolist())
