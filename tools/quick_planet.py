#!/usr/bin/env python
"""Quick device timing of regenerate + remesh on one GPU (development aid): python tools/quick_planet.py 31 47"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import thisspaceofmine_b200 as T  # noqa: E402


def run(ctx, edge, kw, reps=5):
    cc = (edge,) * 3
    n = edge ** 3
    planet = T.Planet(ctx, capacity=n)
    grid = T.Planet.chunk_grid(cc)
    lib = ctx.blockLibrary
    out = []
    for it in range(reps + 1):
        w0 = time.perf_counter()
        planet.GenerateChunks(lib, 42, cc)
        t = ctx.timings()
        w1 = time.perf_counter()
        b = planet.BuildMeshes(grid, views=(it == 0), **kw)
        w2 = time.perf_counter()
        t2 = ctx.timings()
        if it == 0:
            faces = b.total_faces
            continue
        out.append((t["terrain_ms"], t["generate_ms"], t2["classify_ms"], t2["mesh_emit_ms"], (w1 - w0) * 1e3, (w2 - w1) * 1e3))
    a = np.array(out).mean(axis=0)
    planet.close()
    print(f"{edge}^3 {kw}: faces {faces}  terrain {a[0]:.3f}  generate {a[1]:.3f}  classify {a[2]:.3f}  mesh {a[3]:.3f} ms | wall gen {a[4]:.3f} mesh {a[5]:.3f} total {a[4] + a[5]:.3f} ms", flush=True)


if __name__ == "__main__":
    ctx = T.Context(0)
    for e in [int(v) for v in sys.argv[1:]] or [31]:
        run(ctx, e, dict(render=True))
        run(ctx, e, dict(render=True, collider=True, deformed=True, deformRadius=16.0))
    ctx.close()
