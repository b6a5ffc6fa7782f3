"""Planet gen + remesh run used for timing and under ncu (profiles/): python tools/profile_planet.py [edge] [iters] [--deformed] [--collider] [--no-render]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thisspaceofmine_b200 as T

pos = [a for a in sys.argv[1:] if not a.startswith("--")]
flags = {a for a in sys.argv[1:] if a.startswith("--")}
edge = int(pos[0]) if len(pos) > 0 else 15
iters = int(pos[1]) if len(pos) > 1 else 3
ctx = T.Context(0)
cc = (edge, edge, edge)
planet = T.Planet(ctx, capacity=edge ** 3)
grid = T.Planet.chunk_grid(cc)
for it in range(iters):
    w0 = time.perf_counter()
    planet.GenerateChunks(ctx.blockLibrary, 42, cc)
    w1 = time.perf_counter()
    t = ctx.timings()
    b = planet.BuildMeshes(grid, render="--no-render" not in flags, collider="--collider" in flags, deformed="--deformed" in flags, deformRadius=16.0, views=(it == 0))
    w2 = time.perf_counter()
    t2 = ctx.timings()
    print(f"iter {it}: terrain {t['terrain_ms']:.3f} ms, generate {t['generate_ms']:.3f} ms, mesh {t2['mesh_emit_ms']:.3f} ms; wall clock of the calls: generate {(w1 - w0) * 1e3:.3f} ms, mesh {(w2 - w1) * 1e3:.3f} ms" + (" (count-only pass: arenas grew)" if t2["mesh_reemit"] else ""))
print("faces", ctx.arenas()[3], " ".join(sorted(flags)))
