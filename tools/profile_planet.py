"""Small planet gen + remesh run used under ncu (profiles/): python tools/profile_planet.py [edge] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thisspaceofmine_b200 as T

edge = int(sys.argv[1]) if len(sys.argv) > 1 else 15
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = T.Context(0)
cc = (edge, edge, edge)
planet = T.Planet(ctx, capacity=edge ** 3)
grid = T.Planet.chunk_grid(cc)
for it in range(iters):
    planet.GenerateChunks(ctx.blockLibrary, 42, cc)
    t = ctx.timings()
    b = planet.BuildMeshes(grid, render=True, collider=False, views=(it == 0))
    t2 = ctx.timings()
    print(f"iter {it}: terrain {t['terrain_ms']:.3f} ms, generate {t['generate_ms']:.3f} ms, count {t2['mesh_count_ms']:.3f} ms, scan {t2['mesh_scan_ms']:.3f} ms, emit {t2['mesh_emit_ms']:.3f} ms")
print("faces", ctx.arenas()[3])
