#!/usr/bin/env python
"""BASELINE config 4 — edit-stream workload: N random block place/break ops per tick, remesh every dirty chunk and its masked
neighbours (ChunkEntities::Update), tick latency p50 / p99.

    python tools/bench_edits.py [--edge 5] [--edits 10000] [--ticks 300] [--cpu-ticks 3]

Per tick (wall clock around the public calls, host logic included — this is what a server tick pays):
    UpdateBlocks(edits)      tsom_apply_edits   : H2D of the edit list + apply_edits_kernel (+ neighbour masks)
    CollectInvalidated()     tsom_collect_dirty : dirty set = edited chunks + masked neighbours with content
    BuildMeshes(dirty)       tsom_mesh          : render vertices + collider positions + indices, one mesh_fused_kernel launch
Edits: chunk ~ U(all chunks), x, y, z ~ U[0, 32), id ~ U{0 empty, 2 dirt, 7 stone, 13 glass}; numpy MT19937 seed 2025.
Parity of the same stream (final blocks + meshes == oracle replay) is tests/test_gpu_gen_parity.py::test_edit_stream_and_dirty_set.
The CPU arm replays a few ticks on the reference's own code (oracle/_ref) with one task per dirty chunk on all host threads."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def make_edits(rng, grid, n):
    e = np.zeros((n, 7), np.int64)
    e[:, 0:3] = grid[rng.integers(0, len(grid), n)]
    e[:, 3:6] = rng.integers(0, 32, (n, 3))
    e[:, 6] = rng.choice([0, 2, 7, 13], n)
    return e


def run(T, ctx, args):
    """The measurement itself, on an existing context (bench.py calls it for its `edits` sub-object)."""
    cc = (args.edge,) * 3
    n = args.edge ** 3
    planet = T.Planet(ctx, capacity=n)
    planet.GenerateChunks(ctx.blockLibrary, args.seed, cc)
    grid = T.Planet.chunk_grid(cc)
    planet.BuildMeshes(grid, render=True, collider=True, views=False)  # sizes the arenas for the worst case (everything dirty)
    planet.CollectInvalidated(clear=True)
    entities = T.ChunkEntities(planet, render=True, collider=True)

    rng = np.random.Generator(np.random.MT19937(2025))
    streams = [make_edits(rng, grid, args.edits) for _ in range(args.warmup + args.ticks)]
    packed = [T.host.pack_edits(e) for e in streams]  # tsom_edit records, as a server would already hold them

    lat, dirty_counts, faces, k_edit, k_mesh = [], [], [], [], []
    for t, edits in enumerate(packed):
        ctx.synchronize()
        t0 = time.perf_counter()
        planet.UpdateBlocks(edits)
        t1 = time.perf_counter()
        batch = entities.Update()
        ctx.synchronize()
        t2 = time.perf_counter()
        if t == args.warmup - 1 and batch is not None:
            # a server sizes its arenas for the worst tick up front (tsom_reserve_faces); otherwise the arena doubles on demand and
            # the tick that triggers it pays a cudaMalloc (tens of ms) — random edits keep adding faces for hundreds of ticks
            ctx.reserve_faces(int(3 * batch.total_faces), render=True, collider=True)
        if t >= args.warmup:
            lat.append((t2 - t0) * 1e3)
            k_edit.append((t1 - t0) * 1e3)
            k_mesh.append(ctx.timings()["mesh_emit_ms"])
            dirty_counts.append(0 if batch is None else len(batch))
            faces.append(0 if batch is None else batch.total_faces)
    lat = np.array(lat)
    out = {
        "workload": f"config[3]: edit stream, {args.edits} random place/break ops per tick on the seed-{args.seed} {args.edge}^3 planet, remesh dirty chunks + masked neighbours (render + collider)",
        "ticks": args.ticks, "edits_per_tick": args.edits, "chunks": n,
        "tick_ms_p50": float(np.percentile(lat, 50)), "tick_ms_p99": float(np.percentile(lat, 99)), "tick_ms_mean": float(lat.mean()), "tick_ms_max": float(lat.max()),
        "apply_edits_ms_p50": float(np.percentile(k_edit, 50)), "mesh_kernel_ms_p50": float(np.percentile(k_mesh, 50)),
        "dirty_chunks_mean": float(np.mean(dirty_counts)), "faces_per_tick_mean": float(np.mean(faces)),
        "dirty_chunks_per_s": float(np.sum(dirty_counts) / (lat.sum() / 1e3)),
    }

    if args.cpu_ticks > 0:
        import oracle as O

        R, kind = (O.reference(), "reference") if O.reference_available() else (O, "port")
        threads = len(os.sched_getaffinity(0))
        ref = R.Planet()
        assert ref.generate(args.seed, cc, nthreads=threads)
        ref.clear_invalidated()
        from concurrent.futures import ThreadPoolExecutor

        pool = ThreadPoolExecutor(threads)
        cpu = []
        for edits in streams[: args.cpu_ticks]:
            t0 = time.perf_counter()
            ref.apply_edits(edits)
            dirty = ref.collect_dirty()
            # one task per dirty chunk (ctypes releases the GIL): Chunk::BuildMesh with normals + uv
            list(pool.map(lambda idx: ref.mesh(idx), [tuple(int(v) for v in d) for d in dirty]))
            cpu.append((time.perf_counter() - t0) * 1e3)
        out["cpu_baseline"] = {"tick_ms_mean": float(np.mean(cpu)), "ticks": args.cpu_ticks, "cores": threads, "kind": kind,
                               "note": "apply edits + dirty set + Chunk::BuildMesh (render attributes; no collider pass) per dirty chunk, results copied out through ctypes"}
    planet.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=5)
    ap.add_argument("--edits", type=int, default=10000)
    ap.add_argument("--ticks", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--cpu-ticks", type=int, default=3)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()

    import thisspaceofmine_b200 as T

    ctx = T.Context(0)
    print(json.dumps(run(T, ctx, args)))


if __name__ == "__main__":
    main()
