#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native TSOM chunk pipeline.

Metric (BASELINE.json): 32^3 chunks meshed / s.  Workload at every N: BASELINE config 2 — a synthetic batch of 4096 independent
32^3 chunks per GPU with random solid/air fill (p = 0.5 stone, counter-based hash, seed 1234), mesh-only, render attributes
(48-byte vertices + u32 indices).  "step" = one pass of Chunk::BuildMesh over the whole batch (one mesh_fused_kernel launch: classify, count, chain, emit).

  value      : chunks/s with the block ids already resident in HBM (device time, CUDA events on the library's stream)
  e2e        : the same through the C ABI with HOST buffers: H2D of the ids + meshing + D2H of vertices and indices, per step
  roofline   : the dominant kernel (mesh_fused_kernel): algorithmic bytes / its own CUDA-event duration vs the measured HBM peak;
               `traffic` = DRAM bytes of that launch from the committed ncu capture (profiles/r01/traffic.json), same chunk count
  cpu_baseline / --impl reference : the reference's own Chunk::BuildMesh (oracle/_ref: its sources compiled over the third-party
               stand-ins of oracle/ref_shim, kind "reference"; the oracle port, kind "port", only if that library is missing)
               on all host cores, on a bounded sample of the same workload

N > 1 (torchrun): chunks are independent, so each rank meshes its own 4096-chunk batch (weak scaling, no data-path collective);
time = max over ranks, value = N * 4096 / time.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

SEED = 1234
FILL_P = 0.5
STONE = 7
NBLK = 32768
CHUNKS_PER_GPU = 4096
FACE_BYTES = 216  # 4 x 48 B vertices + 6 x 4 B indices
CHUNK_READ_BYTES = 65536 + 8  # ids + (face count, arena offset); isolated chunks have no halo slabs


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(chunks):
    """dram__bytes_read.sum + dram__bytes_write.sum of one mesh_fused_kernel launch of this workload, from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01", "traffic.json")
    try:
        t = json.load(open(p))
        return t["dram_bytes"] if int(t["chunks"]) == int(chunks) else None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- workload (device side, torch plumbing)
def _s64(c):
    return c - (1 << 64) if c >= (1 << 63) else c


def device_fill(torch, device, seed, first_chunk, n_chunks, p=FILL_P, block=STONE):
    """Same function as tests/helpers.hash_fill (splitmix64 of (seed, chunk, idx)), evaluated on the GPU -> int16 [n, 32768]."""
    def lsr(z, k):
        return (z >> k) & ((1 << (64 - k)) - 1)

    i = torch.arange(NBLK, dtype=torch.int64, device=device)[None, :]
    ch = torch.arange(first_chunk, first_chunk + n_chunks, dtype=torch.int64, device=device)[:, None]
    z = (torch.full_like(ch, seed) * _s64(0x9E3779B97F4A7C15)) ^ (ch * _s64(0xD1B54A32D192ED03)) ^ i
    z = z + _s64(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * _s64(0x94D049BB133111EB)
    z = z ^ lsr(z, 31)
    thr = math.ceil(p * float(1 << 53))
    return torch.where(lsr(z, 11) < thr, block, 0).to(torch.int16).contiguous()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _run(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- CPU arm
try:
    ORIGINAL_AFFINITY = os.sched_getaffinity(0)
except Exception:
    ORIGINAL_AFFINITY = None


def host_threads():
    """All host threads this process may use (the affinity it was started with, not the per-GPU NUMA binding of the GPU arm)."""
    if ORIGINAL_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, ORIGINAL_AFFINITY)
        except Exception:
            pass
        return len(ORIGINAL_AFFINITY)
    return os.cpu_count() or 1


def cpu_sample(n_chunks, first_chunk=0):
    from helpers import hash_fill

    return np.stack([hash_fill(SEED, first_chunk + i, FILL_P, STONE) for i in range(n_chunks)])


def cpu_arm():
    """(module-like front-end, kind): the reference's own sources (oracle/_ref, "reference") when built, else the oracle port."""
    import oracle as O

    if O.reference_available():
        try:
            return O.reference(), "reference"
        except Exception:
            pass
    return O, "port"


CPU_KIND_NOTE = {
    "reference": "the reference's own Chunk::BuildMesh (src/CommonLib/Chunk.cpp + FlatChunk.cpp compiled from its sources over the oracle/ref_shim third-party stand-ins)",
    "port": "oracle port of Chunk::BuildMesh",
}


def cpu_mesh_rate(target_seconds):
    """The reference's Chunk::BuildMesh (render attributes) over a bounded sample of the benchmark workload, all host threads."""
    O, kind = cpu_arm()

    threads = host_threads()
    probe = cpu_sample(max(2 * threads, 8))
    t, faces = O.bench_mesh_isolated(probe, threads)
    rate = len(probe) / t
    n = int(max(len(probe), min(CHUNKS_PER_GPU, rate * target_seconds)))
    n = max(threads, (n // threads) * threads)
    blocks = cpu_sample(n)
    t, faces = O.bench_mesh_isolated(blocks, threads)
    t1, _ = O.bench_mesh_isolated(cpu_sample(8), 1)  # the same code on one thread, for the per-core picture
    return dict(value=n / t, seconds=t, chunks=n, faces=faces, threads=threads, kind=kind, value_1thread=8 / t1)


def run_reference(args):
    """--impl reference: the reference's CPU implementation (its own sources via oracle/_ref, else the oracle port) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    O, kind = cpu_arm()

    threads = host_threads()
    probe = cpu_sample(max(2 * threads, 8))
    t, _ = O.bench_mesh_isolated(probe, threads)
    rate = len(probe) / t
    budget = 120.0 / max(1, args.steps + args.warmup)  # whole run within ~2 minutes
    n = int(max(threads, min(CHUNKS_PER_GPU, rate * min(budget, 10.0))))
    n = max(threads, (n // threads) * threads)
    blocks = cpu_sample(n)
    for _ in range(args.warmup):
        O.bench_mesh_isolated(blocks, threads)
    times = []
    faces = 0
    for _ in range(args.steps):
        t, faces = O.bench_mesh_isolated(blocks, threads)
        times.append(t)
    ms = 1000.0 * sum(times) / len(times)
    value = n / (ms / 1000.0)
    sample = f"first {n} of the {CHUNKS_PER_GPU} hash-filled chunks (seed {SEED}, p={FILL_P}), {faces} faces, {CPU_KIND_NOTE[kind]} with normals+uv, one task per chunk"
    print(json.dumps({
        "impl": "reference", "metric": "chunks_meshed_per_sec", "value": value, "unit": "chunks/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16+f32", "data": "synthetic",
        "config": {"workload": "config[1]: 4096 independent 32^3 chunks, random solid/air fill p=0.5 (stone), mesh-only, render attributes", "chunks_per_step": n},
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated (first touch then
    places the e2e staging buffers on the GPU's NUMA node; without it the 45 GB/step D2H of 8 ranks crosses the socket link)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass  # best effort: the measurement is still valid, only possibly NUMA-remote


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    import thisspaceofmine_b200 as T
    from thisspaceofmine_b200 import _lib as L
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    bind_to_gpu_numa_node(local)

    n = args.chunks
    ctx = T.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    cont = T.ChunkContainer(ctx, n, 1.0)
    idx = np.zeros((n, 3), np.int32)
    idx[:, 0] = 2 * np.arange(n)  # every other chunk index along X: no face-adjacent chunk -> independent chunks, all boundary faces drawn

    # ---- inputs resident in HBM (per-rank distinct chunk ids so ranks do not mesh identical data)
    first_chunk = rank * n
    SUBGEN = 512
    for s in range(0, n, SUBGEN):
        m = min(SUBGEN, n - s)
        blk = device_fill(torch, device, SEED, first_chunk + s, m)
        torch.cuda.synchronize()
        cont.ResetChunks(idx[s:s + m], int(blk.data_ptr()))
        del blk
    # self-check of the device generator against the host one (first chunk of this rank)
    from helpers import hash_fill
    assert np.array_equal(cont.GetContent(idx[:1])[0], hash_fill(SEED, first_chunk)), "device fill != host fill"

    # ---- size the arenas once (count pass of an untimed first call tells the exact need)
    probe = cont.BuildMeshes(idx, views=True)
    total_faces = probe.total_faces
    algo_bytes = n * CHUNK_READ_BYTES + FACE_BYTES * total_faces

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        cont.BuildMeshes(idx, views=False)

    # warm-up: at least W (>= 3) steps, and at least ~1 s of them so a box that was idle has ramped its clocks (the first
    # process on a fresh box otherwise reads ~5 % low); the JSON reports the number of warm-up steps actually run
    warm_steps, t_warm = 0, time.perf_counter()
    while warm_steps < max(args.warmup, 3) or time.perf_counter() - t_warm < 1.0:
        one_step()
        warm_steps += 1

    # ---- timed region: exactly K steps, CUDA events on the library's stream, clocks sampled meanwhile
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    emit_ms, count_ms, scan_ms = [], [], []
    e0.record(stream)
    for _ in range(args.steps):
        one_step()
        t = ctx.timings()
        emit_ms.append(t["mesh_emit_ms"])
        count_ms.append(t["mesh_count_ms"])
        scan_ms.append(t["mesh_scan_ms"])
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * n / (ms_per_step / 1000.0)

    # ---- e2e: host buffers in, host buffers out, through the C ABI (sub-batches so the pinned output ring stays bounded)
    SUB = 256
    host_blocks = torch.empty((n, NBLK), dtype=torch.int16, pin_memory=True)
    host_blocks.copy_(torch.from_numpy(cont.GetContent(idx).view(np.int16)))
    max_sub_faces = int(max(probe.face_counts[s:s + SUB].sum() for s in range(0, n, SUB)))
    host_v = torch.empty(max_sub_faces * 192, dtype=torch.uint8, pin_memory=True)
    host_i = torch.empty(max_sub_faces * 24, dtype=torch.uint8, pin_memory=True)
    hb_ptr = host_blocks.data_ptr()

    def e2e_step():
        faces_out = 0
        for s in range(0, n, SUB):
            m = min(SUB, n - s)
            sub_idx = np.ascontiguousarray(idx[s:s + m])
            L.check(L.lib().tsom_chunks_reset(cont._h, sub_idx.ctypes.data_as(C.POINTER(C.c_int32)), m, C.cast(hb_ptr + s * NBLK * 2, C.POINTER(C.c_uint16))))
            cont.BuildMeshes(sub_idx, views=False)
            f = ctx.arenas()[3]
            L.check(L.lib().tsom_mesh_copy_to_host(ctx._h, 0, f, C.c_void_p(host_v.data_ptr()), C.cast(host_i.data_ptr(), C.POINTER(C.c_uint32)), None))
            faces_out += f
        return faces_out

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        faces_out = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert faces_out == total_faces
    if world > 1:
        tt = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * n / e2e_s

    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        emit = sum(emit_ms) / len(emit_ms)
        achieved = algo_bytes / (emit / 1000.0) / 1e9
        out = {
            "metric": "chunks_meshed_per_sec", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": warm_steps,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16+f32", "data": "synthetic",
            "config": {"workload": "config[1]: 4096 independent 32^3 chunks per GPU, random solid/air fill p=0.5 (stone), mesh-only, render attributes (48 B vertices + u32 indices)",
                       "chunks_per_gpu": n, "faces_per_step_per_gpu": total_faces, "l2": "inputs (268 MB) and outputs (45 GB) per step exceed the 126 MB L2; no explicit flush",
                       "parallelism": f"chunks sharded by index over {world} GPU(s), no collective"},
            "e2e": {"value": e2e_value, "unit": "chunks/s", "h2d_bytes_per_step": n * NBLK * 2, "d2h_bytes_per_step": total_faces * FACE_BYTES, "steps": e2e_steps,
                    "note": "tsom_chunks_reset (pinned host ids) + tsom_mesh + tsom_mesh_copy_to_host (pinned host vertices+indices), 256-chunk sub-batches"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "mesh_fused_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind, "frac_of_nominal_8000": achieved / 8000.0,
                         "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": emit, "traffic": ncu_traffic(n)},
            "kernels_ms": {"mesh_count": sum(count_ms) / len(count_ms), "mesh_scan": sum(scan_ms) / len(scan_ms), "mesh_emit": emit},
        }
        if not args.no_cpu and world == 1:
            c = cpu_mesh_rate(args.cpu_seconds)
            out["cpu_baseline"] = {"value": c["value"], "unit": "chunks/s", "cores": c["threads"], "kind": c["kind"], "value_1thread": c["value_1thread"],
                                   "sample": f"first {c['chunks']} of the 4096 chunks ({c['faces']} faces) in {c['seconds']:.1f} s, {CPU_KIND_NOTE[c['kind']]} with normals+uv, one task per chunk"}
        if args.planet and world == 1:
            try:
                out["variants"] = config1_variants(T, ctx, cont, idx, torch, device, peak)
            except Exception as e:  # secondary measurement: never lose the headline line
                out["variants"] = {"error": str(e)}
            try:
                out["planet"] = planet_bench(T, ctx, args.planet)
            except Exception as e:  # secondary measurement: never lose the headline line
                out["planet"] = {"error": str(e)}
            try:
                out["default_planet"] = default_planet_bench(T, ctx, not args.no_cpu)
            except Exception as e:
                out["default_planet"] = {"error": str(e)}
            try:
                # BASELINE config[3]: 10 k random place/break ops per tick on the 31^3 planet, remesh dirty chunks + neighbours
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_edits

                ea = argparse.Namespace(edge=args.planet, edits=10000, ticks=60, warmup=5, cpu_ticks=0 if args.no_cpu else 1, seed=42)
                out["edits"] = bench_edits.run(T, ctx, ea)
            except Exception as e:
                out["edits"] = {"error": str(e)}
    if args.planet and world > 1:
        # BASELINE configs 3 / 5: ONE planet sharded by chunk index over the ranks (strong scaling), halo slabs over NVLink P2P
        try:
            cont.close()
            sharded = {f"{e}^3": planet_sharded_bench(T, ctx, e, rank, world, dist, torch, device) for e in sorted({args.planet, 47})}
        except Exception as e:
            sharded = {"error": str(e)}
        if rank == 0:
            out["planet"] = sharded
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config1_variants(T, ctx, cont, idx, torch, device, peak):
    """SURVEY.md §8(d) config 2, the two other fills (kernel time only, 1024 of the container's chunks, 5 launches each):
    `sparse` p = 0.05 stone, where reading the ids dominates, and `mixed_ids` uniform over {0, 2, 3, 7, 8, 9, 13} = empty, dirt, grass,
    stone, stone_mossy, forcefield (transparent, no collision), glass (transparent, double sided), which exercises the same-type,
    transparent-neighbour and twin-face rules (parity of those rules: tests/test_gpu_mesh_parity.py)."""
    m = 1024
    sub = np.ascontiguousarray(idx[:m])
    table = torch.tensor([0, 2, 3, 7, 8, 9, 13], dtype=torch.int16, device=device)

    def mixed(first, count):
        h = torch.arange(first * NBLK, (first + count) * NBLK, dtype=torch.int64, device=device).view(count, NBLK)
        h = (h * _s64(0x9E3779B97F4A7C15)) ^ ((h >> 29) & 0x7FFFFFFFF)  # a multiplicative hash of the global block ordinal
        h = h * _s64(0xBF58476D1CE4E5B9)
        return table[((h >> 33) & 0x7FFFFFFF) % 7].contiguous()

    fills = {"sparse_p0.05": lambda first, count: device_fill(torch, device, SEED + 2, first, count, p=0.05), "mixed_ids": mixed}
    res = {}
    for name, fill in fills.items():
        for s0 in range(0, m, 256):
            blk = fill(s0, 256)
            torch.cuda.synchronize()
            cont.ResetChunks(sub[s0:s0 + 256], int(blk.data_ptr()))
            del blk
        faces = cont.BuildMeshes(sub, views=True).total_faces
        ms = []
        for _ in range(5):
            cont.BuildMeshes(sub, views=False)
            ms.append(ctx.timings()["mesh_emit_ms"])
        t = sum(ms) / len(ms)
        algo = m * CHUNK_READ_BYTES + FACE_BYTES * faces
        res[name] = {"chunks": m, "faces": int(faces), "kernel_ms": t, "chunks_per_s": m / (t / 1e3), "gbs": algo / (t / 1e3) / 1e9, "frac": algo / (t / 1e3) / 1e9 / peak}
    return res


def planet_bench(T, ctx, count):
    """Secondary numbers (BASELINE configs 3/5): full planet regenerate + remesh with cross-chunk halos on one GPU."""
    cc = (count, count, count)
    n = count ** 3
    planet = T.Planet(ctx, capacity=n)
    grid = T.Planet.chunk_grid(cc)
    lib = ctx.blockLibrary
    res = {"chunk_count": list(cc), "chunks": n}
    gen, cnt, emit, scan, wall = [], [], [], [], []
    for it in range(4):
        w0 = time.perf_counter()
        planet.GenerateChunks(lib, 42, cc)
        t = ctx.timings()
        batch = planet.BuildMeshes(grid, views=(it == 0))
        wall.append((time.perf_counter() - w0) * 1e3)  # both calls block until their kernels are done
        t2 = ctx.timings()
        if it == 0:
            faces = batch.total_faces
            nonempty = int((batch.face_counts > 0).sum())
            continue
        gen.append(t["terrain_ms"] + t["generate_ms"])
        cnt.append(t2["mesh_count_ms"])
        scan.append(t2["mesh_scan_ms"])
        emit.append(t2["mesh_emit_ms"])
    g, c, s, e = (sum(x) / len(x) for x in (gen, cnt, scan, emit))
    peak, _ = measured_peak_gbs()
    halo = 6 * 2048  # upper bound; boundary chunks have fewer neighbours
    mesh_bytes = n * (65536 + halo + 8) + 216 * faces
    w = sum(wall[1:]) / len(wall[1:])
    res.update(faces=faces, chunks_with_faces=nonempty, calls_wall_ms=w, gen_plus_mesh_chunks_per_s_wall=n / (w / 1e3), gen_ms=g, mesh_count_ms=c, mesh_scan_ms=s, mesh_emit_ms=e,
               gen_chunks_per_s=n / (g / 1e3), mesh_chunks_per_s=n / ((c + s + e) / 1e3), gen_plus_mesh_chunks_per_s=n / ((g + c + s + e) / 1e3),
               gen_gbs=n * 65536 / (g / 1e3) / 1e9, gen_frac=n * 65536 / (g / 1e3) / 1e9 / peak,
               mesh_gbs=mesh_bytes / ((c + s + e) / 1e3) / 1e9, mesh_frac=mesh_bytes / ((c + s + e) / 1e3) / 1e9 / peak)
    planet.close()
    return res


def default_planet_bench(T, ctx, with_cpu):
    """BASELINE config[0]: the default serverconfig spawn planet (seed 42, 5x5x5 chunks, src/Server/main.cpp:62,
    src/ServerLib/ServerPlanetEnvironment.cpp:49-59): generate + mesh every chunk, as server and client do at start-up.
    GPU: wall clock around GenerateChunks + BuildMeshes through the C ABI (125 chunks: launch-latency bound).
    CPU: the reference's own Planet::GenerateChunks + Chunk::BuildMesh for every chunk on all host threads (oracle/_ref)."""
    cc = (5, 5, 5)
    planet = T.Planet(ctx, capacity=125)
    grid = T.Planet.chunk_grid(cc)
    lib = ctx.blockLibrary
    planet.GenerateChunks(lib, 42, cc)
    faces = planet.BuildMeshes(grid, views=True).total_faces
    times = []
    for _ in range(20):
        ctx.synchronize()
        t0 = time.perf_counter()
        planet.GenerateChunks(lib, 42, cc)
        planet.BuildMeshes(grid, views=False)
        ctx.synchronize()
        times.append(time.perf_counter() - t0)
    planet.close()
    sec = statistics.median(times)
    res = {"chunk_count": list(cc), "chunks": 125, "faces": int(faces), "gpu_gen_plus_mesh_ms": sec * 1e3, "gpu_chunks_per_s": 125 / sec}
    if with_cpu:
        O, kind = cpu_arm()
        threads = host_threads()
        r = min((O.bench_planet(42, cc, threads) for _ in range(3)), key=lambda d: d["gen_s"] + d["mesh_s"])
        res["cpu_baseline"] = {"kind": kind, "cores": threads, "gen_ms": r["gen_s"] * 1e3, "mesh_ms": r["mesh_s"] * 1e3,
                               "chunks_per_s": 125 / (r["gen_s"] + r["mesh_s"]), "faces": r["faces"]}
    return res


def planet_sharded_bench(T, ctx, count, rank, world, dist, torch, device):
    """One count^3 planet over `world` GPUs: each rank generates its chunk range, pulls the facing slabs of remote neighbours
    over P2P (tsom_halo_pull, no NCCL on the data path) and meshes its range.  Time = max over ranks of the wall clock around
    generate + halo pull + mesh (device work synchronised on both sides); value = count^3 / time."""
    from thisspaceofmine_b200.sharding import ShardedPlanet

    cc = (count, count, count)
    sp = ShardedPlanet(ctx, cc, rank, world)
    lib = ctx.blockLibrary
    res = {"chunk_count": list(cc), "chunks": count ** 3, "n_gpus": world, "scaling": "strong", "halo_bytes_rank0": sp.plan.halo_bytes(0)}
    variants = [("flat_render", dict(render=True)), ("deformed_render_collider", dict(render=True, collider=True, deformed=True, deformRadius=16.0))]
    for name, kw in variants:
        times, gen_ms, mesh_ms, faces = [], [], [], 0
        for it in range(4):
            ctx.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            sp.GenerateChunks(lib, 42)
            t = ctx.timings()
            sp.PullHalos()
            batch = sp.BuildMeshes(views=(it == 0), **kw)
            ctx.synchronize()
            dt = time.perf_counter() - t0
            t2 = ctx.timings()
            if it == 0:
                faces = int(batch.total_faces)
                continue  # the first pass sizes the arenas
            tt = torch.tensor([dt, t["terrain_ms"] + t["generate_ms"], t2["mesh_emit_ms"]], device=device, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            times.append(float(tt[0]))
            gen_ms.append(float(tt[1]))
            mesh_ms.append(float(tt[2]))
        ft = torch.tensor([faces], device=device, dtype=torch.int64)
        dist.all_reduce(ft)
        sec = sum(times) / len(times)
        res[name] = {"faces": int(ft.item()), "step_ms": sec * 1e3, "gen_kernels_ms_max": sum(gen_ms) / len(gen_ms), "mesh_kernel_ms_max": sum(mesh_ms) / len(mesh_ms),
                     "gen_plus_mesh_chunks_per_s": count ** 3 / sec}
    sp.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--chunks", type=int, default=CHUNKS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--planet", type=int, default=31, help="edge (chunks) of the secondary full-planet gen+remesh measurement, 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
