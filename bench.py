#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native TSOM chunk pipeline.

Metric (BASELINE.json): 32^3 chunks meshed / s (gen + mesh) at 1 / 2 / 4 / 8 B200, % of the HBM roofline.

Workload at EVERY N (BASELINE configs[2] / [4], the planet north_star asks to scale): ONE synthetic planet of 47 x 47 x 47 = 103,823
chunks (seed 42, block size 1, default BlockLibrary), regenerated and remeshed in full — Planet::GenerateChunks, the halo exchange,
Chunk::BuildMesh with render attributes (48-byte vertices + u32 indices) for every chunk — sharded by chunk index over the N GPUs
(one box of chunks per GPU, tsom_shard_plan; facing slabs of neighbour chunks on other GPUs pulled over NVLink P2P, ordered on the
device).  The work is fixed as N grows: `scaling` is "strong".  One process per GPU (torchrun), no collective on the data path.

  value        : chunks/s of a whole step (generate + exchange + mesh), block ids resident in HBM, CUDA events on the library's stream
                 around the step, barrier + synchronize on both sides, max over ranks
  e2e          : the same step through the C ABI with HOST results: + every chunk's vertices and indices copied to pinned host memory
  roofline     : the dominant kernel, mesh_fused_kernel, on the rank that ran longest: algorithmic bytes / its CUDA-event
                 duration vs the measured HBM peak; `traffic` = DRAM bytes of that launch from the committed ncu capture (N = 1)
  cpu_baseline / --impl reference : the reference's own Planet::GenerateChunk + Chunk::BuildMesh (oracle/_ref: its sources compiled
                 over the third-party stand-ins of oracle/ref_shim; the oracle port only if that library is missing) on all host cores,
                 on a bounded planet-wide sample of the same planet (every 53rd chunk, real neighbours)
  sharded_parity : mesh(N GPUs) == mesh(1 GPU) chunk by chunk (face counts + device-side checksums of vertex and index bytes) and the
                 sharded 5^3 planet == tests/golden (block ids, face counts, index buffers); at N = 1 the single-process multi-part
                 planet (tsom_planet_create_sharded, 3 parts on the one GPU) plays the sharded side
  microbench   : (N = 1) BASELINE config[1], the kernel roofline microbench of round 1: 4096 independent random-fill chunks, mesh-only

Secondary sub-objects at N = 1: 31^3 planet (flat / deformed + collider), default planet incl. its 4 platforms (config[0]), edit
stream (config[3]), batched box colliders.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

PLANET_EDGE = 47
PLANET_SEED = 42
SAMPLE_STRIDE = 53  # CPU arm: every 53rd chunk of the planet (1,959 chunks), coprime to 47 so the sample covers the planet evenly
SEED = 1234
FILL_P = 0.5
STONE = 7
NBLK = 32768
MICRO_CHUNKS = 4096
CHUNKS_PER_GPU = MICRO_CHUNKS  # config[1] (tests/test_gpu_fullsize.py checks the microbench workload at full size)
FACE_BYTES = 216  # 4 x 48 B vertices + 6 x 4 B indices
METRIC = "chunks_generated_and_meshed_per_sec"


def workload_config(edge, n_gpus):
    """`config` of the JSON line: the same object on both arms (the reference arm times a bounded sample of THIS workload)."""
    return {"workload": workload_name(edge), "chunk_count": [edge, edge, edge], "chunks": edge ** 3, "seed": PLANET_SEED,
            "l2": "per step the GPUs write 6.8 GB of ids and 37.9 GB of vertices + indices between them: far beyond the 126 MB L2, no explicit flush",
            "parallelism": f"chunks sharded by chunk index over {n_gpus} GPU(s): one box per GPU, P2P halo slabs, no collective on the data path"}


def workload_name(edge):
    return (f"one {edge}^3-chunk planet ({edge ** 3} chunks, seed {PLANET_SEED}): Planet::GenerateChunks + halo exchange + Chunk::BuildMesh (render attributes, "
            f"48 B vertices + u32 indices) of every chunk, sharded by chunk index over the GPUs (BASELINE configs[2]/[4])")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one mesh_fused_kernel launch of the named workload, from the committed ncu capture."""
    for rnd in ("r02", "r01"):
        p = os.path.join(ROOT, "profiles", rnd, "traffic.json")
        try:
            t = json.load(open(p))
            if key in t:
                return t[key]["dram_bytes"]
            if key == "config1_4096" and int(t.get("chunks", 0)) == 4096:
                return t["dram_bytes"]
        except Exception:
            continue
    return None


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- workload helpers
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


def neighbour_counts(grid, edge):
    """Existing face-adjacent neighbours of every chunk of a cubic planet (each one contributes a 2,048-byte halo slab)."""
    lo, hi = -(edge // 2), edge - edge // 2 - 1
    return ((grid > lo).sum(axis=1) + (grid < hi).sum(axis=1)).astype(np.int64)


def mesh_algo_bytes(n_chunks, nbr_total, faces, with_collider=False):
    """65,536 B ids + 2,048 B per neighbour slab + 8 B (count, offset) per chunk, 216 B per face (+ 48 B collider positions)."""
    return n_chunks * (65536 + 8) + 2048 * nbr_total + (FACE_BYTES + (48 if with_collider else 0)) * faces


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons of one GPU, polled through NVML every few milliseconds on a thread (the timed region of a multi-GPU
    run lasts tens of milliseconds: `nvidia-smi -lms` cannot sample inside it).  start() before the warm-up, mark_begin() / mark_end()
    around the timed region; the summary is over the samples taken between the two marks."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, period_s=0.004):
        self.rows, self.t0, self.t1, self.err = [], None, None, None
        self.period = period_s
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # no NVML: say so in the line instead of failing the run
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), mhz, reasons))
            except Exception as e:
                self.err = str(e)
                return
            self._stop.wait(self.period)

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self._stop.set()
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.err}"]}
        self.t.join(timeout=1.0)
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        window = "timed region"
        if len(inside) < 3:  # a very short region: the load is the same since the ramp began, so say what was sampled instead
            inside = [r for r in self.rows if self.t0 is None or r[0] >= self.t0 - 0.5]
            window = "timed region + the 0.5 s of identical steps before it"
        mask = 0
        for r in inside:
            mask |= r[2]
        return {"sm_mhz": statistics.median(r[1] for r in inside) if inside else None, "sm_min_mhz": min((r[1] for r in inside), default=None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if mask & b), "samples": len(inside), "window": window, "source": "NVML, 4 ms period"}


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
    "reference": "the reference's own Planet::GenerateChunk + Chunk::BuildMesh (src/CommonLib/Planet.cpp, Chunk.cpp, FlatChunk.cpp compiled from its sources over the oracle/ref_shim stand-ins)",
    "port": "oracle port of Planet::GenerateChunk + Chunk::BuildMesh",
}


def cpu_planet_sample(edge, steps, warmup):
    """The reference's generate + mesh over the bounded sample of the benchmark planet, all host threads -> (list of step dicts, kind, threads)."""
    O, kind = cpu_arm()
    threads = host_threads()
    sample = O.PlanetSample(PLANET_SEED, (edge,) * 3, SAMPLE_STRIDE, 0, nthreads=threads)
    for _ in range(warmup):
        sample.step(threads)
    return [sample.step(threads) for _ in range(steps)], kind, threads


def sample_note(kind, threads, r):
    return (f"every {SAMPLE_STRIDE}rd chunk of the planet in GenerateChunks order ({r['chunks']} of {PLANET_EDGE ** 3} chunks, {r['faces']} faces; their neighbours generated once at "
            f"set-up so the sample meshes with real halos), regenerated + meshed per step: {CPU_KIND_NOTE[kind]} with normals + uv, one task per chunk on {threads} threads")


def run_reference(args):
    """--impl reference: the reference's CPU implementation (its own sources via oracle/_ref, else the oracle port) on the host cores; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, kind, threads = cpu_planet_sample(args.edge, args.steps, args.warmup)
    sec = sum(s["gen_s"] + s["mesh_s"] for s in steps) / len(steps)
    r = steps[-1]
    value = r["chunks"] / sec
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16+f32", "data": "synthetic",
        "config": workload_config(args.edge, args.gpus), "sample_chunks_per_step": r["chunks"],
        "cpu_baseline": {"value": value, "unit": "chunks/s", "cores": threads, "kind": kind, "sample": sample_note(kind, threads, r),
                         "gen_ms": 1e3 * sum(s["gen_s"] for s in steps) / len(steps), "mesh_ms": 1e3 * sum(s["mesh_s"] for s in steps) / len(steps)},
        "e2e": {"value": value, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, before any pinned host buffer is allocated (first touch then
    places the e2e staging buffers on the GPU's NUMA node)."""
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
    from thisspaceofmine_b200.sharding import ShardedPlanet

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON of rank 0: everything else that writes to fd 1 while the run lasts (NCCL prints its version
    # banner there when the communicator is created) goes to stderr; the line itself is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    bind_to_gpu_numa_node(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return int(v)
        t = torch.tensor([int(v)], device=device, dtype=torch.int64)
        dist.all_reduce(t)
        return int(t.item())

    def gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    edge = args.edge
    cc = (edge,) * 3
    n_total = edge ** 3
    ctx = T.Context(local)
    lib = ctx.blockLibrary
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    sp = ShardedPlanet(ctx, cc, rank, world)
    n_local = len(sp.local_grid)
    nbr_local = int(neighbour_counts(sp.local_grid, edge).sum())

    def step(views=False):
        # three calls, one host wait: generate and the exchange are only queued, the mesh call returns when all three are done
        sp.GenerateChunks(lib, PLANET_SEED, wait=False)  # terrain + template + gen_classify + generate kernels on this rank's box
        sp.PullHalos()                                   # publish -> wait for the peers on the device -> halo_pull_kernel -> publish
        batch = sp.BuildMeshes(views=views)              # job_classify_kernel + mesh_fused_kernel
        tm = ctx.timings()
        return batch, tm, tm

    # ---- first pass sizes the arenas (the kernel counts, the host grows the arenas, the kernel runs again) and gives the per-chunk faces
    first, _, _ = step(views=True)
    faces_local = int(first.total_faces)
    face_counts = np.array(first.face_counts, np.uint32)
    faces_total = allsum(faces_local)

    # ---- ramp the clocks (a fresh box reads ~5 % low for its first second), then exactly W warm-up steps, then exactly K timed steps
    # (the ranks agree on when the ramp ends: a step is collective, every rank must run the same number of them)
    ramp_steps, t_ramp = 0, time.perf_counter()
    while allmax(1.0 if time.perf_counter() - t_ramp < args.ramp_seconds else 0.0) > 0.0:
        step()
        ramp_steps += 1
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step()
    launches0 = ctx.launch_count
    barrier()
    if sampler:
        sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt = []
    e0.record(stream)
    for _ in range(args.steps):
        _, tg, tm = step()
        kt.append((tg["terrain_ms"], tg["generate_ms"], tm["halo_ms"], tm["classify_ms"], tm["mesh_emit_ms"]))
    e1.record(stream)
    barrier()
    if sampler:
        sampler.mark_end()
    elapsed_ms = allmax(e0.elapsed_time(e1))
    launches = allsum(ctx.launch_count - launches0)
    clocks = sampler.stop() if sampler else None
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step / 1000.0)
    k_terrain, k_gen, k_halo, k_classify, k_mesh = (float(v) for v in np.mean(np.array(kt), axis=0))
    per_rank = gather(dict(rank=rank, chunks=n_local, faces=faces_local, nbr=nbr_local, mesh_ms=k_mesh, gen_ms=k_terrain + k_gen, halo_ms=k_halo,
                           halo_bytes=sp.plan.halo_bytes(rank), box=sp.plan.boxes[rank].tolist()))

    # ---- e2e: the same step with HOST results: every chunk's vertices + indices land in pinned host memory (a bounded ring, reused slice by slice)
    SLICE_FACES = 4 << 20  # 4 Mi faces = 864 MiB of vertices + 96 MiB of indices per slice
    host_v = torch.empty(SLICE_FACES * 192, dtype=torch.uint8, pin_memory=True)
    host_i = torch.empty(SLICE_FACES * 24, dtype=torch.uint8, pin_memory=True)

    def e2e_step(with_indices=True):
        batch, _, _ = step(views=True)
        total = int(batch.total_faces)
        for f0 in range(0, total, SLICE_FACES):
            f = min(SLICE_FACES, total - f0)
            L.check(L.lib().tsom_mesh_copy_to_host(ctx._h, f0, f, C.c_void_p(host_v.data_ptr()), C.cast(host_i.data_ptr(), C.POINTER(C.c_uint32)) if with_indices else None, None))
        return total

    def time_e2e(with_indices):
        e2e_step(with_indices)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            got = e2e_step(with_indices)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / args.e2e_steps
        assert got == faces_local
        return allmax(sec)

    e2e_s = time_e2e(True)
    e2e_noidx_s = time_e2e(False)
    del host_v, host_i

    # ---- parity of the sharded result
    parity = sharded_parity(T, L, ctx, sp, first, rank, world, gather, torch, dist)

    # ---- BASELINE config[3] on the sharded planet: 10 k random place/break ops per tick, every rank applies the edits of its own chunks,
    # one exchange round, every rank rebuilds its share of the dirty set (render + collider); tick latency = max over the ranks
    edits_sharded = None
    if world > 1:
        try:
            edits_sharded = sharded_edit_ticks(sp, ctx, world, barrier, allmax, allsum)
        except Exception as ex:  # secondary measurement: never lose the headline line
            edits_sharded = {"error": str(ex)}

    out = None
    if rank == 0:
        peak, peak_kind = measured_peak_gbs()
        slow = max(per_rank, key=lambda r: r["mesh_ms"])
        algo = mesh_algo_bytes(slow["chunks"], slow["nbr"], slow["faces"])
        achieved = algo / (slow["mesh_ms"] / 1e3) / 1e9
        mesh_all = [r["mesh_ms"] for r in per_rank]
        step_bytes = n_total * 65536 + sum(mesh_algo_bytes(r["chunks"], r["nbr"], r["faces"]) for r in per_rank)
        out = {
            "metric": METRIC, "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ramp_steps": ramp_steps,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16+f32", "data": "synthetic",
            "config": workload_config(edge, world), "faces_per_step": faces_total,
            "e2e": {"value": n_total / e2e_s, "unit": "chunks/s", "h2d_bytes_per_step": n_total * 24 + 1536 * world, "d2h_bytes_per_step": faces_total * FACE_BYTES + n_total * 16,
                    "steps": args.e2e_steps, "ms_per_step": e2e_s * 1e3,
                    "note": "tsom_planet_generate_chunks + tsom_halo_pull + tsom_mesh + tsom_mesh_copy_to_host of every face (pinned host vertices + indices, 4 Mi-face slices)"
                            + ("; all ranks copy at once, so the box's host memory bandwidth is shared" if world > 1 else ""),
                    "indices_on_host": {"value": n_total / e2e_noidx_s, "ms_per_step": e2e_noidx_s * 1e3, "d2h_bytes_per_step": faces_total * 192 + n_total * 16,
                                        "note": "indices are a pure function of the face ordinal (Chunk.cpp:95-101): left on the device, tsom_mesh_indices_for_faces writes them on the host"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "mesh_fused_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_kind": peak_kind,
                         "frac_of_nominal_8000": achieved / 8000.0, "algorithmic_bytes_per_launch": algo, "kernel_ms": slow["mesh_ms"], "rank": slow["rank"],
                         "traffic": ncu_traffic(f"planet{edge}") if world == 1 else None,
                         "note": "slowest rank; chunks whose summaries prove they have no face skip their 64 KiB stage, so DRAM traffic is below the algorithmic bytes"},
            "step": {"algorithmic_bytes": step_bytes, "gbs_by_wall": step_bytes / (ms_per_step / 1e3) / 1e9, "frac_by_wall": step_bytes / (ms_per_step / 1e3) / 1e9 / (peak * world),
                     "kernels_ms_rank0": {"terrain": k_terrain, "generate": k_gen, "halo": k_halo, "classify": k_classify, "mesh": k_mesh},
                     "mesh_kernel_ms_max": max(mesh_all), "mesh_kernel_ms_mean": sum(mesh_all) / len(mesh_all), "mesh_max_over_mean": max(mesh_all) / (sum(mesh_all) / len(mesh_all)),
                     "gen_kernels_ms_max": max(r["gen_ms"] for r in per_rank), "per_rank": per_rank},
            "sharded_parity": parity["status"], "sharded_parity_detail": parity,
        }
        if edits_sharded is not None:
            out["edits_sharded"] = edits_sharded
    if world == 1:
        extras(args, T, L, ctx, sp, out, torch, device)
    if rank == 0:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    sp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def sharded_edit_ticks(sp, ctx, world, barrier, allmax, allsum, edits_per_tick=10000, ticks=30, warmup=5):
    """ChunkEntities::Update on the sharded planet (Planet.cpp:39-58 across GPUs): the tick's edit list is the same on every rank."""
    sp.BuildMeshes(render=True, collider=True, views=False)  # collider arena for the worst case
    sp.CollectInvalidated(clear=True)
    rng = np.random.Generator(np.random.MT19937(2025))
    grid = sp.plan.grid
    lat, dirty_total, faces_total = [], [], []
    for t in range(warmup + ticks):
        e = np.zeros((edits_per_tick, 7), np.int64)
        e[:, 0:3] = grid[rng.integers(0, len(grid), edits_per_tick)]
        e[:, 3:6] = rng.integers(0, 32, (edits_per_tick, 3))
        e[:, 6] = rng.choice([0, 2, 7, 13], edits_per_tick)
        barrier()
        t0 = time.perf_counter()
        sp.UpdateBlocks(e)
        dirty, batch = sp.Update(render=True, collider=True, views=True)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        faces = 0 if batch is None else int(batch.total_faces)
        if t == warmup - 1:
            ctx.reserve_faces(int(3 * max(faces, 1)), render=True, collider=True)  # a server sizes its arenas for the worst tick up front
        dt = allmax(dt)
        nd, nf = allsum(len(dirty)), allsum(faces)
        if t >= warmup:
            lat.append(dt)
            dirty_total.append(nd)
            faces_total.append(nf)
    lat = np.array(lat)
    return {"workload": f"config[3] on the sharded {sp.plan.chunkCount[0]}^3 planet: {edits_per_tick} random place/break ops per tick, owners apply, halo exchange, every rank rebuilds its dirty chunks (render + collider)",
            "n_gpus": world, "ticks": ticks, "tick_ms_p50": float(np.percentile(lat, 50)), "tick_ms_p99": float(np.percentile(lat, 99)), "tick_ms_max": float(lat.max()),
            "dirty_chunks_mean": float(np.mean(dirty_total)), "faces_per_tick_mean": float(np.mean(faces_total)), "dirty_chunks_per_s": float(np.sum(dirty_total) / (lat.sum() / 1e3))}


def sharded_parity(T, L, ctx, sp, first, rank, world, gather, torch, dist):
    """mesh(N GPUs) == mesh(1 GPU), chunk by chunk, and the sharded 5^3 planet == the golden fixture."""
    res = {"status": "ok"}
    try:
        lib = ctx.blockLibrary
        edge = sp.plan.chunkCount[0]
        cc = sp.plan.chunkCount
        if world > 1:
            # every rank: per-chunk face counts + device-side checksums of what it just meshed; rank 0: the whole planet alone on its GPU
            sp.GenerateChunks(lib, PLANET_SEED)
            sp.PullHalos()
            batch = sp.BuildMeshes(views=True)
            mine = (sp.plan.local_ids(rank), np.array(batch.face_counts), batch.checksums())
            parts = gather(mine)
            if rank == 0:
                whole = T.Planet(ctx, capacity=len(sp.plan.grid))
                times = []
                for it in range(4):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    whole.GenerateChunks(lib, PLANET_SEED, cc)
                    ref = whole.BuildMeshes(sp.plan.grid, views=(it == 0))
                    torch.cuda.synchronize()
                    times.append((time.perf_counter() - t0) * 1e3)
                    if it == 0:
                        want_faces, want_sums = np.array(ref.face_counts), ref.checksums()
                whole.close()
                bad = 0
                for ids, fc, sums in parts:
                    bad += int((fc != want_faces[ids]).sum()) + int((sums != want_sums[ids]).any(axis=1).sum())
                res["vs_single_gpu"] = {"chunks": len(sp.plan.grid), "mismatching_chunks": bad, "what": "face count + checksums of vertex bytes and index bytes per chunk",
                                        "single_gpu_step_ms_same_box": statistics.median(times[1:])}
                if bad:
                    res["status"] = f"MISMATCH vs single GPU in {bad} chunks"
            dist.barrier()  # rank 0 was busy alone: nobody starts an exchange round (5 s device-side time-out) before it is back
        # the golden 5^3 planet through the same sharded path
        g = np.load(os.path.join(ROOT, "tests", "golden", "planet_42_5x5x5.npz"))
        from thisspaceofmine_b200.sharding import ShardedPlanet

        small = ShardedPlanet(ctx, tuple(int(v) for v in g["count"]), rank, world)
        small.GenerateChunks(lib, int(g["seed"]))
        small.PullHalos()
        b = small.BuildMeshes(views=True)
        ids = small.plan.local_ids(rank)
        ok = True
        if len(ids):
            blocks = small.planet.GetContent(small.local_grid)
            ok &= np.array_equal(np.array([crc(x) for x in blocks], np.uint32), g["block_crc"][ids])
            ok &= np.array_equal(np.array(b.face_counts), g["face_count"][ids])
            for i in range(len(ids)):
                m = b.chunk(i)
                ok &= (0 if m is None else crc(m.indices)) == int(g["index_crc"][ids[i]])
        oks = gather(bool(ok))
        small.close()
        res["vs_golden_5x5x5"] = "ok" if all(oks) else "MISMATCH"
        if not all(oks):
            res["status"] = "MISMATCH vs tests/golden/planet_42_5x5x5.npz"
        if world == 1:
            # one GPU: the single-process multi-part planet (3 parts on this device) against the plain planet, 31^3
            e = 31
            grid = T.Planet.chunk_grid((e,) * 3)
            one = T.Planet(ctx, capacity=e ** 3)
            one.GenerateChunks(lib, PLANET_SEED, (e,) * 3)
            ref = one.BuildMeshes(grid, views=True)
            want_faces, want_sums = np.array(ref.face_counts), ref.checksums()
            one.close()
            mp = T.MultiGpuPlanet([ctx.device] * 3, (e,) * 3)
            mp.GenerateChunks(PLANET_SEED)
            views, where, flags = mp.BuildMeshes(render=True)
            bad = int((views["face_count"] != want_faces).sum())
            for part in range(mp.parts):
                sel = np.nonzero(where == part)[0]
                sums = np.zeros((len(sel), 2), np.uint64)
                v = np.ascontiguousarray(views[sel])
                L.check(L.lib().tsom_mesh_checksums(L.lib().tsom_sharded_ctx(mp._h, part), C.cast(v.ctypes.data, C.POINTER(L.MeshView)), len(v), sums.ctypes.data_as(C.POINTER(C.c_uint64))))
                bad += int((sums != want_sums[sel]).any(axis=1).sum())
            mp.close()
            res["vs_single_gpu"] = {"chunks": e ** 3, "mismatching_chunks": bad, "what": "tsom_planet_create_sharded with 3 parts on this GPU vs the plain planet: face count + checksums per chunk"}
            if bad:
                res["status"] = f"MISMATCH vs single GPU in {bad} chunks"
    except Exception as ex:  # never lose the headline line
        res["status"] = f"error: {ex}"
    return res


# ---------------------------------------------------------------------------------------------- N = 1 extras
def extras(args, T, L, ctx, sp, out, torch, device):
    peak, _ = measured_peak_gbs()
    if not args.no_cpu:
        steps, kind, threads = cpu_planet_sample(args.edge, 3, 1)
        sec = sum(s["gen_s"] + s["mesh_s"] for s in steps) / len(steps)
        r = steps[-1]
        out["cpu_baseline"] = {"value": r["chunks"] / sec, "unit": "chunks/s", "cores": threads, "kind": kind, "sample": sample_note(kind, threads, r),
                               "gen_ms": 1e3 * sum(s["gen_s"] for s in steps) / len(steps), "mesh_ms": 1e3 * sum(s["mesh_s"] for s in steps) / len(steps)}
    if args.lean:
        return
    sp.planet.close()  # the 47^3 blocks are not needed any more
    for name, fn in (("microbench", lambda: microbench(args, T, ctx, torch, device, peak)), ("planet31", lambda: planet_bench(T, ctx, 31, peak)),
                     ("default_planet", lambda: default_planet_bench(T, ctx, not args.no_cpu)), ("box_colliders", lambda: box_collider_bench(T, ctx, not args.no_cpu)),
                     ("edits", lambda: edits_bench(T, ctx, args))):
        try:
            out[name] = fn()
        except Exception as e:  # secondary measurements: never lose the headline line
            out[name] = {"error": str(e)}


def microbench(args, T, ctx, torch, device, peak):
    """BASELINE config[1]: 4096 independent 32^3 chunks, random solid/air fill p = 0.5 (stone), mesh-only, render attributes — the kernel
    roofline microbench (round 1's headline).  Kernel time from the library's own CUDA events, mean of 10 launches after 3."""
    n = MICRO_CHUNKS
    cont = T.ChunkContainer(ctx, n, 1.0)
    idx = np.zeros((n, 3), np.int32)
    idx[:, 0] = 2 * np.arange(n)  # every other chunk index along X: no face-adjacent chunk -> independent chunks, all boundary faces drawn
    for s in range(0, n, 512):
        m = min(512, n - s)
        blk = device_fill(torch, device, SEED, s, m)
        torch.cuda.synchronize()
        cont.ResetChunks(idx[s:s + m], int(blk.data_ptr()))
        del blk
    from helpers import hash_fill
    assert np.array_equal(cont.GetContent(idx[:1])[0], hash_fill(SEED, 0)), "device fill != host fill"

    def timed(sub, reps=10, warm=3):
        faces = int(cont.BuildMeshes(sub, views=True).total_faces)
        ms = []
        for it in range(warm + reps):
            cont.BuildMeshes(sub, views=False)
            if it >= warm:
                ms.append(ctx.timings()["mesh_emit_ms"])
        return faces, sum(ms) / len(ms)

    faces, t = timed(idx)
    algo = mesh_algo_bytes(n, 0, faces)
    res = {"workload": "config[1]: 4096 independent 32^3 chunks, random solid/air fill p=0.5 (stone), mesh-only, render attributes", "chunks": n, "faces": faces, "kernel_ms": t,
           "value": n / (t / 1e3), "unit": "chunks/s",
           "roofline": {"bound": "hbm", "kernel": "mesh_fused_kernel", "achieved": algo / (t / 1e3) / 1e9, "peak": peak, "unit": "GB/s", "frac": algo / (t / 1e3) / 1e9 / peak,
                        "algorithmic_bytes_per_launch": algo, "traffic": ncu_traffic("config1_4096"),
                        "note": "the measured peak is a COPY (half reads, half writes); this launch is 94 % writes issued as streaming stores, so it can land a little above it"}}
    # the two other fills of SURVEY.md §8(d) config 2 (1024 chunks): sparse p = 0.05, and ids uniform over {0, 2, 3, 7, 8, 9, 13}
    m = 1024
    sub = np.ascontiguousarray(idx[:m])
    table = torch.tensor([0, 2, 3, 7, 8, 9, 13], dtype=torch.int16, device=device)

    def mixed(first, count):
        h = torch.arange(first * NBLK, (first + count) * NBLK, dtype=torch.int64, device=device).view(count, NBLK)
        h = (h * _s64(0x9E3779B97F4A7C15)) ^ ((h >> 29) & 0x7FFFFFFFF)
        h = h * _s64(0xBF58476D1CE4E5B9)
        return table[((h >> 33) & 0x7FFFFFFF) % 7].contiguous()

    variants = {}
    for name, fill in {"sparse_p0.05": lambda f, c: device_fill(torch, device, SEED + 2, f, c, p=0.05), "mixed_ids": mixed}.items():
        for s0 in range(0, m, 256):
            blk = fill(s0, 256)
            torch.cuda.synchronize()
            cont.ResetChunks(sub[s0:s0 + 256], int(blk.data_ptr()))
            del blk
        f, t = timed(sub, reps=5, warm=2)
        a = mesh_algo_bytes(m, 0, f)
        variants[name] = {"chunks": m, "faces": f, "kernel_ms": t, "chunks_per_s": m / (t / 1e3), "gbs": a / (t / 1e3) / 1e9, "frac": a / (t / 1e3) / 1e9 / peak}
    res["variants"] = variants
    cont.close()
    return res


def planet_bench(T, ctx, count, peak):
    """Full planet regenerate + remesh on one GPU, flat render and deformed render + collider (kernel times + wall clock of the two calls)."""
    cc = (count,) * 3
    n = count ** 3
    planet = T.Planet(ctx, capacity=n)
    grid = T.Planet.chunk_grid(cc)
    lib = ctx.blockLibrary
    nbr = int(neighbour_counts(grid, count).sum())
    res = {"chunk_count": list(cc), "chunks": n}
    for name, kw in (("flat_render", dict(render=True)), ("deformed_render_collider", dict(render=True, collider=True, deformed=True, deformRadius=16.0))):
        rows, faces = [], 0
        for it in range(5):
            w0 = time.perf_counter()
            planet.GenerateChunks(lib, PLANET_SEED, cc)
            t = ctx.timings()
            batch = planet.BuildMeshes(grid, views=(it == 0), **kw)
            w = (time.perf_counter() - w0) * 1e3
            t2 = ctx.timings()
            if it == 0:
                faces = int(batch.total_faces)
                with_faces = int((batch.face_counts > 0).sum())
                continue
            rows.append((t["terrain_ms"], t["generate_ms"], t2["mesh_emit_ms"], w))
        te, ge, me, w = (float(v) for v in np.mean(np.array(rows), axis=0))
        mb = mesh_algo_bytes(n, nbr, faces, with_collider="collider" in kw)
        res[name] = {"faces": faces, "chunks_with_faces": with_faces, "terrain_ms": te, "generate_ms": ge, "mesh_ms": me, "calls_wall_ms": w,
                     "gen_frac": n * 65536 / ((te + ge) / 1e3) / 1e9 / peak, "mesh_frac": mb / (me / 1e3) / 1e9 / peak,
                     "gen_plus_mesh_frac_by_wall": (n * 65536 + mb) / (w / 1e3) / 1e9 / peak, "gen_plus_mesh_chunks_per_s_wall": n / (w / 1e3)}
    planet.close()
    return res


PLATFORMS = ((5, (0, 71, 0)), (0, (3, 5, 70)), (3, (-70, 1, 2)), (1, (2, -71, 1)))  # (up direction, centre): one platform on four sides of the default planet


def default_planet_bench(T, ctx, with_cpu):
    """BASELINE config[0]: the default serverconfig spawn planet (seed 42, 5x5x5 chunks, src/Server/main.cpp:62,
    src/ServerLib/ServerPlanetEnvironment.cpp:49-63): generate + 4 platforms + mesh of every chunk, as the server does at start-up.
    GPU: wall clock around the six C-ABI calls (125 chunks: launch-latency bound).  CPU: the reference's own code on all host threads."""
    cc = (5, 5, 5)
    planet = T.Planet(ctx, capacity=125)
    grid = T.Planet.chunk_grid(cc)
    lib = ctx.blockLibrary

    def run():
        planet.GenerateChunks(lib, PLANET_SEED, cc)
        for d, c in PLATFORMS:
            planet.GeneratePlatform(lib, d, c)
        return planet.BuildMeshes(grid, views=False)

    run()
    faces = int(planet.BuildMeshes(grid, views=True).total_faces)
    times = []
    for _ in range(20):
        ctx.synchronize()
        t0 = time.perf_counter()
        run()
        ctx.synchronize()
        times.append(time.perf_counter() - t0)
    planet.close()
    sec = statistics.median(times)
    res = {"workload": "gen + 4 platforms + mesh", "chunk_count": list(cc), "chunks": 125, "faces": faces, "gpu_ms": sec * 1e3, "gpu_chunks_per_s": 125 / sec}
    if with_cpu:
        O, kind = cpu_arm()
        threads = host_threads()
        best = None
        for _ in range(3):
            p = O.Planet()
            t0 = time.perf_counter()
            assert p.generate(PLANET_SEED, cc, nthreads=threads)
            for d, c in PLATFORMS:
                p.generate_platform(d, c)
            t1 = time.perf_counter()
            r = O.bench_planet(PLANET_SEED, cc, threads)  # its mesh leg: every chunk, one task per chunk
            sec_cpu = (t1 - t0) + r["mesh_s"]
            best = sec_cpu if best is None else min(best, sec_cpu)
        res["cpu_baseline"] = {"kind": kind, "cores": threads, "ms": best * 1e3, "chunks_per_s": 125 / best}
    return res


def box_collider_bench(T, ctx, with_cpu):
    """FlatChunk::BuildCollider for every chunk of the 31^3 planet (the server's physics path, ChunkEntities.cpp:198) in ONE batched call."""
    e = 31
    cc = (e,) * 3
    grid = T.Planet.chunk_grid(cc)
    planet = T.Planet(ctx, capacity=e ** 3)
    planet.GenerateChunks(ctx.blockLibrary, PLANET_SEED, cc)
    planet.BuildBoxColliders(grid, copy=False)
    wall, kern = [], []
    for _ in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        first, cnt, _ = planet.BuildBoxColliders(grid, copy=False)
        wall.append((time.perf_counter() - t0) * 1e3)
        kern.append(ctx.timings()["collider_ms"])
    res = {"chunks": e ** 3, "boxes": int(cnt.sum()), "call_ms": statistics.median(wall), "kernels_ms": statistics.median(kern), "chunks_per_s": e ** 3 / (statistics.median(wall) / 1e3)}
    planet.close()
    if with_cpu:
        O, kind = cpu_arm()
        threads = host_threads()
        sc = (15, 15, 15)  # the CPU arm builds every chunk of a 15^3 planet (3,375 chunks): the per-chunk rate is what is compared
        p = O.Planet()
        assert p.generate(PLANET_SEED, sc, nthreads=threads)
        t, boxes, chunks = O.bench_box_colliders(p, sc, threads)
        res["cpu_baseline"] = {"kind": kind, "cores": threads, "chunks": chunks, "boxes": boxes, "seconds": t, "chunks_per_s": chunks / t, "planet": list(sc)}
    return res


def edits_bench(T, ctx, args):
    """BASELINE config[3]: 10 k random place/break ops per tick on the 31^3 planet, remesh dirty chunks + neighbours."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_edits

    ea = argparse.Namespace(edge=31, edits=10000, ticks=60, warmup=5, cpu_ticks=0 if args.no_cpu else 1, seed=42)
    return bench_edits.run(T, ctx, ea)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--edge", type=int, default=PLANET_EDGE, help="edge of the benchmark planet in chunks")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ramp-seconds", type=float, default=1.0, help="untimed steps before the warm-up so an idle box has ramped its clocks")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lean", action="store_true", help="headline line only (no N = 1 sub-objects)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
