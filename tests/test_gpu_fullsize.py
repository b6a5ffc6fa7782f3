"""GPU, BASELINE.json's FULL sizes, checked through size-independent properties (the CPU oracle would need minutes to hours here).

config[1]: 4096 independent random-fill chunks, 207.6 M faces, 45 GB of vertices + indices:
  * per-chunk face counts == an independent count written in plain torch tensor ops on the same ids (bit-exact);
  * every index of the 1.25 G-entry index buffer == 4 * (chunk-relative face ordinal) + {0,2,1,1,2,3} (bit-exact, Chunk.cpp:95-101);
  * every vertex: integer block-corner position inside the chunk, one axis-aligned unit normal shared by the face's 4 vertices,
    the 4 corners span a unit square in the plane of the normal, uv in {0,1} (to 1e-6), slice == the stone texture, tangent slot == 0;
  * every chunk's mesh is a closed surface: its normals sum to zero, per axis as many +faces as -faces.
configs[2]/[4]-sized planets (31^3 = 29,791 and 47^3 = 103,823 chunks) generated and meshed on the device:
  * generation invariants of SURVEY.md §8c(4) on the whole volume (free space, shaft, solid core);
  * per-chunk face counts == torch count on the assembled world volume with real cross-chunk neighbours (bit-exact);
  * regenerate + remesh is idempotent (same counts, same arena checksum)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


class _DevArray:
    """__cuda_array_interface__ view of a raw device pointer (the library's arenas / chunk storage) for torch.as_tensor."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def dev_tensor(torch, ptr, shape, typestr):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device="cuda")


def torch_face_counts_isolated(torch, ids):
    """ids [n, 32(z), 32(y), 32(x)] int16, opaque blocks only, no neighbour chunks -> faces per chunk (every solid block side that
    touches Empty or the chunk boundary)."""
    solid = ids != 0
    total = torch.zeros(ids.shape[0], dtype=torch.int64, device=ids.device)
    for dim in (1, 2, 3):
        for shift in (1, -1):
            nb = torch.zeros_like(solid)
            src = [slice(None)] * 4
            dst = [slice(None)] * 4
            if shift == 1:
                src[dim], dst[dim] = slice(1, None), slice(0, -1)  # neighbour at +1
            else:
                src[dim], dst[dim] = slice(0, -1), slice(1, None)  # neighbour at -1
            nb[tuple(dst)] = solid[tuple(src)]
            total += (solid & ~nb).sum(dim=(1, 2, 3))
    return total


def test_config1_full_size_properties(ctx):
    import torch

    import bench
    import thisspaceofmine_b200 as T
    from thisspaceofmine_b200 import _lib as L

    device = torch.device("cuda", 0)
    n = bench.CHUNKS_PER_GPU
    cont = T.ChunkContainer(ctx, n, 1.0)
    idx = np.zeros((n, 3), np.int32)
    idx[:, 0] = 2 * np.arange(n)
    want_counts = torch.empty(n, dtype=torch.int64, device=device)
    for s in range(0, n, 512):
        blk = bench.device_fill(torch, device, bench.SEED, s, 512)
        torch.cuda.synchronize()
        cont.ResetChunks(idx[s:s + 512], int(blk.data_ptr()))
        want_counts[s:s + 512] = torch_face_counts_isolated(torch, blk.view(512, 32, 32, 32))
        del blk

    batch = cont.BuildMeshes(idx, views=True)
    counts = torch.from_numpy(batch.face_counts.astype(np.int64)).to(device)
    firsts = torch.from_numpy(batch.first_faces.astype(np.int64)).to(device)
    assert torch.equal(counts, want_counts), "per-chunk face counts differ from the torch count"
    F = int(counts.sum())
    assert F == batch.total_faces == 207619226  # the workload's face count (bench.py)
    order = torch.argsort(firsts, stable=True)
    order = order[counts[order] > 0]
    assert torch.equal(firsts[order], torch.cumsum(counts[order], 0) - counts[order]), "the chunks' face ranges must tile the arena without gaps or overlaps"

    vptr, iptr, _, total = ctx.arenas()
    assert total == F
    verts = dev_tensor(torch, vptr, (F * 4, 12), "<f4")
    inds = dev_tensor(torch, iptr, (F, 6), "<i4")
    pattern = torch.tensor([0, 2, 1, 1, 2, 3], dtype=torch.int32, device=device)
    stone_tex = float(ctx.blockLibrary.GetBlockData(bench.STONE).texIndices[0])
    chunk_of_first = torch.repeat_interleave(order, counts[order])  # face -> chunk in arena order, F int64 (1.6 GB)
    normal_sum = torch.zeros((n, 3), dtype=torch.float64, device=device)
    STEP = 1 << 23
    for f0 in range(0, F, STEP):
        f1 = min(F, f0 + STEP)
        ch = chunk_of_first[f0:f1]
        rel = (torch.arange(f0, f1, device=device) - firsts[ch]).to(torch.int32)
        assert torch.equal(inds[f0:f1], rel[:, None] * 4 + pattern[None, :]), f"index buffer differs in faces [{f0}, {f1})"
        v = verts[f0 * 4:f1 * 4].view(-1, 4, 12)
        pos, nrm, uv, sl, tan = v[:, :, 0:3], v[:, :, 3:6], v[:, :, 6:8], v[:, :, 8], v[:, :, 9:12]
        assert not tan.any(), "tangent slot must stay zero"
        assert bool((sl == stone_tex).all()), "texture slice"
        uvr = uv.round()  # side faces carry the reference's own ~6e-8 quaternion rounding (identical in the CPU oracle)
        assert bool((((uvr == 0) | (uvr == 1)) & ((uv - uvr).abs() <= 1e-6)).all()), "uv corners"
        assert bool((pos == pos.round()).all()) and bool((pos.abs() <= 16).all()), "positions are block corners inside the chunk"
        assert bool((nrm == nrm[:, :1, :]).all()), "one normal per face"
        n0 = nrm[:, 0, :]
        assert bool(((n0.abs().sum(1) == 1) & ((n0 == 0) | (n0.abs() == 1)).all(1)).all()), "axis-aligned unit normals"
        ext = pos.max(1).values - pos.min(1).values                      # [f, 3] extent of the quad per axis
        assert bool((ext == (1 - n0.abs())).all()), "each quad is a unit square in the plane of its normal"
        normal_sum.index_add_(0, ch, n0.to(torch.float64))
    assert not normal_sum.any(), "every isolated chunk's mesh is a closed surface (normals sum to zero)"
    cont.close()


def _world_volume(torch, blocks, edge):
    """[edge^3 chunks in z, y, x chunk order][32 z][32 y][32 x] -> world volume [Y, Z, X] (local z is world Y, local y is world Z)."""
    v = blocks.view(edge, edge, edge, 32, 32, 32)      # cz, cy, cx, lz, ly, lx
    v = v.permute(1, 3, 0, 4, 2, 5)                    # cy, lz, cz, ly, cx, lx
    return v.reshape(edge * 32, edge * 32, edge * 32)  # world Y, world Z, world X


def _planet_counts_from_volume(torch, vol, edge):
    solid = vol != 0
    faces = torch.zeros(solid.shape, dtype=torch.uint8, device=vol.device)
    for dim in (0, 1, 2):
        for shift in (1, -1):
            nb = torch.zeros_like(solid)
            src, dst = [slice(None)] * 3, [slice(None)] * 3
            if shift == 1:
                src[dim], dst[dim] = slice(1, None), slice(0, -1)
            else:
                src[dim], dst[dim] = slice(0, -1), slice(1, None)
            nb[tuple(dst)] = solid[tuple(src)]
            faces += (solid & ~nb).to(torch.uint8)
            del nb
    per = faces.view(edge, 32, edge, 32, edge, 32).to(torch.int32).sum(dim=(1, 3, 5))  # [cy, cz, cx]
    return per.permute(1, 0, 2).reshape(-1).to(torch.int64)                             # chunk order z, y, x


@pytest.mark.parametrize("edge", [31, 47])
def test_planet_full_size_properties(ctx, edge):
    import torch

    import thisspaceofmine_b200 as T
    from thisspaceofmine_b200 import _lib as L
    import ctypes as C

    device = torch.device("cuda", 0)
    n = edge ** 3
    planet = T.Planet(ctx, capacity=n)
    cc = (edge, edge, edge)
    planet.GenerateChunks(ctx.blockLibrary, 42, cc)
    grid = T.Planet.chunk_grid(cc)
    first = np.ascontiguousarray(grid[0], np.int32)
    base = L.lib().tsom_chunk_device_ptr(planet._h, first.ctypes.data_as(C.POINTER(C.c_int32)))
    assert base
    blocks = dev_tensor(torch, base, (n, 32, 32, 32), "<i2")  # GenerateChunks fills slots in grid order
    lib = ctx.blockLibrary
    opaque = [lib.GetBlockIndex(k) for k in ("dirt", "grass", "stone", "stone_mossy", "snow")]
    assert bool(torch.isin(blocks, torch.tensor([0] + opaque, dtype=torch.int16, device=device)).all())

    vol = _world_volume(torch, blocks, edge).contiguous()
    half = (edge // 2) * 32 + 16            # world coordinate of voxel 0 is -half
    maxH = ((edge + 1) // 2) * 32
    coord = (torch.arange(edge * 32, device=device) - half).to(torch.int16)
    a = coord.abs()
    cheb = torch.maximum(torch.maximum(a[:, None, None], a[None, :, None]), a[None, None, :])
    assert not vol[cheb > maxH - 30].any(), "free space: depth < 30 is Empty (Planet.cpp:205-209)"
    shaft = (a[None, :, None] <= 2) & (a[None, None, :] <= 2)  # |x| <= 2 and |z| <= 2 (Planet.cpp:221-222)
    assert not vol[shaft.expand_as(vol)].any(), "central shaft is Empty"
    core = cheb <= maxH // 2 - 2
    assert bool((vol[core & ~shaft.expand_as(vol)] != 0).all()), "below the deepest terrain everything is solid"
    del cheb, core

    want = _planet_counts_from_volume(torch, vol, edge)
    del vol
    batch = planet.BuildMeshes(grid, views=True)
    got = torch.from_numpy(batch.face_counts.astype(np.int64)).to(device)
    assert torch.equal(got, want), f"{int((got != want).sum())} chunks differ from the torch count"
    F = int(got.sum())
    assert F > 10 * n

    def checksum():
        vptr, iptr, _, total = ctx.arenas()
        assert total == F
        acc = 0
        for ptr, words in ((iptr, F * 6), (vptr, F * 48)):  # both arenas as 32-bit words, summed exactly in int64
            t = dev_tensor(torch, ptr, (words,), "<i4")
            for w0 in range(0, words, 1 << 28):
                acc += int(t[w0:w0 + (1 << 28)].to(torch.int64).sum())
        return acc

    ck1 = checksum()
    planet.GenerateChunks(lib, 42, cc)
    batch2 = planet.BuildMeshes(grid, views=True)
    assert np.array_equal(batch2.face_counts, batch.face_counts)
    assert checksum() == ck1, "regenerate + remesh must be idempotent"
    planet.close()


def test_deformed_planet_shortcuts_equal_generic_math_on_every_chunk(ctx):
    """31^3 planet, DeformedChunk corners (radius 16) + collider: the run that sends undeformed chunks / blocks through the flat
    path's attribute table must write the same bytes as the run that evaluates DrawFace for every face (TSOM_MESH_GENERIC_MATH) —
    device checksums of every chunk's vertex words and indices, plus the face counts; and the same for the collider positions of a
    collider-only run."""
    import thisspaceofmine_b200 as T

    cc = (31, 31, 31)
    planet = T.Planet(ctx, capacity=31 ** 3)
    try:
        planet.GenerateChunks(ctx.blockLibrary, 42, cc)
        grid = T.Planet.chunk_grid(cc)
        for kw in (dict(render=True, collider=True), dict(render=False, collider=True)):
            fast = planet.BuildMeshes(grid, deformed=True, deformRadius=16.0, ordered=True, **kw)
            fc, fs = np.asarray(fast.face_counts).copy(), fast.checksums()
            gen = planet.BuildMeshes(grid, deformed=True, deformRadius=16.0, ordered=True, genericMath=True, **kw)
            assert np.array_equal(fc, gen.face_counts)
            gs = gen.checksums()
            bad = np.nonzero((fs != gs).any(axis=1))[0]
            assert len(bad) == 0, f"{kw}: {len(bad)} chunks differ, first {grid[bad[0]].tolist()}"
            assert int(fc.sum()) == 53483600
    finally:
        planet.close()
