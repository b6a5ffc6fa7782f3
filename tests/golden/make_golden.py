#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/tsom_oracle.hpp).

The reference ships no golden vectors for generation / meshing and cannot be built or imported here (C++20 + NazaraEngine), so
these fixtures freeze the ORACLE's output: they guard the oracle against regressions (tests/test_golden.py, CPU) and give the
GPU path a committed target that does not depend on the oracle being built on the GPU box (tests/test_gpu_golden.py).
Parity against the real reference stays "unpinned" (DESIGN.md).

    python tests/golden/make_golden.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
from helpers import hash_fill  # noqa: E402

PLATFORMS = [(O.RIGHT, (65, -18, -39)), (O.BACK, (-34, 2, 53)), (O.FRONT, (22, -35, -59)), (O.DOWN, (23, -62, 26))]  # ServerPlanetEnvironment.cpp:60-63


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def grid(count):
    nx, ny, nz = count
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return np.stack([x.ravel() - nx // 2, y.ravel() - ny // 2, z.ravel() - nz // 2], axis=1).astype(np.int32)


def planet_fixture(seed, count, platforms, keep):
    p = O.Planet(1.0, 16.0)
    assert p.generate(seed, count, nthreads=8)
    if platforms:
        for d, c in PLATFORMS:
            p.generate_platform(d, c)
    g = grid(count)
    out = dict(seed=np.uint32(seed), count=np.array(count, np.uint32), platforms=np.uint8(platforms), grid=g)
    out["block_crc"] = np.array([crc(p.get_blocks(i)) for i in g], np.uint32)
    faces, icrc = [], []
    for i in g:
        m = p.mesh(i)
        faces.append(0 if m is None else m.face_count)
        icrc.append(0 if m is None else crc(m.indices))
    out["face_count"] = np.array(faces, np.uint32)
    out["index_crc"] = np.array(icrc, np.uint32)
    for k, idx in enumerate(keep):
        m = p.mesh(idx)
        out[f"keep{k}_idx"] = np.array(idx, np.int32)
        out[f"keep{k}_blocks"] = p.get_blocks(idx)
        out[f"keep{k}_vertices"] = m.interleaved48()
        out[f"keep{k}_indices"] = m.indices
        out[f"keep{k}_boxes"] = p.box_collider(idx)
    return out


def isolated_fixture():
    """BASELINE config 2 inputs (hash fill, seed 1234) + a mixed-id chunk, flat and deformed."""
    out = {}
    p = O.Planet(1.0, 16.0)
    chunks = {(0, 0, 0): hash_fill(1234, 0), (2, 0, 0): hash_fill(1234, 1), (4, 0, 0): hash_fill(1234, 2, p=0.05)}
    mixed = hash_fill(77, 0, p=0.7).astype(np.uint16)
    pick = np.array([2, 3, 7, 8, 9, 13], np.uint16)[(np.arange(len(mixed)) * 2654435761 >> 7) % 6]
    chunks[(6, 0, 0)] = np.where(mixed != 0, pick, 0).astype(np.uint16)
    for idx, b in chunks.items():
        p.add_chunk(idx, b)
    keys = list(chunks)
    out["idx"] = np.array(keys, np.int32)
    out["blocks_crc"] = np.array([crc(chunks[k]) for k in keys], np.uint32)
    out["mixed_blocks"] = chunks[(6, 0, 0)]
    fc, ic, vc = [], [], []
    for k in keys:
        m = p.mesh(k)
        fc.append(m.face_count)
        ic.append(crc(m.indices))
        vc.append(crc(m.interleaved48()))
    out["face_count"] = np.array(fc, np.uint32)
    out["index_crc"] = np.array(ic, np.uint32)
    out["vertex_crc_oracle"] = np.array(vc, np.uint32)  # bit-exact only for the oracle itself
    m = p.mesh((4, 0, 0))
    out["sparse_vertices"] = m.interleaved48()
    out["sparse_indices"] = m.indices
    m = p.mesh((6, 0, 0))
    out["mixed_vertices_head"] = m.interleaved48()[:8000]  # first 2000 faces
    md = p.mesh((4, 0, 0), deformed=True, deform_radius=16.0)
    out["sparse_deformed_vertices"] = md.interleaved48()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "planet_42_5x5x5.npz"), **planet_fixture(42, (5, 5, 5), False, [(2, 0, 0), (0, 2, 0)]))
    np.savez_compressed(os.path.join(HERE, "planet_42_5x5x5_platforms.npz"), **planet_fixture(42, (5, 5, 5), True, [(2, -1, -2)]))
    np.savez_compressed(os.path.join(HERE, "planet_7_3x3x3.npz"), **planet_fixture(7, (3, 3, 3), False, [(0, 1, 0)]))
    np.savez_compressed(os.path.join(HERE, "isolated_chunks.npz"), **isolated_fixture())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
