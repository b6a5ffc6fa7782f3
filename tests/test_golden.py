"""CPU: the oracle must keep reproducing the committed golden fixtures (tests/golden/make_golden.py wrote them from the oracle;
the reference ships none for generation / meshing)."""
import os
import zlib

import numpy as np
import pytest

import oracle as O
from helpers import hash_fill

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLATFORMS = [(O.RIGHT, (65, -18, -39)), (O.BACK, (-34, 2, 53)), (O.FRONT, (22, -35, -59)), (O.DOWN, (23, -62, 26))]


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


@pytest.mark.parametrize("name", ["planet_42_5x5x5", "planet_42_5x5x5_platforms", "planet_7_3x3x3"])
def test_planet_fixture(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = O.Planet(1.0, 16.0)
    assert p.generate(int(g["seed"]), tuple(int(v) for v in g["count"]), nthreads=8)
    if int(g["platforms"]):
        for d, c in PLATFORMS:
            p.generate_platform(d, c)
    for i, idx in enumerate(g["grid"]):
        assert crc(p.get_blocks(idx)) == g["block_crc"][i], f"blocks of chunk {tuple(idx)}"
        m = p.mesh(idx)
        assert (0 if m is None else m.face_count) == g["face_count"][i]
        assert (0 if m is None else crc(m.indices)) == g["index_crc"][i]
    k = 0
    while f"keep{k}_idx" in g:
        idx = g[f"keep{k}_idx"]
        assert np.array_equal(p.get_blocks(idx), g[f"keep{k}_blocks"])
        m = p.mesh(idx)
        assert np.array_equal(m.interleaved48(), g[f"keep{k}_vertices"])
        assert np.array_equal(m.indices, g[f"keep{k}_indices"])
        assert np.array_equal(p.box_collider(idx), g[f"keep{k}_boxes"])
        k += 1
    assert k >= 1


def test_isolated_fixture():
    g = np.load(os.path.join(GOLDEN, "isolated_chunks.npz"))
    chunks = [hash_fill(1234, 0), hash_fill(1234, 1), hash_fill(1234, 2, p=0.05), g["mixed_blocks"]]
    p = O.Planet(1.0, 16.0)
    for idx, b in zip(g["idx"], chunks):
        assert crc(b) == g["blocks_crc"][list(map(tuple, g["idx"])).index(tuple(idx))]
        p.add_chunk(idx, b)
    for i, idx in enumerate(g["idx"]):
        m = p.mesh(idx)
        assert m.face_count == g["face_count"][i]
        assert crc(m.indices) == g["index_crc"][i]
        assert crc(m.interleaved48()) == g["vertex_crc_oracle"][i]
    assert np.array_equal(p.mesh(g["idx"][2]).interleaved48(), g["sparse_vertices"])
    assert np.array_equal(p.mesh(g["idx"][3]).interleaved48()[:8000], g["mixed_vertices_head"])
    assert np.array_equal(p.mesh(g["idx"][2], deformed=True, deform_radius=16.0).interleaved48(), g["sparse_deformed_vertices"])
    # BASELINE config 2 sanity: p = 0.5 isolated chunk ~ 50.7 k faces
    assert 50000 < g["face_count"][0] < 51500
