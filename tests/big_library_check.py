"""Run in its OWN process (it appends to the oracle's process-wide BlockLibrary):  python tests/big_library_check.py cpu|gpu

A block library of 14 + 300 types with random collision / double-sided / transparent flags and per-direction textures, chunks
filled with ids from the whole range (so ids beyond the mesher's 128-entry shared-memory flag table take its slow path, and
"same type => no face" is exercised between many transparent types), neighbours present / without content / absent.
  cpu: the restated oracle vs the reference's own code (oracle/_ref), bit-exact;
  gpu: the CUDA mesher (through the C ABI) vs the oracle, indices bit-exact, attributes within 1e-5; box colliders identical."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle as O  # noqa: E402
from helpers import NBLK, assert_mesh_equal  # noqa: E402

EXTRA = 300


def extra_blocks(rng):
    out = []
    for i in range(EXTRA):
        kind = rng.integers(0, 4)
        out.append(dict(name=f"extra_{i}", basePath=f"blocks/extra_{i}", sidePath=f"blocks/extra_{i}_side" if kind == 1 else "",
                        upPath=f"blocks/extra_{i}_up" if kind == 2 else "", hasCollisions=bool(rng.random() < 0.8), isDoubleSided=bool(rng.random() < 0.2),
                        isTransparent=bool(rng.random() < 0.35)))
    return out


def make_chunks(rng, n_types):
    def fill(p, lo, hi):
        ids = rng.integers(lo, hi, NBLK).astype(np.uint16)
        return np.where(rng.random(NBLK) < p, ids, 0).astype(np.uint16)

    return {
        (0, 0, 0): fill(0.6, 1, n_types),          # everything, most ids >= 128
        (1, 0, 0): fill(0.5, 100, 160),            # straddles the table boundary
        (-1, 0, 0): fill(0.9, 200, n_types),
        (0, 1, 0): fill(0.3, 1, 14),               # built-in types only
        (0, 0, 1): None,                           # exists, no content
        (0, 0, -1): fill(0.7, 120, 136),
        (5, 5, 5): fill(0.5, n_types - 3, n_types),  # isolated, the last few ids
    }


def main(mode):
    rng = np.random.default_rng(77)
    blocks = extra_blocks(rng)
    for b in blocks:
        idx = O.register_block(**b)
    n_types = idx + 1
    assert n_types == 14 + EXTRA
    chunks = make_chunks(rng, n_types)

    op = O.Planet()
    for k, v in chunks.items():
        op.add_chunk(k, v)

    faces = 0
    if mode == "cpu":
        R = O.reference()
        for b in blocks:
            R.register_block(**b)
        a, c = O.block_library(), R.block_library()
        assert len(a) == len(c) == n_types
        for x, y in zip(a, c):
            assert x[0] == y[0] and list(x[1]) == list(y[1]) and x[2:] == y[2:], (x, y)
        rp = R.Planet()
        for k, v in chunks.items():
            rp.add_chunk(k, v)
        for k, v in chunks.items():
            ma, mb = op.mesh(k), rp.mesh(k)
            if v is None:
                assert ma is None and mb is None
                continue
            assert ma.face_count == mb.face_count and np.array_equal(ma.indices, mb.indices), k
            assert np.array_equal(ma.position, mb.position) and np.array_equal(ma.normal, mb.normal) and np.array_equal(ma.uvw, mb.uvw), k
            assert np.array_equal(op.box_collider(k), rp.box_collider(k)), k
            faces += ma.face_count
    else:
        import thisspaceofmine_b200 as T

        lib = T.BlockLibrary()
        for b in blocks:
            lib.RegisterBlock(b["name"], T.host.BlockInfo(basePath=b["basePath"], baseSidePath=b["sidePath"], baseUpPath=b["upPath"], hasCollisions=b["hasCollisions"],
                                                          isDoubleSided=b["isDoubleSided"], isTransparent=b["isTransparent"]))
        for i, e in enumerate(O.block_library()):
            d = lib.GetBlockData(i)
            assert (d.name, list(d.texIndices), d.hasCollisions, d.isDoubleSided, d.isTransparent) == (e[0], list(e[1]), e[2], e[3], e[4])
        ctx = T.Context(0, lib)
        gc = T.ChunkContainer(ctx, len(chunks), 1.0)
        for k, v in chunks.items():
            gc.AddChunk(k, None if v is None else v[None, :])
        keys = [k for k, v in chunks.items() if v is not None]
        for kw in (dict(render=True, collider=True), dict(render=True, deformed=True, deformRadius=16.0)):
            batch = gc.BuildMeshes(keys, **kw)
            for i, k in enumerate(keys):
                om = op.mesh(k, deformed=kw.get("deformed", False), deform_radius=kw.get("deformRadius", 0.0))
                faces += assert_mesh_equal(batch.chunk(i), om, what=f"{kw} chunk {k}")["faces"]
        for k in keys:
            assert np.array_equal(gc.BuildBoxCollider(k), op.box_collider(k)), k
        gc.close()
        ctx.close()
    assert faces > 100_000
    print(f"big library ok ({mode}): {n_types} block types, {faces} faces compared")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "cpu")
