"""GPU: TSOM_MESH_TANGENTS — what Nz::StaticMesh::GenerateTangents adds to a chunk mesh right after Chunk::BuildMesh
(src/ClientLib/ClientChunkEntities.cpp:169), fused into the meshing kernel.  NazaraEngine is not part of the reference tree: the
algorithm is restated once for the CPU in oracle/nazara_tangents.hpp (UNPINNED third-party arithmetic, see its header) and this test
holds the kernel to that restatement; everything else of the vertex stays exactly what the call without the flag writes."""
import numpy as np
import pytest

import oracle as O
import thisspaceofmine_b200 as T
from helpers import ATOL, RTOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [dict(), dict(genericMath=True), dict(deformed=True, deformRadius=16.0)], ids=["flat_table", "flat_generic", "deformed"])
def test_tangent_slot_matches_generate_tangents(ctx, kw):
    count = (3, 3, 3)
    grid = T.Planet.chunk_grid(count)
    planet = T.Planet(ctx, capacity=27)
    want = O.Planet()
    try:
        planet.GenerateChunks(ctx.blockLibrary, 42, count)
        assert want.generate(42, count, nthreads=4)
        planet.UpdateBlocks(np.array([[0, 1, 0, 5, 5, 20, 13], [0, 1, 0, 6, 5, 20, 9]]))  # glass (double sided) and forcefield on the surface
        want.apply_edits(np.array([[0, 1, 0, 5, 5, 20, 13], [0, 1, 0, 6, 5, 20, 9]]))
        plain = planet.BuildMeshes(grid, render=True, **kw)
        plain_meshes = [plain.chunk(i) for i in range(len(grid))]
        batch = planet.BuildMeshes(grid, render=True, tangents=True, **kw)
        faces = 0
        for i, idx in enumerate(grid):
            got = batch.chunk(i)
            ref = want.mesh(idx, deformed=bool(kw.get("deformed")), deform_radius=kw.get("deformRadius", 0.0),
                            deform_center=None if not kw.get("deformed") else -np.array(idx, np.float32) * np.float32(32.0))
            if ref is None:
                assert got is None
                continue
            assert np.array_equal(got.indices, ref.indices)
            # the other nine floats of every vertex are untouched by the flag
            assert np.array_equal(got.vertices[:, 0:9], plain_meshes[i].vertices[:, 0:9]) and not plain_meshes[i].tangent.any()
            # tangents of the GPU's own positions / uvs through the CPU restatement: same arithmetic, same inputs
            mine = O.Mesh(got.position.copy(), got.normal.copy(), got.uvw.copy(), got.indices)
            np.testing.assert_allclose(got.tangent, O.generate_tangents(mine), rtol=RTOL, atol=ATOL, err_msg=f"chunk {tuple(idx)}")
            # and of the oracle's mesh (positions / uvs within 1e-5 of ours)
            np.testing.assert_allclose(got.tangent, O.generate_tangents(ref), rtol=1e-4, atol=1e-4, err_msg=f"chunk {tuple(idx)} vs oracle mesh")
            lens = np.linalg.norm(got.tangent, axis=1)
            assert np.allclose(lens, 1.0, atol=1e-5), "unit tangents"
            if not kw.get("deformed"):  # a deformed quad is not planar: its vertex tangents need not be orthogonal to the face normal
                assert np.all(np.abs(np.einsum("ij,ij->i", got.tangent, got.normal)) < 1e-3), "a flat quad's tangent lies in its plane"
            faces += ref.face_count
        assert faces > 10000
    finally:
        planet.close()


def test_tangents_need_render_vertices(ctx):
    planet = T.Planet(ctx, capacity=1)
    try:
        planet.ResetChunks([(0, 0, 0)], np.full((1, 32768), 7, np.uint16))
        with pytest.raises(T.TsomError):
            planet.BuildMeshes([(0, 0, 0)], render=False, collider=True, tangents=True)
    finally:
        planet.close()
