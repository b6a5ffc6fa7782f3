"""thisspaceofmine_b200 — B200-native (sm_100a) voxel chunk pipeline for This Space Of Mine.

Chunk generation, face-culled meshing, collider triangles and block edits run as hand-written CUDA kernels behind the C ABI of
``include/tsom_b200.h`` (``libtsom_b200.so``); ``host.py`` mirrors the reference's BlockLibrary / Planet / Chunk call surface on top.
"""
from ._lib import SO_PATH, TsomError, build  # noqa: F401
from .host import (  # noqa: F401
    BlockData,
    BlockInfo,
    BlockLibrary,
    ChunkBlockCount,
    ChunkContainer,
    ChunkEntities,
    ChunkMesh,
    ChunkSize,
    Context,
    Direction,
    DirectionMask_All,
    EmptyBlockIndex,
    InvalidBlockIndex,
    MeshBatch,
    MultiGpuPlanet,
    Planet,
    Ship,
)
