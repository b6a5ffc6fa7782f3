"""Host-side mirror of the reference's call surface for the chunk pipeline, over the C ABI of libtsom_b200.so.

Names, argument meaning and error behaviour follow the reference classes (method names keep the reference's CamelCase):

  BlockLibrary      include/CommonLib/BlockLibrary.hpp, src/CommonLib/BlockLibrary.cpp
  ChunkContainer    include/CommonLib/ChunkContainer.hpp/.inl
  Planet            include/CommonLib/Planet.hpp, src/CommonLib/Planet.cpp
  Chunk             include/CommonLib/Chunk.hpp, src/CommonLib/Chunk.cpp (FlatChunk / DeformedChunk corner providers)
  ChunkEntities     src/CommonLib/ChunkEntities.cpp + src/ClientLib/ClientChunkEntities.cpp (per-tick batch of invalidated chunks)

All block data lives in HBM; every compute call runs sm_100a kernels.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from dataclasses import dataclass, field
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _lib as L

ChunkSize = 32
ChunkBlockCount = ChunkSize ** 3
EmptyBlockIndex = 0
InvalidBlockIndex = 0xFFFF


class Direction:
    """include/CommonLib/Direction.hpp:18-28"""
    Back, Down, Front, Left, Right, Up = range(6)


DirectionMask_All = 0x3F

# Direction.hpp:83-90 (world axes)
s_chunkDirOffset = np.array([[0, 0, 1], [0, -1, 0], [0, 0, -1], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.int32)


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


# numpy mirrors of tsom_edit / tsom_mesh_view (include/tsom_b200.h), so batches cross the C ABI without per-element Python work
EDIT_DTYPE = np.dtype([("chunk", np.int32, 3), ("x", np.uint8), ("y", np.uint8), ("z", np.uint8), ("reserved", np.uint8), ("block", np.uint16), ("reserved2", np.uint16)])
VIEW_DTYPE = np.dtype([("first_face", np.uint64), ("face_count", np.uint32), ("reserved", np.uint32)])
BOX_VIEW_DTYPE = np.dtype([("first_box", np.uint64), ("box_count", np.uint32), ("reserved", np.uint32)])


def pack_edits(edits) -> np.ndarray:
    """[n, 7] ints (cx, cy, cz, x, y, z, block) -> tsom_edit records."""
    e = np.ascontiguousarray(edits, dtype=np.int64).reshape(-1, 7)
    raw = np.zeros(len(e), EDIT_DTYPE)
    raw["chunk"] = e[:, 0:3]
    raw["x"], raw["y"], raw["z"], raw["block"] = e[:, 3], e[:, 4], e[:, 5], e[:, 6]
    return raw


def _mesh_flags(render, collider, deformed, genericMath, ordered, tangents) -> int:
    return ((L.MESH_RENDER if render else 0) | (L.MESH_COLLIDER if collider else 0) | (L.MESH_DEFORMED if deformed else 0) | (L.MESH_GENERIC_MATH if genericMath else 0)
            | (L.MESH_ORDERED if ordered else 0) | (L.MESH_TANGENTS if tangents else 0))


def _idx(v) -> np.ndarray:
    a = np.ascontiguousarray(v, dtype=np.int32).reshape(-1, 3)
    return a


# ------------------------------------------------------------------------------------------------ BlockLibrary
@dataclass
class BlockInfo:
    """BlockLibrary::BlockInfo (BlockLibrary.hpp:37-50)"""
    basePath: str = ""
    baseBackPath: str = ""
    baseDownPath: str = ""
    baseFrontPath: str = ""
    baseLeftPath: str = ""
    baseRightPath: str = ""
    baseSidePath: str = ""
    baseUpPath: str = ""
    hasCollisions: bool = True
    isDoubleSided: bool = False
    isTransparent: bool = False
    permeability: float = 0.0


@dataclass
class BlockData:
    """BlockLibrary::BlockData (BlockLibrary.hpp:52-60)"""
    name: str
    texIndices: list = field(default_factory=lambda: [0] * 6)
    hasCollisions: bool = True
    isDoubleSided: bool = False
    isTransparent: bool = False
    permeability: float = 0.0


class BlockLibrary:
    """The 14 built-in block types and their texture slices (BlockLibrary.cpp:10-143)."""

    def __init__(self, register_defaults: bool = True):
        self._blocks: list[BlockData] = []
        self._blockIndices: dict[str, int] = {}
        self._textureIndices: dict[str, int] = {}
        if register_defaults:
            R = self.RegisterBlock
            R("empty", BlockInfo(hasCollisions=False, isTransparent=True, permeability=1.0))
            R("debug", BlockInfo(baseBackPath="blocks/debug_back", baseDownPath="blocks/debug_down", baseFrontPath="blocks/debug_front",
                                 baseLeftPath="blocks/debug_left", baseRightPath="blocks/debug_right", baseUpPath="blocks/debug_up"))
            R("dirt", BlockInfo(basePath="blocks/dirt", permeability=0.1))
            R("grass", BlockInfo(basePath="blocks/grass_top", baseDownPath="blocks/dirt", baseSidePath="blocks/grass_side", permeability=0.1))
            R("hull", BlockInfo(basePath="blocks/smooth_stone"))
            R("hull2", BlockInfo(basePath="blocks/smooth_stone_slab_side"))
            R("snow", BlockInfo(basePath="blocks/snow", permeability=0.5))
            R("stone", BlockInfo(basePath="blocks/cobblestone"))
            R("stone_mossy", BlockInfo(basePath="blocks/mossy_cobblestone"))
            R("forcefield", BlockInfo(basePath="blocks/forcefield", hasCollisions=False, isTransparent=True))
            R("planks", BlockInfo(basePath="blocks/planks"))
            R("stone_bricks", BlockInfo(basePath="blocks/stone_bricks"))
            R("copper_block", BlockInfo(basePath="blocks/copper_block"))
            R("glass", BlockInfo(basePath="blocks/glass", isDoubleSided=True, isTransparent=True))

    def RegisterBlock(self, name: str, info: BlockInfo) -> int:
        index = len(self._blocks)
        data = BlockData(name=name, hasCollisions=info.hasCollisions, isDoubleSided=info.isDoubleSided, isTransparent=info.isTransparent, permeability=info.permeability)
        base = self._RegisterTexture(info.basePath) if info.basePath else 0
        side = self._RegisterTexture(info.baseSidePath) if info.baseSidePath else base
        per_dir = [info.baseBackPath, info.baseDownPath, info.baseFrontPath, info.baseLeftPath, info.baseRightPath, info.baseUpPath]
        for d, path in enumerate(per_dir):
            if path:
                data.texIndices[d] = self._RegisterTexture(path)
            elif d not in (Direction.Up, Direction.Down):
                data.texIndices[d] = side
            else:
                data.texIndices[d] = base
        assert name not in self._blockIndices
        self._blocks.append(data)
        self._blockIndices[name] = index
        return index

    def _RegisterTexture(self, path: str) -> int:
        if path in self._textureIndices:
            return self._textureIndices[path]
        tex = len(self._textureIndices) + 1  # slice #0 stays empty
        self._textureIndices[path] = tex
        return tex

    def GetBlockData(self, blockIndex: int) -> BlockData:
        return self._blocks[blockIndex]

    def GetBlockIndex(self, blockName: str) -> int:
        return self._blockIndices.get(blockName, InvalidBlockIndex)

    def __len__(self):
        return len(self._blocks)

    def to_descs(self):
        arr = (L.BlockDesc * len(self._blocks))()
        for i, b in enumerate(self._blocks):
            for d in range(6):
                arr[i].tex[d] = b.texIndices[d]
            arr[i].has_collisions = int(b.hasCollisions)
            arr[i].is_double_sided = int(b.isDoubleSided)
            arr[i].is_transparent = int(b.isTransparent)
        return arr


def _timings_dict(t) -> dict:
    """enum tsom_timing (include/tsom_b200.h) -> names"""
    return dict(terrain_ms=t[0], generate_ms=t[1], halo_ms=t[2], classify_ms=t[3], mesh_emit_ms=t[4], mesh_reemit=bool(t[5]), io_ms=t[6], collider_ms=t[7])


# ------------------------------------------------------------------------------------------------ device context
class Context:
    """One per process and GPU (tsom_ctx): stream, device block table, mesh arenas."""

    def __init__(self, device: int = 0, blockLibrary: Optional[BlockLibrary] = None, _borrowed=None):
        self._owned = _borrowed is None
        self.device = device
        self.blockLibrary = blockLibrary or BlockLibrary()
        if _borrowed is not None:  # a context owned by a tsom_sharded_planet
            self._h = C.c_void_p(_borrowed)
            return
        self._h = C.c_void_p()
        L.check(L.lib().tsom_create(device, C.byref(self._h)))
        descs = self.blockLibrary.to_descs()
        L.check(L.lib().tsom_set_block_table(self._h, descs, len(descs)))

    def close(self):
        if self._h:
            if self._owned:
                L.lib().tsom_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return int(L.lib().tsom_stream(self._h) or 0)

    def synchronize(self):
        L.check(L.lib().tsom_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(L.lib().tsom_launch_count(self._h))

    def timings(self) -> dict:
        """Device-side kernel durations (ms) of the most recent generate / mesh call on this context."""
        t = (C.c_float * L.TIMING_COUNT)()
        L.check(L.lib().tsom_get_timings(self._h, t, L.TIMING_COUNT))
        return _timings_dict(t)

    def reserve_faces(self, faces: int, render=True, collider=False):
        L.check(L.lib().tsom_reserve_faces(self._h, int(faces), (L.MESH_RENDER if render else 0) | (L.MESH_COLLIDER if collider else 0)))

    @contextlib.contextmanager
    def locked(self):
        """tsom_ctx_lock / tsom_ctx_unlock: keeps "mesh, then read the arenas" together when several threads share the context."""
        L.check(L.lib().tsom_ctx_lock(self._h))
        try:
            yield self
        finally:
            L.lib().tsom_ctx_unlock(self._h)

    def check_guards(self):
        """Raises if a kernel wrote behind one of the mesh arenas (each carries a 0xA5 guard behind its last face)."""
        L.check(L.lib().tsom_debug_check_guards(self._h))

    def arenas(self):
        v, i, c, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        L.check(L.lib().tsom_mesh_arenas(self._h, C.byref(v), C.byref(i), C.byref(c), C.byref(n)))
        return v.value, i.value, c.value, int(n.value)


@dataclass
class ChunkMesh:
    """What ClientChunkEntities::BuildMesh / DeformedChunk::BuildCollider hand to Nazara, as host arrays."""
    vertices: Optional[np.ndarray]            # [4F, 12] float32: position, normal, uvw, tangent (48 B stride)
    indices: np.ndarray                       # [6F] uint32, chunk-relative
    collider_positions: Optional[np.ndarray]  # [4F, 3] float32

    @property
    def face_count(self) -> int:
        return len(self.indices) // 6

    @property
    def position(self):
        return self.vertices[:, 0:3]

    @property
    def normal(self):
        return self.vertices[:, 3:6]

    @property
    def uvw(self):
        return self.vertices[:, 6:9]

    @property
    def tangent(self):
        return self.vertices[:, 9:12]


class MeshBatch:
    """Result of one batched tsom_mesh call: per-chunk views into the device arenas."""

    def __init__(self, ctx: Context, indices: np.ndarray, views, flags: int):
        """`views`: numpy structured array laid out like tsom_mesh_view (VIEW_DTYPE), or None when the caller skipped them."""
        self.ctx, self.chunk_indices, self._views, self.flags = ctx, indices, views, flags
        if views is None:
            self.face_counts = np.zeros(0, np.uint32)
            self.first_faces = np.zeros(0, np.uint64)
        else:
            self.face_counts = views["face_count"]
            self.first_faces = views["first_face"]

    @property
    def total_faces(self) -> int:
        return int(self.face_counts.sum())

    def __len__(self):
        return len(self.chunk_indices)

    def chunk(self, i: int) -> Optional[ChunkMesh]:
        """Host copy of chunk i's mesh; None when it has no face (the reference returns a null mesh, ClientChunkEntities.cpp:161-162)."""
        f = int(self.face_counts[i])
        if f == 0:
            return None
        first = int(self.first_faces[i])
        verts = np.empty((4 * f, 12), np.float32) if self.flags & L.MESH_RENDER else None
        coll = np.empty((4 * f, 3), np.float32) if self.flags & L.MESH_COLLIDER else None
        ind = np.empty(6 * f, np.uint32)
        L.check(L.lib().tsom_mesh_copy_to_host(self.ctx._h, first, f, None if verts is None else verts.ctypes.data_as(C.c_void_p), _p(ind, C.c_uint32),
                                               None if coll is None else _p(coll, C.c_float)))
        return ChunkMesh(verts, ind, coll)

    def checksums(self) -> np.ndarray:
        """tsom_mesh_checksums: [n, 2] uint64 — a position-weighted sum over each chunk's vertex words (collider positions after a
        collider-only call) and one over its indices, computed on the device; independent of where the faces sit in the arena."""
        out = np.zeros((len(self._views), 2), np.uint64)
        v = np.ascontiguousarray(self._views)
        L.check(L.lib().tsom_mesh_checksums(self.ctx._h, C.cast(v.ctypes.data, C.POINTER(L.MeshView)), len(v), out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def aabbs(self) -> np.ndarray:
        """Nz::StaticMesh::GenerateAABB for every chunk of the batch (ClientChunkEntities.cpp:168), computed on the device from the
        arena -> [n, 6] float32: min.xyz, max.xyz of the vertex positions; zeros for a chunk without faces."""
        out = np.zeros((len(self._views), 6), np.float32)
        v = np.ascontiguousarray(self._views)
        L.check(L.lib().tsom_mesh_aabb(self.ctx._h, C.cast(v.ctypes.data, C.POINTER(L.MeshView)), len(v), _p(out, C.c_float)))
        return out


# ------------------------------------------------------------------------------------------------ ChunkContainer / Planet
class ChunkContainer:
    """ChunkContainer (ChunkContainer.hpp): a set of 32^3 chunks resident in HBM, addressed by world-axis chunk indices."""

    ChunkSize = ChunkSize

    def __init__(self, ctx: Context, capacity: int, tileSize: float = 1.0):
        self.ctx = ctx
        self.capacity = capacity
        self._tileSize = float(tileSize)
        self._h = C.c_void_p()
        L.check(L.lib().tsom_container_create(ctx._h, capacity, self._tileSize, C.byref(self._h)))

    def close(self):
        if self._h:
            if self.ctx._h:  # a container must not outlive its context; if it did, its device memory went with the context
                L.lib().tsom_container_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- index math (ChunkContainer.inl:14-56) — host integer helpers
    @staticmethod
    def GetBlockIndices(chunkIndices, indices):
        cx, cy, cz = (int(v) for v in chunkIndices)
        x, y, z = (int(v) for v in indices)
        h = ChunkSize // 2
        return (cx * ChunkSize + x - h, cy * ChunkSize + z - h, cz * ChunkSize + y - h)

    @staticmethod
    def GetChunkIndicesByBlockIndices(indices):
        def trunc_div(a, b):  # C++ integer division truncates toward zero
            q = abs(a) // b
            return -q if a < 0 else q

        adj = [int(v) + ChunkSize // 2 for v in indices]
        chunk = [trunc_div(a - ChunkSize + 1 if a < 0 else a, ChunkSize) for a in adj]
        local = [a - c * ChunkSize for a, c in zip(adj, chunk)]
        local = [l + ChunkSize if l < 0 else l for l in local]
        return tuple(chunk), (local[0], local[2], local[1])

    def GetChunkOffset(self, indices):
        return tuple(np.float32(np.float32(np.float32(i) * np.float32(ChunkSize)) * np.float32(self._tileSize)) for i in indices)

    def GetTileSize(self) -> float:
        return self._tileSize

    def GetCenter(self):
        return (0.0, 0.0, 0.0)

    def GetChunkCount(self) -> int:
        return int(L.lib().tsom_container_chunk_count(self._h))

    # ---- chunk set
    def AddChunk(self, indices, blocks: Optional[np.ndarray] = None):
        """Planet::AddChunk (Planet.cpp:28-68); `blocks` plays the initCallback."""
        idx = _idx(indices)
        if blocks is None:
            L.check(L.lib().tsom_container_add_chunks(self._h, _p(idx, C.c_int32), len(idx)))
        else:
            self.ResetChunks(idx, blocks)

    def RemoveChunk(self, indices):
        idx = _idx(indices)
        L.check(L.lib().tsom_container_remove_chunk(self._h, _p(idx, C.c_int32)))

    def ResetChunks(self, indices, blocks):
        """Chunk::Reset(F) for a batch: F copies 32768 ids per chunk from host memory (or a device pointer given as int)."""
        idx = _idx(indices)
        if isinstance(blocks, int):
            L.check(L.lib().tsom_chunks_reset_device(self._h, _p(idx, C.c_int32), len(idx), C.c_void_p(blocks)))
            return
        b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(len(idx), ChunkBlockCount)
        L.check(L.lib().tsom_chunks_reset(self._h, _p(idx, C.c_int32), len(idx), _p(b, C.c_uint16)))

    def GetContent(self, indices) -> np.ndarray:
        """Chunk::GetContent for a batch -> [n, 32768] uint16."""
        idx = _idx(indices)
        out = np.empty((len(idx), ChunkBlockCount), np.uint16)
        L.check(L.lib().tsom_chunks_get_content(self._h, _p(idx, C.c_int32), len(idx), _p(out, C.c_uint16)))
        return out

    # ---- save format (Chunk::Serialize / Chunk::Deserialize, Chunk.cpp:252-346)
    ChunkBinaryVersion = 1  # include/CommonLib/InternalConstants.hpp:24

    def SerializeChunks(self, indices, blockLibrary: "BlockLibrary") -> list:
        """Chunk::Serialize for a batch -> one bytes object per chunk, byte for byte what the reference writes to `<x>_<y>_<z>.chunk`.
        The device produces the palette and the per-block palette indices; the header (version, size, block NAMES) is composed here.
        Stream conventions (Nazara ByteStream, restated in oracle/ref_shim): big-endian integers, string = UInt32 length + bytes."""
        import struct

        idx = _idx(indices)
        n = len(idx)
        stride = len(blockLibrary)
        pal = np.zeros((n, stride), np.uint16)
        cnt = np.zeros(n, np.uint32)
        payload = np.zeros((n, 2 * ChunkBlockCount), np.uint8)
        L.check(L.lib().tsom_chunks_palettize(self._h, _p(idx, C.c_int32), n, _p(pal, C.c_uint16), stride, _p(cnt, C.c_uint32), _p(payload, C.c_uint8), 1))
        out = []
        for i in range(n):
            k = int(cnt[i])
            head = struct.pack(">IIIIH", self.ChunkBinaryVersion, ChunkSize, ChunkSize, ChunkSize, k)
            for b in pal[i, :k]:
                name = blockLibrary.GetBlockData(int(b)).name.encode()
                head += struct.pack(">I", len(name)) + name
            out.append(head + payload[i, : ChunkBlockCount * (2 if k > 8 else 1)].tobytes())
        return out

    def DeserializeChunks(self, indices, blobs, blockLibrary: "BlockLibrary"):
        """Chunk::Deserialize for a batch: parse the headers (same exceptions as the reference), expand the payloads on the device."""
        import struct

        idx = _idx(indices)
        n = len(idx)
        assert n == len(blobs)
        stride = max(len(blockLibrary), 1)
        pal = np.zeros((n, stride), np.uint16)
        cnt = np.zeros(n, np.uint32)
        payload = np.zeros((n, 2 * ChunkBlockCount), np.uint8)
        for i, blob in enumerate(blobs):
            if len(blob) < 18:
                raise RuntimeError("ByteStream: read past end")
            version, sx, sy, sz, k = struct.unpack_from(">IIIIH", blob, 0)
            if version != self.ChunkBinaryVersion:
                raise RuntimeError("incompatible chunk version")
            if (sx, sy, sz) != (ChunkSize,) * 3:
                raise RuntimeError("incompatible chunk size")
            cur = 18
            if k > stride:
                raise RuntimeError("more block types than the library holds")
            for j in range(k):
                (ln,) = struct.unpack_from(">I", blob, cur)
                name = blob[cur + 4: cur + 4 + ln].decode(errors="replace")
                cur += 4 + ln
                b = blockLibrary.GetBlockIndex(name)
                if b == InvalidBlockIndex:
                    raise RuntimeError("unknown block " + name)
                pal[i, j] = b
            need = ChunkBlockCount * (2 if k > 8 else 1)
            if len(blob) - cur < need:
                raise RuntimeError("ByteStream: read past end")
            payload[i, :need] = np.frombuffer(blob, np.uint8, need, cur)
            cnt[i] = k
        L.check(L.lib().tsom_chunks_reset_palettized(self._h, _p(idx, C.c_int32), n, _p(pal, C.c_uint16), stride, _p(cnt, C.c_uint32), _p(payload, C.c_uint8), 1))

    # ---- wire format (Packets::ChunkReset)
    LZ4_BOUND = ChunkBlockCount * 2 + ChunkBlockCount * 2 // 255 + 16  # LZ4_compressBound(65536)

    def CompressChunks(self, indices) -> list:
        """The `content` payload of Packets::ChunkReset (Packets.cpp:172-196) for a batch: one LZ4 block per chunk, byte for byte what
        BinaryCompressor::Compress (= LZ4_compress_fast_extState, acceleration 1) makes of the chunk's block array; built on the device."""
        idx = _idx(indices)
        n = len(idx)
        offs = np.zeros(n + 1, np.uint64)
        buf = np.zeros(max(n * 16384, self.LZ4_BOUND), np.uint8)  # planet chunks compress to <= 14 KiB; retried at the exact size if not
        rc = L.lib().tsom_chunks_lz4_compress(self._h, _p(idx, C.c_int32), n, _p(buf, C.c_uint8), len(buf), _p(offs, C.c_uint64))
        if rc == L.ERR_CAPACITY:
            buf = np.zeros(int(offs[n]), np.uint8)
            rc = L.lib().tsom_chunks_lz4_compress(self._h, _p(idx, C.c_int32), n, _p(buf, C.c_uint8), len(buf), _p(offs, C.c_uint64))
        L.check(rc)
        return [buf[int(offs[i]):int(offs[i + 1])].tobytes() for i in range(n)]

    def ResetChunksCompressed(self, indices, streams):
        """The receiving side of Packets::ChunkReset (Packets.cpp:197-211, ClientSessionHandler.cpp:156-176): LZ4_decompress_safe on the
        device + Chunk::Reset; raises (and resets nothing) if any block is malformed or is not 65,536 bytes long."""
        idx = _idx(indices)
        n = len(idx)
        assert n == len(streams)
        offs = np.zeros(n + 1, np.uint64)
        offs[1:] = np.cumsum([len(b) for b in streams], dtype=np.uint64)
        buf = np.frombuffer(b"".join(streams) or b"\0", np.uint8).copy()
        L.check(L.lib().tsom_chunks_reset_lz4(self._h, _p(idx, C.c_int32), n, _p(buf, C.c_uint8), _p(offs, C.c_uint64)))

    # ---- edits
    def UpdateBlocks(self, edits: np.ndarray):
        """n x Chunk::UpdateBlock, in order; `edits` is [n, 7] int: cx, cy, cz, x, y, z, block."""
        if isinstance(edits, np.ndarray) and edits.dtype == EDIT_DTYPE:
            raw = np.ascontiguousarray(edits)  # already in the wire form of tsom_edit (see pack_edits)
        else:
            raw = pack_edits(edits)
        L.check(L.lib().tsom_apply_edits(self._h, C.cast(raw.ctypes.data, C.POINTER(L.Edit)), len(raw)))

    def UpdateBlock(self, chunkIndices, indices, newBlock: int):
        self.UpdateBlocks(np.array([[*chunkIndices, *indices, newBlock]]))

    def CollectInvalidated(self, clear: bool = True) -> np.ndarray:
        """The chunks ChunkEntities::Update would rebuild this tick, sorted by (x, y, z) -> [n, 3] int32."""
        n = C.c_uint32()
        cap = max(self.GetChunkCount(), 1)
        out = np.zeros((cap, 3), np.int32)
        L.check(L.lib().tsom_collect_dirty(self._h, _p(out, C.c_int32), cap, C.byref(n), int(clear)))
        return out[: n.value].copy()

    # ---- meshing
    def BuildMeshes(self, indices=None, render: bool = True, collider: bool = False, deformed: bool = False, deformRadius: float = 0.0,
                    gravityCenters: Optional[np.ndarray] = None, views: bool = True, genericMath: bool = False, ordered: bool = False, tangents: bool = False) -> MeshBatch:
        """Chunk::BuildMesh for a whole batch in one call (one mesh_fused_kernel launch).  `ordered`: arena ranges in job order;
        `tangents`: also what Nz::StaticMesh::GenerateTangents adds afterwards (ClientChunkEntities.cpp:169)."""
        flags = _mesh_flags(render, collider, deformed, genericMath, ordered, tangents)
        params = L.MeshParams(flags, float(deformRadius), None)
        g = None
        if gravityCenters is not None:
            g = np.ascontiguousarray(gravityCenters, dtype=np.float32).reshape(-1, 3)
            params.gravity_centers = _p(g, C.c_float)
        if indices is None:
            raise ValueError("indices is required (use list_chunks() for all)")
        idx = _idx(indices)
        v = np.zeros(len(idx), VIEW_DTYPE) if views else None
        L.check(L.lib().tsom_mesh(self._h, _p(idx, C.c_int32), len(idx), C.byref(params), None if v is None else C.cast(v.ctypes.data, C.POINTER(L.MeshView))))
        return MeshBatch(self.ctx, idx, v, flags)

    def ComputeVoxelCorners(self, chunkIndices, blockIndices, deformed: bool = False, deformRadius: float = 0.0, deformCenters: Optional[np.ndarray] = None) -> np.ndarray:
        """Chunk::ComputeVoxelCorners (Chunk.hpp:54) for n blocks: FlatChunk's corner provider, or DeformedChunk's with `deformed`
        (deformation centre per block in `deformCenters`, default container centre - chunk offset) -> [n, 8, 3] float32 in
        Nz::BoxCorner enum order.  What the game's block picking calls (GameState.cpp:859, PlayerSessionHandler.cpp:367,489)."""
        idx = _idx(chunkIndices)
        loc = np.ascontiguousarray(blockIndices, dtype=np.uint32).reshape(-1, 3)
        assert len(idx) == len(loc)
        params = L.MeshParams(L.MESH_DEFORMED if deformed else 0, float(deformRadius), None)
        g = None
        if deformCenters is not None:
            g = np.ascontiguousarray(deformCenters, dtype=np.float32).reshape(-1, 3)
            assert len(g) == len(idx)
            params.gravity_centers = _p(g, C.c_float)
        out = np.zeros((len(idx), 8, 3), np.float32)
        L.check(L.lib().tsom_voxel_corners(self._h, _p(idx, C.c_int32), _p(loc, C.c_uint32), len(idx), C.byref(params), _p(out, C.c_float)))
        return out

    def BuildBoxCollider(self, indices) -> np.ndarray:
        """FlatChunk::BuildCollider boxes -> [n, 6] float32 (pos.xyz, size.xyz) in block units."""
        idx = _idx(indices)
        out = np.zeros((16384, 6), np.float32)
        n = C.c_uint32()
        L.check(L.lib().tsom_box_collider(self._h, _p(idx, C.c_int32), _p(out, C.c_float), len(out), C.byref(n)))
        return out[: n.value].copy()

    def BuildBoxColliders(self, indices, copy: bool = True):
        """FlatChunk::BuildCollider for a batch of chunks in ONE call (what ChunkEntities::UpdateChunkEntity asks per invalidated chunk,
        ChunkEntities.cpp:198) -> (first_box [n] uint64, box_count [n] uint32, boxes [total, 6] float32 or None when copy is False)."""
        idx = _idx(indices)
        views = np.zeros(len(idx), BOX_VIEW_DTYPE)
        total = C.c_uint64()
        L.check(L.lib().tsom_box_colliders(self._h, _p(idx, C.c_int32), len(idx), C.cast(views.ctypes.data, C.POINTER(L.BoxView)), C.byref(total)))
        boxes = None
        if copy:
            boxes = np.zeros((int(total.value), 6), np.float32)
            if total.value:
                L.check(L.lib().tsom_box_colliders_copy_to_host(self.ctx._h, 0, total.value, _p(boxes, C.c_float)))
        return views["first_box"], views["box_count"], boxes

    # ---- multi-GPU exchange (see tsom_halo_publish / tsom_halo_pull in include/tsom_b200.h)
    def PublishHalo(self):
        L.check(L.lib().tsom_halo_publish(self._h))

    def PullHalos(self):
        L.check(L.lib().tsom_halo_pull(self._h))


class Planet(ChunkContainer):
    """Planet (Planet.hpp / Planet.cpp): a ChunkContainer of FlatChunks with generation."""

    def __init__(self, ctx: Context, tileSize: float = 1.0, cornerRadius: float = 16.0, gravity: float = 9.81, capacity: int = 125):
        super().__init__(ctx, capacity, tileSize)
        self._cornerRadius, self._gravity = cornerRadius, gravity

    def GetCornerRadius(self) -> float:
        return self._cornerRadius

    def GetGravity(self) -> float:
        return self._gravity

    def _gen_ids(self, lib: BlockLibrary):
        return L.GenBlocks(*(lib.GetBlockIndex(n) for n in ("dirt", "grass", "stone", "stone_mossy", "snow")))

    def GenerateChunks(self, blockLibrary: BlockLibrary, seed: int, chunkCount: Sequence[int]):
        """Planet::GenerateChunks (Planet.cpp:385-403); the TaskScheduler argument has no counterpart (the GPU is the scheduler)."""
        cc = np.ascontiguousarray(chunkCount, dtype=np.uint32)
        ids = self._gen_ids(blockLibrary)
        L.check(L.lib().tsom_planet_generate(self._h, seed & 0xFFFFFFFF, _p(cc, C.c_uint32), C.byref(ids)))

    def GenerateChunk(self, blockLibrary: BlockLibrary, indices, seed: int, chunkCount: Sequence[int], wait: bool = True):
        """Planet::GenerateChunk (Planet.cpp:160-383) on one or more chunks.  wait=False: return once the work is queued
        (tsom_planet_generate_chunks_async); the next blocking call on the context completes it."""
        cc = np.ascontiguousarray(chunkCount, dtype=np.uint32)
        idx = _idx(indices)
        ids = self._gen_ids(blockLibrary)
        fn = L.lib().tsom_planet_generate_chunks if wait else L.lib().tsom_planet_generate_chunks_async
        L.check(fn(self._h, seed & 0xFFFFFFFF, _p(cc, C.c_uint32), C.byref(ids), _p(idx, C.c_int32), len(idx)))

    def GeneratePlatform(self, blockLibrary: BlockLibrary, upDirection: int, platformCenter: Sequence[int]):
        """Planet::GeneratePlatform (Planet.cpp:405-512)."""
        ids = L.PlatformBlocks(*(blockLibrary.GetBlockIndex(n) for n in ("copper_block", "stone_bricks", "planks")))
        ctr = np.ascontiguousarray(platformCenter, dtype=np.int32)
        L.check(L.lib().tsom_planet_generate_platform(self._h, int(upDirection), _p(ctr, C.c_int32), C.byref(ids)))

    @staticmethod
    def chunk_grid(chunkCount: Sequence[int]) -> np.ndarray:
        """Chunk indices in GenerateChunks order (z, y, x loops; (x,y,z) - chunkCount/2) -> [n, 3] int32."""
        nx, ny, nz = (int(v) for v in chunkCount)
        z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        return np.stack([x.ravel() - nx // 2, y.ravel() - ny // 2, z.ravel() - nz // 2], axis=1).astype(np.int32)


class ChunkEntities:
    """Per-tick glue (ChunkEntities::Update, ClientChunkEntities::ProcessChunkUpdate): rebuild every invalidated chunk and the
    masked neighbours in ONE batched GPU call instead of one TaskScheduler task per chunk."""

    def __init__(self, chunkContainer: ChunkContainer, render: bool = True, collider: bool = True):
        self.container, self.render, self.collider = chunkContainer, render, collider

    def Update(self) -> Optional[MeshBatch]:
        dirty = self.container.CollectInvalidated(clear=True)
        if len(dirty) == 0:
            return None
        return self.container.BuildMeshes(dirty, render=self.render, collider=self.collider)


# ------------------------------------------------------------------------------------------------ Ship
class Ship(ChunkContainer):
    """Ship (include/CommonLib/Ship.hpp, src/CommonLib/Ship.cpp): the second ChunkContainer of the reference — a handful of FlatChunks
    whose block edits raise Ship::AddChunk's neighbour mask (local z == 0 -> Direction::Up, z == 31 -> Direction::Down, Ship.cpp:36-54:
    the opposite of Planet's) and whose physics body is the compound of its chunks' box colliders (Ship::BuildHullCollider)."""

    def __init__(self, ctx: Context, tileSize: float = 1.0, capacity: int = 8):
        super().__init__(ctx, capacity, tileSize)
        L.check(L.lib().tsom_container_set_kind(self._h, L.CONTAINER_SHIP))
        self._order: list = []  # chunk indices in insertion order (the reference iterates a hopscotch_map: no defined order)

    def AddChunk(self, blockLibrary: BlockLibrary, indices, blocks: Optional[np.ndarray] = None):
        """Ship::AddChunk (Ship.cpp:21-61); `blocks` plays the initCallback."""
        super().AddChunk(indices, blocks)
        key = tuple(int(v) for v in np.asarray(indices).reshape(3))
        if key not in self._order:
            self._order.append(key)

    def RemoveChunk(self, indices):
        super().RemoveChunk(indices)
        key = tuple(int(v) for v in np.asarray(indices).reshape(3))
        if key in self._order:
            self._order.remove(key)

    def Generate(self, blockLibrary: BlockLibrary, small: bool):
        """Ship::Generate (Ship.cpp:115-152): chunk (0,0,0) with the hull box and its forcefield door; nothing without a "hull" block."""
        hull = blockLibrary.GetBlockIndex("hull")
        if (0, 0, 0) not in self._order:
            super().AddChunk((0, 0, 0))
            self._order.append((0, 0, 0))
        if hull == InvalidBlockIndex:
            return
        L.check(L.lib().tsom_ship_generate(self._h, int(bool(small)), hull, blockLibrary.GetBlockIndex("forcefield")))

    def BuildHullCollider(self):
        """Ship::BuildHullCollider (Ship.cpp:63-93): every chunk's FlatChunk::BuildCollider in ONE batched call.
        -> None when no chunk has a colliding block; with one chunk its boxes [n, 6] (pos.xyz, size.xyz in block units, as
        FlatChunk::BuildCollider's compound lists them); with several a list of (GetChunkOffset(indices), boxes) children."""
        if not self._order:
            return None
        idx = np.array(self._order, np.int32)
        first, count, boxes = self.BuildBoxColliders(idx)
        per_chunk = [boxes[int(f):int(f) + int(c)] for f, c in zip(first, count)]
        if len(self._order) > 1:
            children = [(self.GetChunkOffset(k), b) for k, b in zip(self._order, per_chunk) if len(b)]
            return children or None
        return per_chunk[0] if len(per_chunk[0]) else None


# ------------------------------------------------------------------------------------------------ one planet over several GPUs, one process
class MultiGpuPlanet:
    """tsom_planet_create_sharded: a Planet cut into one box of chunks per GPU, all GPUs driven from THIS process (a TSOM server is one
    process, src/Server/main.cpp:24-62).  Same call surface as Planet for the whole-planet operations; meshes come back per chunk
    with the part (GPU) whose arenas hold them.  `devices` may name a GPU twice (two parts on one GPU)."""

    def __init__(self, devices: Sequence[int], chunkCount: Sequence[int], tileSize: float = 1.0, blockLibrary: Optional[BlockLibrary] = None, weights: Optional[np.ndarray] = None):
        self.chunkCount = tuple(int(v) for v in chunkCount)
        self.blockLibrary = blockLibrary or BlockLibrary()
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        cc = np.ascontiguousarray(self.chunkCount, dtype=np.uint32)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
        self._h = C.c_void_p()
        L.check(L.lib().tsom_planet_create_sharded(dev, len(devices), _p(cc, C.c_uint32), float(tileSize), None if w is None else _p(w, C.c_float), C.byref(self._h)))
        descs = self.blockLibrary.to_descs()
        L.check(L.lib().tsom_sharded_set_block_table(self._h, descs, len(descs)))
        self.parts = int(L.lib().tsom_sharded_part_count(self._h))
        self.grid = Planet.chunk_grid(self.chunkCount)

    def close(self):
        if self._h:
            L.lib().tsom_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def part_box(self, part: int) -> np.ndarray:
        b = np.zeros(6, np.int32)
        L.check(L.lib().tsom_sharded_part_box(self._h, part, _p(b, C.c_int32)))
        return b

    def part_context(self, part: int) -> Context:
        return Context(blockLibrary=self.blockLibrary, _borrowed=L.lib().tsom_sharded_ctx(self._h, part))

    def owner(self, indices) -> np.ndarray:
        idx = _idx(indices)
        out = np.zeros(len(idx), np.uint32)
        o = C.c_uint32()
        for i in range(len(idx)):
            L.check(L.lib().tsom_sharded_owner(self._h, _p(idx[i:i + 1], C.c_int32), C.byref(o)))
            out[i] = o.value
        return out

    def GenerateChunks(self, seed: int):
        """Planet::GenerateChunks (Planet.cpp:385-403) over all GPUs + one halo exchange round."""
        lib = self.blockLibrary
        ids = L.GenBlocks(*(lib.GetBlockIndex(n) for n in ("dirt", "grass", "stone", "stone_mossy", "snow")))
        L.check(L.lib().tsom_sharded_generate(self._h, seed & 0xFFFFFFFF, C.byref(ids)))

    def UpdateBlocks(self, edits):
        raw = np.ascontiguousarray(edits) if isinstance(edits, np.ndarray) and edits.dtype == EDIT_DTYPE else pack_edits(edits)
        L.check(L.lib().tsom_sharded_apply_edits(self._h, C.cast(raw.ctypes.data, C.POINTER(L.Edit)), len(raw)))

    def CollectInvalidated(self, clear: bool = True) -> np.ndarray:
        n = C.c_uint32()
        out = np.zeros((len(self.grid), 3), np.int32)
        L.check(L.lib().tsom_sharded_collect_dirty(self._h, _p(out, C.c_int32), len(out), C.byref(n), int(clear)))
        return out[: n.value].copy()

    def GetContent(self, indices) -> np.ndarray:
        idx = _idx(indices)
        out = np.empty((len(idx), ChunkBlockCount), np.uint16)
        L.check(L.lib().tsom_sharded_get_content(self._h, _p(idx, C.c_int32), len(idx), _p(out, C.c_uint16)))
        return out

    def BuildMeshes(self, indices=None, render: bool = True, collider: bool = False, deformed: bool = False, deformRadius: float = 0.0, views: bool = True,
                    genericMath: bool = False, tangents: bool = False):
        """Chunk::BuildMesh for the listed chunks (None: the whole planet in GenerateChunks order), each on the GPU that owns it.
        -> (views [n] VIEW_DTYPE, parts [n] uint32, flags); a chunk's faces are read with `chunk_mesh`."""
        flags = _mesh_flags(render, collider, deformed, genericMath, False, tangents)
        params = L.MeshParams(flags, float(deformRadius), None)
        idx = None if indices is None else _idx(indices)
        n = len(self.grid) if idx is None else len(idx)
        v = np.zeros(n, VIEW_DTYPE) if views else None
        parts = np.zeros(n, np.uint32) if views else None
        L.check(L.lib().tsom_sharded_mesh(self._h, None if idx is None else _p(idx, C.c_int32), n, C.byref(params), None if v is None else C.cast(v.ctypes.data, C.POINTER(L.MeshView)),
                                          None if parts is None else _p(parts, C.c_uint32)))
        return v, parts, flags

    def chunk_mesh(self, view, part: int, flags: int) -> Optional[ChunkMesh]:
        f = int(view["face_count"])
        if f == 0:
            return None
        verts = np.empty((4 * f, 12), np.float32) if flags & L.MESH_RENDER else None
        coll = np.empty((4 * f, 3), np.float32) if flags & L.MESH_COLLIDER else None
        ind = np.empty(6 * f, np.uint32)
        L.check(L.lib().tsom_sharded_mesh_copy_to_host(self._h, int(part), int(view["first_face"]), f, None if verts is None else verts.ctypes.data_as(C.c_void_p), _p(ind, C.c_uint32),
                                                       None if coll is None else _p(coll, C.c_float)))
        return ChunkMesh(verts, ind, coll)

    def timings(self) -> list:
        t = np.zeros((self.parts, L.TIMING_COUNT), np.float32)
        L.check(L.lib().tsom_sharded_get_timings(self._h, _p(t, C.c_float)))
        return [_timings_dict(row) for row in t]
