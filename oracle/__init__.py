"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front-end of ``libtsom_oracle.so`` (``oracle/capi.cpp`` over ``oracle/tsom_oracle.hpp``), the CPU restatement of
ThisSpaceOfMine's chunk generation / meshing / collider / edit rules.

Two backends share this front-end (same C entry points, ``orc_*``):

* the default one, ``libtsom_oracle.so`` — the restatement ("port"), travels everywhere;
* ``reference()`` — ``oracle/_ref/libtsom_ref.so``: the reference's OWN ``Chunk.cpp / FlatChunk.cpp / DeformedChunk.cpp / Planet.cpp /
  BlockLibrary.cpp`` compiled unmodified from ``/root/reference`` against the stand-in third-party headers of ``oracle/ref_shim``
  (built by ``oracle/ref_shim/Makefile`` wherever ``/root/reference`` exists; the prebuilt ``.so`` travels to the GPU box).

Parity status: the restatement's structure is pinned against the reference's real code by ``tests/test_reference_vs_oracle.py``;
what stays *unpinned* is the third-party arithmetic that both the shim and the oracle restate (Nazara ``Box`` corner naming,
``Quaternion::RotationBetween``, ``siv::PerlinNoise``), for which the reference ships no test.  The reference's only KATs
(``tests/UnitTests/Common/ChunkTests.cpp:17-43``) are checked in ``tests/test_oracle_kat.py`` against both backends.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtsom_oracle.so")

CHUNK = 32
NBLK = CHUNK ** 3

# Direction.hpp:18-28
BACK, DOWN, FRONT, LEFT, RIGHT, UP = range(6)
CHUNK_DIR_OFFSET = np.array([[0, 0, 1], [0, -1, 0], [0, 0, -1], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.int32)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "tsom_oracle.hpp")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libtsom_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_REF_SO = os.path.join(_HERE, "_ref", "libtsom_ref.so")
_REF_DIR = os.path.join(_HERE, "ref_shim")


def build_reference() -> str | None:
    """(Re)build oracle/_ref/libtsom_ref.so from the reference's sources when /root/reference is present; else keep the prebuilt one."""
    subprocess.check_call(["make", "-C", _REF_DIR], stdout=subprocess.DEVNULL)
    return _REF_SO if os.path.exists(_REF_SO) else None


def _bind(path: str, full: bool):
    L = C.CDLL(path)
    i32p, u32p, u16p, f32p, u8p = (C.POINTER(t) for t in (C.c_int32, C.c_uint32, C.c_uint16, C.c_float, C.c_uint8))
    L.orc_get_block_indices.argtypes = [i32p, u32p, i32p]
    L.orc_get_chunk_indices_by_block_indices.argtypes = [i32p, i32p, u32p]
    L.orc_direction_from_normal.argtypes = [f32p]
    L.orc_blocklib_count.restype = C.c_uint32
    L.orc_blocklib_get.argtypes = [C.c_uint32, u32p, u8p, C.c_char_p]
    L.orc_blocklib_index.argtypes = [C.c_char_p]
    L.orc_blocklib_register.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    if full:  # probes of the restatement's internals; the reference build has no such entry points
        L.orc_lz4_bound.argtypes = [C.c_int]
        L.orc_lz4_compress.argtypes = [u8p, C.c_int, u8p, C.c_int]
        L.orc_lz4_decompress.argtypes = [u8p, C.c_int, u8p, C.c_int]
        L.orc_perlin_perm.argtypes = [C.c_uint32, u8p]
        L.orc_perlin_noise3d_perm.argtypes = [u8p, C.c_double, C.c_double, C.c_double]
        L.orc_perlin_noise3d_perm.restype = C.c_double
        L.orc_perlin_octave01.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_int]
        L.orc_perlin_octave01.restype = C.c_double
        L.orc_bernoulli_draws.argtypes = [C.c_uint32, C.c_uint32, u8p]
    L.orc_generate_chunk.argtypes = [C.c_uint32, u32p, i32p, u16p]
    L.orc_planet_create.argtypes = [C.c_float, C.c_float]
    L.orc_planet_create.restype = C.c_void_p
    L.orc_planet_destroy.argtypes = [C.c_void_p]
    L.orc_planet_add_chunk.argtypes = [C.c_void_p, i32p, u16p]
    L.orc_planet_reset_chunk.argtypes = [C.c_void_p, i32p, u16p]
    L.orc_planet_generate.argtypes = [C.c_void_p, C.c_uint32, u32p, C.c_int]
    L.orc_planet_generate_platform.argtypes = [C.c_void_p, C.c_int, i32p]
    L.orc_planet_get_blocks.argtypes = [C.c_void_p, i32p, u16p]
    L.orc_planet_update_block.argtypes = [C.c_void_p, i32p, u32p, C.c_uint16]
    L.orc_planet_apply_edits.argtypes = [C.c_void_p, i32p, C.c_uint32]
    L.orc_planet_clear_invalidated.argtypes = [C.c_void_p]
    L.orc_planet_collect_dirty.argtypes = [C.c_void_p, i32p, C.c_uint32]
    L.orc_planet_collect_dirty.restype = C.c_uint32
    L.orc_mesh_build.argtypes = [C.c_void_p, i32p, C.c_int, f32p, f32p, C.c_float, C.c_int, C.c_int]
    L.orc_mesh_build.restype = C.c_void_p
    L.orc_voxel_corners.argtypes = [C.c_void_p, i32p, u32p, C.c_int, f32p, C.c_float, f32p]
    L.orc_mesh_face_count.argtypes = [C.c_void_p]
    L.orc_mesh_face_count.restype = C.c_uint32
    L.orc_mesh_copy.argtypes = [C.c_void_p, f32p, f32p, f32p, u32p]
    L.orc_mesh_free.argtypes = [C.c_void_p]
    L.orc_box_collider.argtypes = [C.c_void_p, i32p, f32p, C.c_uint32]
    L.orc_box_collider.restype = C.c_uint32
    L.orc_chunk_serialize.argtypes = [C.c_void_p, i32p, u8p, C.c_uint32]
    L.orc_chunk_serialize.restype = C.c_uint32
    L.orc_chunk_deserialize.argtypes = [C.c_void_p, i32p, u8p, C.c_uint32, C.c_char_p]
    L.orc_bench_mesh_isolated.argtypes = [u16p, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    L.orc_bench_mesh_isolated.restype = C.c_double
    L.orc_bench_planet.argtypes = [C.c_uint32, u32p, C.c_int, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint64), u32p]
    L.orc_generate_tangents.argtypes = [f32p, f32p, u32p, C.c_uint32, C.c_uint32, f32p]
    L.orc_sample_create.argtypes = [C.c_uint32, u32p, C.c_uint32, C.c_uint32, C.c_int]
    L.orc_sample_create.restype = C.c_void_p
    L.orc_sample_size.argtypes = [C.c_void_p]
    L.orc_sample_size.restype = C.c_uint32
    L.orc_sample_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.orc_sample_step.restype = C.c_uint64
    L.orc_sample_destroy.argtypes = [C.c_void_p]
    L.orc_bench_box_colliders.argtypes = [C.c_void_p, u32p, C.c_uint32, C.c_int, C.POINTER(C.c_uint64), u32p]
    L.orc_bench_box_colliders.restype = C.c_double
    L.orc_ship_create.argtypes = [C.c_float]
    L.orc_ship_create.restype = C.c_void_p
    L.orc_ship_destroy.argtypes = [C.c_void_p]
    L.orc_ship_generate.argtypes = [C.c_void_p, C.c_int]
    L.orc_ship_add_chunk.argtypes = [C.c_void_p, i32p, u16p]
    L.orc_ship_get_blocks.argtypes = [C.c_void_p, i32p, u16p]
    L.orc_ship_update_block.argtypes = [C.c_void_p, i32p, u32p, C.c_uint16]
    L.orc_ship_take_invalidated.argtypes = [C.c_void_p, i32p, u32p, C.c_uint32]
    L.orc_ship_take_invalidated.restype = C.c_uint32
    L.orc_ship_hull_collider.argtypes = [C.c_void_p, f32p, C.c_uint32]
    L.orc_ship_hull_collider.restype = C.c_uint32
    return L


_lib = None
_ref_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = _bind(_SO, full=True)
    return _lib


def reference_available() -> bool:
    """True when oracle/_ref/libtsom_ref.so exists (prebuilt) or can be built here (/root/reference present)."""
    if os.path.exists(_REF_SO):
        return True
    try:
        return build_reference() is not None
    except Exception:
        return False


def ref_lib():
    global _ref_lib
    if _ref_lib is None:
        if os.path.isdir("/root/reference"):
            build_reference()  # make: no-op when up to date
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libtsom_ref.so is not built (needs /root/reference; see oracle/ref_shim/Makefile)")
        _ref_lib = _bind(_REF_SO, full=False)
    return _ref_lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _i3(v):
    return np.ascontiguousarray(v, dtype=np.int32)


def _u3(v):
    return np.ascontiguousarray(v, dtype=np.uint32)


# ---------------------------------------------------------------- index math / tables
def get_block_indices(chunk, local, L=None):
    out = np.zeros(3, np.int32)
    (L or lib()).orc_get_block_indices(_p(_i3(chunk), C.c_int32), _p(_u3(local), C.c_uint32), _p(out, C.c_int32))
    return tuple(int(v) for v in out)


def get_chunk_indices_by_block_indices(block, L=None):
    c = np.zeros(3, np.int32)
    l = np.zeros(3, np.uint32)
    (L or lib()).orc_get_chunk_indices_by_block_indices(_p(_i3(block), C.c_int32), _p(c, C.c_int32), _p(l, C.c_uint32))
    return tuple(int(v) for v in c), tuple(int(v) for v in l)


def direction_from_normal(n, L=None):
    a = np.ascontiguousarray(n, dtype=np.float32)
    return int((L or lib()).orc_direction_from_normal(_p(a, C.c_float)))


def block_library(L=None):
    """[(name, tex[6], hasCollisions, isDoubleSided, isTransparent)] in BlockIndex order (BlockLibrary.cpp:10-83)."""
    res = []
    for i in range((L or lib()).orc_blocklib_count()):
        tex = np.zeros(6, np.uint32)
        fl = np.zeros(3, np.uint8)
        name = C.create_string_buffer(64)
        (L or lib()).orc_blocklib_get(i, _p(tex, C.c_uint32), _p(fl, C.c_uint8), name)
        res.append((name.value.decode(), tex.copy(), bool(fl[0]), bool(fl[1]), bool(fl[2])))
    return res


def register_block(name, basePath="", sidePath="", upPath="", hasCollisions=True, isDoubleSided=False, isTransparent=False, L=None) -> int:
    """BlockLibrary::RegisterBlock on the backend's process-wide library (appends; use a fresh process for such tests)."""
    return int((L or lib()).orc_blocklib_register(name.encode(), basePath.encode(), sidePath.encode(), upPath.encode(), int(hasCollisions), int(isDoubleSided), int(isTransparent)))


def block_index(name: str, L=None) -> int:
    return int((L or lib()).orc_blocklib_index(name.encode()))


def perlin_perm(seed: int, L=None) -> np.ndarray:
    out = np.zeros(256, np.uint8)
    (L or lib()).orc_perlin_perm(seed & 0xFFFFFFFF, _p(out, C.c_uint8))
    return out


def perlin_octave01(seed: int, x: float, y: float, octaves: int = 4, L=None) -> float:
    return float((L or lib()).orc_perlin_octave01(seed & 0xFFFFFFFF, x, y, octaves))


def bernoulli_draws(seed: int, n: int, L=None) -> np.ndarray:
    out = np.zeros(n, np.uint8)
    (L or lib()).orc_bernoulli_draws(seed & 0xFFFFFFFF, n, _p(out, C.c_uint8))
    return out


def generate_chunk(seed, chunk_count, chunk_idx, L=None):
    """Planet::GenerateChunk for one chunk -> (blocks u16[32768], ok)."""
    out = np.zeros(NBLK, np.uint16)
    rc = (L or lib()).orc_generate_chunk(seed & 0xFFFFFFFF, _p(_u3(chunk_count), C.c_uint32), _p(_i3(chunk_idx), C.c_int32), _p(out, C.c_uint16))
    return out, rc == 0


class Mesh:
    """One chunk's mesh in the reference's emission order."""

    def __init__(self, pos, normal, uvw, indices):
        self.position, self.normal, self.uvw, self.indices = pos, normal, uvw, indices

    @property
    def face_count(self):
        return len(self.indices) // 6

    def interleaved48(self) -> np.ndarray:
        """VertexStruct{pos,normal,uvw,tangent} stride 48 B (ClientChunkEntities.hpp:27-33); tangent left zero (never written by Chunk::BuildMesh)."""
        n = len(self.position)
        v = np.zeros((n, 12), np.float32)
        v[:, 0:3] = self.position
        if self.normal is not None:
            v[:, 3:6] = self.normal
        if self.uvw is not None:
            v[:, 6:9] = self.uvw
        return v


class Planet:
    """Handle on an oracle Planet (ChunkContainer of FlatChunks, Planet.cpp)."""

    def __init__(self, tile: float = 1.0, corner_radius: float = 16.0, L=None):
        self._L = L or lib()
        self._h = C.c_void_p(self._L.orc_planet_create(tile, corner_radius))
        self.tile = tile

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_planet_destroy(self._h)
            self._h = None

    def add_chunk(self, idx, blocks=None):
        b = None if blocks is None else _p(np.ascontiguousarray(blocks, dtype=np.uint16), C.c_uint16)
        return self._L.orc_planet_add_chunk(self._h, _p(_i3(idx), C.c_int32), b)

    def reset_chunk(self, idx, blocks):
        return self._L.orc_planet_reset_chunk(self._h, _p(_i3(idx), C.c_int32), _p(np.ascontiguousarray(blocks, dtype=np.uint16), C.c_uint16))

    def generate(self, seed, chunk_count, nthreads=1) -> bool:
        return self._L.orc_planet_generate(self._h, seed & 0xFFFFFFFF, _p(_u3(chunk_count), C.c_uint32), nthreads) == 0

    def generate_platform(self, direction, center):
        self._L.orc_planet_generate_platform(self._h, int(direction), _p(_i3(center), C.c_int32))

    def get_blocks(self, idx):
        out = np.zeros(NBLK, np.uint16)
        rc = self._L.orc_planet_get_blocks(self._h, _p(_i3(idx), C.c_int32), _p(out, C.c_uint16))
        return out if rc == 0 else None

    def update_block(self, idx, local, block_id):
        return self._L.orc_planet_update_block(self._h, _p(_i3(idx), C.c_int32), _p(_u3(local), C.c_uint32), int(block_id))

    def apply_edits(self, edits):
        e = np.ascontiguousarray(edits, dtype=np.int32).reshape(-1, 7)
        self._L.orc_planet_apply_edits(self._h, _p(e, C.c_int32), len(e))

    def clear_invalidated(self):
        self._L.orc_planet_clear_invalidated(self._h)

    def collect_dirty(self, cap=1 << 20):
        out = np.zeros((cap, 3), np.int32)
        n = self._L.orc_planet_collect_dirty(self._h, _p(out, C.c_int32), cap)
        return out[:n].copy()

    def mesh(self, idx, deformed=False, gravity_center=None, deform_center=None, deform_radius=0.0, normal=True, uv=True):
        g = None if gravity_center is None else _p(np.ascontiguousarray(gravity_center, dtype=np.float32), C.c_float)
        d = None if deform_center is None else _p(np.ascontiguousarray(deform_center, dtype=np.float32), C.c_float)
        h = self._L.orc_mesh_build(self._h, _p(_i3(idx), C.c_int32), int(deformed), g, d, float(deform_radius), int(normal), int(uv))
        if not h:
            return None
        h = C.c_void_p(h)
        f = self._L.orc_mesh_face_count(h)
        pos = np.zeros((4 * f, 3), np.float32)
        nor = np.zeros((4 * f, 3), np.float32) if normal else None
        uvw = np.zeros((4 * f, 3), np.float32) if uv else None
        ind = np.zeros(6 * f, np.uint32)
        self._L.orc_mesh_copy(h, _p(pos, C.c_float), None if nor is None else _p(nor, C.c_float), None if uvw is None else _p(uvw, C.c_float), _p(ind, C.c_uint32))
        self._L.orc_mesh_free(h)
        return Mesh(pos, nor, uvw, ind)

    def serialize(self, idx) -> bytes:
        """Chunk::Serialize (Chunk.cpp:312-346): the bytes of the chunk's save file."""
        buf = np.zeros(1 << 17, np.uint8)
        n = self._L.orc_chunk_serialize(self._h, _p(_i3(idx), C.c_int32), _p(buf, C.c_uint8), len(buf))
        return bytes(buf[:n])

    def deserialize(self, idx, data: bytes):
        """Chunk::Deserialize (Chunk.cpp:252-310) into an existing chunk; raises RuntimeError with the reference's message."""
        a = np.frombuffer(data, np.uint8).copy()
        err = C.create_string_buffer(128)
        rc = self._L.orc_chunk_deserialize(self._h, _p(_i3(idx), C.c_int32), _p(a, C.c_uint8), len(a), err)
        if rc == 1:
            raise KeyError(tuple(idx))
        if rc == 2:
            raise RuntimeError(err.value.decode())

    def voxel_corners(self, idx, local, deformed=False, deform_center=None, deform_radius=0.0) -> np.ndarray:
        """Chunk::ComputeVoxelCorners (Chunk.hpp:54) of block `local` of chunk `idx` -> [8, 3] float32 in Nz::BoxCorner enum order."""
        out = np.zeros((8, 3), np.float32)
        d = None if deform_center is None else _p(np.ascontiguousarray(deform_center, dtype=np.float32), C.c_float)
        rc = self._L.orc_voxel_corners(self._h, _p(_i3(idx), C.c_int32), _p(_u3(local), C.c_uint32), int(deformed), d, float(deform_radius), _p(out, C.c_float))
        if rc:
            raise KeyError(tuple(idx))
        return out

    def box_collider(self, idx, cap=32768):
        out = np.zeros((cap, 6), np.float32)
        n = self._L.orc_box_collider(self._h, _p(_i3(idx), C.c_int32), _p(out, C.c_float), cap)
        return out[:n].copy()


class Ship:
    """Handle on an oracle Ship (ChunkContainer of FlatChunks, Ship.cpp): the Up/Down-swapped dirty mask, Generate, hull collider."""

    def __init__(self, tile: float = 1.0, L=None):
        self._L = L or lib()
        self._h = C.c_void_p(self._L.orc_ship_create(tile))
        self.tile = tile

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_ship_destroy(self._h)
            self._h = None

    def generate(self, small: bool):
        self._L.orc_ship_generate(self._h, int(small))

    def add_chunk(self, idx, blocks=None):
        b = None if blocks is None else _p(np.ascontiguousarray(blocks, dtype=np.uint16), C.c_uint16)
        return self._L.orc_ship_add_chunk(self._h, _p(_i3(idx), C.c_int32), b)

    def get_blocks(self, idx):
        out = np.zeros(NBLK, np.uint16)
        rc = self._L.orc_ship_get_blocks(self._h, _p(_i3(idx), C.c_int32), _p(out, C.c_uint16))
        return out if rc == 0 else None

    def update_block(self, idx, local, block_id):
        return self._L.orc_ship_update_block(self._h, _p(_i3(idx), C.c_int32), _p(_u3(local), C.c_uint32), int(block_id))

    def take_invalidated(self, cap=4096):
        """{chunk index: direction mask | 0x80} raised since the last call (what ChunkEntities collects from OnChunkUpdated); cleared."""
        idx = np.zeros((cap, 3), np.int32)
        mask = np.zeros(cap, np.uint32)
        n = self._L.orc_ship_take_invalidated(self._h, _p(idx, C.c_int32), _p(mask, C.c_uint32), cap)
        return {tuple(int(v) for v in idx[i]): int(mask[i]) for i in range(n)}

    def hull_collider(self, cap=1 << 16) -> np.ndarray:
        """Ship::BuildHullCollider flattened -> [n, 9] float32: child offset, box centre, box lengths (world units); n == 0: null."""
        out = np.zeros((cap, 9), np.float32)
        n = self._L.orc_ship_hull_collider(self._h, _p(out, C.c_float), cap)
        return out[:n].copy()


class PlanetSample:
    """A bounded, planet-wide sample of "generate + mesh" (every stride-th chunk, neighbours generated once at set-up): the workload of
    bench.py's CPU arm.  step() regenerates and meshes the sample, one task per chunk on nthreads workers."""

    def __init__(self, seed, chunk_count, stride, offset=0, nthreads=1, L=None):
        self._L = L or lib()
        self._h = C.c_void_p(self._L.orc_sample_create(seed & 0xFFFFFFFF, _p(_u3(chunk_count), C.c_uint32), int(stride), int(offset), int(nthreads)))
        if not self._h:
            raise RuntimeError("planet sample: unsupported planet")
        self.size = int(self._L.orc_sample_size(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_sample_destroy(self._h)
            self._h = None

    def step(self, nthreads):
        t = (C.c_double * 2)()
        faces = int(self._L.orc_sample_step(self._h, int(nthreads), t))
        return dict(gen_s=t[0], mesh_s=t[1], faces=faces, chunks=self.size)


def bench_box_colliders(planet: "Planet", chunk_count, nthreads, stride=1, L=None):
    """Time FlatChunk::BuildCollider over every stride-th chunk of a generated planet (the backend is the planet's) -> (seconds, boxes, chunks)."""
    boxes = C.c_uint64(0)
    chunks = C.c_uint32(0)
    t = planet._L.orc_bench_box_colliders(planet._h, _p(_u3(chunk_count), C.c_uint32), int(stride), int(nthreads), C.byref(boxes), C.byref(chunks))
    return float(t), int(boxes.value), int(chunks.value)


def generate_tangents(mesh: "Mesh", L=None) -> np.ndarray:
    """Nz::SubMesh::GenerateTangents on a chunk mesh (ClientChunkEntities.cpp:169; restated in oracle/nazara_tangents.hpp, unpinned
    third-party arithmetic) -> [4F, 3] float32."""
    pos = np.ascontiguousarray(mesh.position, np.float32)
    uvw = np.ascontiguousarray(mesh.uvw, np.float32)
    ind = np.ascontiguousarray(mesh.indices, np.uint32)
    out = np.zeros_like(pos)
    (L or lib()).orc_generate_tangents(_p(pos, C.c_float), _p(uvw, C.c_float), _p(ind, C.c_uint32), len(pos), len(ind), _p(out, C.c_float))
    return out


def perlin_noise3d_perm(perm, x: float, y: float, z: float) -> float:
    """siv::PerlinNoise::noise3D over a caller-supplied permutation table (deserialize)."""
    p = np.ascontiguousarray(perm, dtype=np.uint8)
    assert p.shape == (256,)
    return float(lib().orc_perlin_noise3d_perm(_p(p, C.c_uint8), float(x), float(y), float(z)))


def lz4_compress(data) -> bytes:
    """BinaryCompressor::Compress (BinaryCompressor.cpp:17-36) = LZ4_compress_fast_extState(acceleration 0 -> 1): restated in
    oracle/lz4_block.hpp for inputs shorter than 64 KiB + 11 (a chunk's block array is exactly 64 KiB); b"" if it does not apply."""
    a = np.frombuffer(bytes(data), np.uint8).copy() if not isinstance(data, np.ndarray) else np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    L = lib()
    out = np.zeros(max(L.orc_lz4_bound(len(a)), 1), np.uint8)
    n = L.orc_lz4_compress(_p(a, C.c_uint8), len(a), _p(out, C.c_uint8), len(out))
    return bytes(out[:n])


def lz4_decompress(data: bytes, size: int) -> bytes:
    """BinaryCompressor::Decompress (BinaryCompressor.cpp:38-47) = LZ4_decompress_safe; raises ValueError on a malformed stream."""
    a = np.frombuffer(bytes(data), np.uint8).copy()
    out = np.zeros(max(size, 1), np.uint8)
    n = lib().orc_lz4_decompress(_p(a, C.c_uint8), len(a), _p(out, C.c_uint8), size)
    if n < 0:
        raise ValueError("malformed LZ4 block")
    return bytes(out[:n])


def bench_mesh_isolated(blocks: np.ndarray, nthreads: int, L=None):
    """Time Chunk::BuildMesh (render attributes) over n isolated chunks -> (seconds, faces)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint16).reshape(-1, NBLK)
    faces = C.c_uint64(0)
    t = (L or lib()).orc_bench_mesh_isolated(_p(b, C.c_uint16), len(b), nthreads, C.byref(faces))
    return float(t), int(faces.value)


def bench_planet(seed, chunk_count, nthreads, stride=1, L=None):
    """Generate a planet and mesh every stride-th chunk -> dict(gen_s, mesh_s, faces, meshed, ok)."""
    times = (C.c_double * 2)()
    faces = C.c_uint64(0)
    meshed = C.c_uint32(0)
    rc = (L or lib()).orc_bench_planet(seed & 0xFFFFFFFF, _p(_u3(chunk_count), C.c_uint32), nthreads, stride, times, C.byref(faces), C.byref(meshed))
    return dict(gen_s=times[0], mesh_s=times[1], faces=int(faces.value), meshed=int(meshed.value), ok=rc == 0)


class _Reference:
    """The same front-end bound to oracle/_ref/libtsom_ref.so — the reference's own sources (see the module docstring)."""

    def __getattr__(self, name):
        target = globals()[name]
        if name in ("Planet", "Ship", "PlanetSample"):
            cls = target
            return lambda *a, **k: cls(*a, L=ref_lib(), **k)
        if callable(target):
            return lambda *a, **k: target(*a, L=ref_lib(), **k)
        return target


def reference() -> _Reference:
    ref_lib()
    return _Reference()
