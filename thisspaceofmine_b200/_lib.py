"""ctypes loader of libtsom_b200.so (the C ABI declared in include/tsom_b200.h).

The library is built in-tree (thisspaceofmine_b200/csrc/Makefile -> thisspaceofmine_b200/libtsom_b200.so) for sm_100a only.
There is no CPU fallback: if the library is missing, or no B200-class device is visible, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("TSOM_B200_SO") or os.path.join(_HERE, "libtsom_b200.so")  # the override is a development aid (kernel variants)
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NOMEM, ERR_CAPACITY, ERR_INTERNAL = range(7)
MESH_RENDER, MESH_COLLIDER, MESH_DEFORMED, MESH_GENERIC_MATH, MESH_ORDERED, MESH_TANGENTS = 1, 2, 4, 8, 16, 32
CONTAINER_PLANET, CONTAINER_SHIP = 0, 1
TIMING_COUNT = 8


class TsomError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tsom error {code}: {msg}")
        self.code = code


class BlockDesc(C.Structure):
    _fields_ = [("tex", C.c_uint32 * 6), ("has_collisions", C.c_uint8), ("is_double_sided", C.c_uint8), ("is_transparent", C.c_uint8), ("reserved", C.c_uint8)]


class GenBlocks(C.Structure):
    _fields_ = [(n, C.c_uint16) for n in ("dirt", "grass", "stone", "stone_mossy", "snow")]


class PlatformBlocks(C.Structure):
    _fields_ = [(n, C.c_uint16) for n in ("copper_block", "stone_bricks", "planks")]


class Edit(C.Structure):
    _fields_ = [("chunk", C.c_int32 * 3), ("x", C.c_uint8), ("y", C.c_uint8), ("z", C.c_uint8), ("reserved", C.c_uint8), ("block", C.c_uint16), ("reserved2", C.c_uint16)]


class MeshView(C.Structure):
    _fields_ = [("first_face", C.c_uint64), ("face_count", C.c_uint32), ("reserved", C.c_uint32)]


class MeshParams(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("deform_radius", C.c_float), ("gravity_centers", C.POINTER(C.c_float))]


class BoxView(C.Structure):
    _fields_ = [("first_box", C.c_uint64), ("box_count", C.c_uint32), ("reserved", C.c_uint32)]


class IpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_ubyte * 200)]


# every symbol include/tsom_b200.h declares: name -> (restype, argtypes)
_vp, _i32p, _u32p, _u16p, _f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_float)
SYMBOLS = {
    "tsom_last_error": (C.c_char_p, []),
    "tsom_version": (C.c_char_p, []),
    "tsom_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "tsom_destroy": (None, [_vp]),
    "tsom_stream": (_vp, [_vp]),
    "tsom_synchronize": (C.c_int, [_vp]),
    "tsom_set_block_table": (C.c_int, [_vp, C.POINTER(BlockDesc), C.c_uint32]),
    "tsom_launch_count": (C.c_uint64, [_vp]),
    "tsom_ctx_lock": (C.c_int, [_vp]),
    "tsom_ctx_unlock": (C.c_int, [_vp]),
    "tsom_reserve_faces": (C.c_int, [_vp, C.c_uint64, C.c_uint32]),
    "tsom_debug_check_guards": (C.c_int, [_vp]),
    "tsom_get_timings": (C.c_int, [_vp, _f32p, C.c_uint32]),
    "tsom_container_create": (C.c_int, [_vp, C.c_uint32, C.c_float, C.POINTER(_vp)]),
    "tsom_container_destroy": (None, [_vp]),
    "tsom_container_add_chunks": (C.c_int, [_vp, _i32p, C.c_uint32]),
    "tsom_container_remove_chunk": (C.c_int, [_vp, _i32p]),
    "tsom_container_chunk_count": (C.c_uint32, [_vp]),
    "tsom_chunks_reset": (C.c_int, [_vp, _i32p, C.c_uint32, _u16p]),
    "tsom_chunks_reset_device": (C.c_int, [_vp, _i32p, C.c_uint32, _vp]),
    "tsom_chunks_get_content": (C.c_int, [_vp, _i32p, C.c_uint32, _u16p]),
    "tsom_chunk_device_ptr": (_vp, [_vp, _i32p]),
    "tsom_chunks_palettize": (C.c_int, [_vp, _i32p, C.c_uint32, _u16p, C.c_uint32, _u32p, C.POINTER(C.c_uint8), C.c_int]),
    "tsom_chunks_reset_palettized": (C.c_int, [_vp, _i32p, C.c_uint32, _u16p, C.c_uint32, _u32p, C.POINTER(C.c_uint8), C.c_int]),
    "tsom_chunks_lz4_compress": (C.c_int, [_vp, _i32p, C.c_uint32, C.POINTER(C.c_uint8), C.c_uint64, C.POINTER(C.c_uint64)]),
    "tsom_chunks_reset_lz4": (C.c_int, [_vp, _i32p, C.c_uint32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)]),
    "tsom_planet_generate": (C.c_int, [_vp, C.c_uint32, _u32p, C.POINTER(GenBlocks)]),
    "tsom_planet_generate_chunks": (C.c_int, [_vp, C.c_uint32, _u32p, C.POINTER(GenBlocks), _i32p, C.c_uint32]),
    "tsom_planet_generate_chunks_async": (C.c_int, [_vp, C.c_uint32, _u32p, C.POINTER(GenBlocks), _i32p, C.c_uint32]),
    "tsom_planet_generate_platform": (C.c_int, [_vp, C.c_int, _i32p, C.POINTER(PlatformBlocks)]),
    "tsom_apply_edits": (C.c_int, [_vp, C.POINTER(Edit), C.c_uint32]),
    "tsom_collect_dirty": (C.c_int, [_vp, _i32p, C.c_uint32, _u32p, C.c_int]),
    "tsom_mesh": (C.c_int, [_vp, _i32p, C.c_uint32, C.POINTER(MeshParams), C.POINTER(MeshView)]),
    "tsom_mesh_arenas": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_uint64)]),
    "tsom_mesh_copy_to_host": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _u32p, _f32p]),
    "tsom_mesh_aabb": (C.c_int, [_vp, C.POINTER(MeshView), C.c_uint32, _f32p]),
    "tsom_mesh_checksums": (C.c_int, [_vp, C.POINTER(MeshView), C.c_uint32, C.POINTER(C.c_uint64)]),
    "tsom_mesh_indices_for_faces": (C.c_int, [C.c_uint64, _u32p]),
    "tsom_voxel_corners": (C.c_int, [_vp, _i32p, _u32p, C.c_uint32, C.POINTER(MeshParams), _f32p]),
    "tsom_box_collider": (C.c_int, [_vp, _i32p, _f32p, C.c_uint32, _u32p]),
    "tsom_container_add_remote_chunks": (C.c_int, [_vp, _i32p, _u32p, _i32p, C.c_uint32]),
    "tsom_chunk_slot": (C.c_int, [_vp, _i32p, _u32p]),
    "tsom_ipc_export": (C.c_int, [_vp, C.POINTER(IpcHandle)]),
    "tsom_ipc_attach": (C.c_int, [_vp, C.c_int, C.POINTER(IpcHandle)]),
    "tsom_peer_attach_local": (C.c_int, [_vp, C.c_int, _vp]),
    "tsom_halo_publish": (C.c_int, [_vp]),
    "tsom_halo_pull": (C.c_int, [_vp]),
    "tsom_halo_set_timeout": (C.c_int, [_vp, C.c_uint32]),
    "tsom_box_colliders": (C.c_int, [_vp, _i32p, C.c_uint32, C.POINTER(BoxView), C.POINTER(C.c_uint64)]),
    "tsom_box_colliders_copy_to_host": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _f32p]),
    "tsom_container_set_kind": (C.c_int, [_vp, C.c_int]),
    "tsom_ship_generate": (C.c_int, [_vp, C.c_int, C.c_uint16, C.c_uint16]),
    "tsom_shard_plan": (C.c_int, [_u32p, C.c_uint32, _f32p, _i32p]),
    "tsom_planet_create_sharded": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, _u32p, C.c_float, _f32p, C.POINTER(_vp)]),
    "tsom_sharded_destroy": (None, [_vp]),
    "tsom_sharded_part_count": (C.c_uint32, [_vp]),
    "tsom_sharded_ctx": (_vp, [_vp, C.c_uint32]),
    "tsom_sharded_container": (_vp, [_vp, C.c_uint32]),
    "tsom_sharded_part_box": (C.c_int, [_vp, C.c_uint32, _i32p]),
    "tsom_sharded_owner": (C.c_int, [_vp, _i32p, _u32p]),
    "tsom_sharded_set_block_table": (C.c_int, [_vp, C.POINTER(BlockDesc), C.c_uint32]),
    "tsom_sharded_generate": (C.c_int, [_vp, C.c_uint32, C.POINTER(GenBlocks)]),
    "tsom_sharded_apply_edits": (C.c_int, [_vp, C.POINTER(Edit), C.c_uint32]),
    "tsom_sharded_collect_dirty": (C.c_int, [_vp, _i32p, C.c_uint32, _u32p, C.c_int]),
    "tsom_sharded_mesh": (C.c_int, [_vp, _i32p, C.c_uint32, C.POINTER(MeshParams), C.POINTER(MeshView), _u32p]),
    "tsom_sharded_mesh_copy_to_host": (C.c_int, [_vp, C.c_uint32, C.c_uint64, C.c_uint64, _vp, _u32p, _f32p]),
    "tsom_sharded_get_content": (C.c_int, [_vp, _i32p, C.c_uint32, _u16p]),
    "tsom_sharded_get_timings": (C.c_int, [_vp, _f32p]),
}


def build(force: bool = False) -> str:
    """Compile libtsom_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC]
    if force:
        args.append("-B")
    subprocess.check_call(args, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL if not force else None)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise TsomError(rc, lib().tsom_last_error().decode(errors="replace"))
