"""CPU: libtsom_b200.so loads, exports every symbol include/tsom_b200.h declares, the ctypes mirror matches the header's
struct layouts, and — with no CUDA device — every compute entry point fails loudly (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from thisspaceofmine_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tsom_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?(?:int|void|uint32_t|uint64_t|char|tsom_ctx|tsom_container)\s*\*?\s*(tsom_[a-z0-9_]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_what_ctypes_binds():
    names = declared_functions()
    assert len(names) >= 30
    assert names == sorted(L.SYMBOLS), set(names) ^ set(L.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L.build()
    assert os.path.exists(L.SO_PATH)
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.SO_PATH], text=True)
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    lib = L.lib()  # dlopen + getattr of every symbol
    assert lib.tsom_version().decode().startswith("tsom_b200")


def test_library_is_sm_100a_only():
    out = subprocess.check_output(["cuobjdump", "--list-elf", L.SO_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_the_header():
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirror."""
    import tempfile

    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "tsom_b200.h"
int main(void) {
  printf("%zu %zu %zu ", sizeof(tsom_box_view), offsetof(tsom_box_view, box_count), (size_t)TSOM_TIMING_COUNT);
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(tsom_block_desc), sizeof(tsom_gen_blocks), sizeof(tsom_platform_blocks), sizeof(tsom_edit),
         offsetof(tsom_edit, block), sizeof(tsom_mesh_view), offsetof(tsom_mesh_view, face_count), sizeof(tsom_mesh_params), offsetof(tsom_mesh_params, gravity_centers),
         sizeof(tsom_ipc_handle));
  return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-o", exe])  # also proves the header is plain C
        got = [int(v) for v in subprocess.check_output([exe], text=True).split()]
    want = [C.sizeof(L.BoxView), L.BoxView.box_count.offset, L.TIMING_COUNT, C.sizeof(L.BlockDesc), C.sizeof(L.GenBlocks), C.sizeof(L.PlatformBlocks), C.sizeof(L.Edit), L.Edit.block.offset, C.sizeof(L.MeshView),
            L.MeshView.face_count.offset, C.sizeof(L.MeshParams), L.MeshParams.gravity_centers.offset, C.sizeof(L.IpcHandle)]
    assert got == want


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a CUDA device")
def test_no_device_fails_loudly():
    """No CPU fallback: without a GPU tsom_create returns TSOM_ERR_CUDA with a message, and the host mirror raises."""
    import thisspaceofmine_b200 as T

    h = C.c_void_p()
    rc = L.lib().tsom_create(0, C.byref(h))
    assert rc == L.ERR_CUDA and not h.value
    assert L.lib().tsom_last_error()
    with pytest.raises(T.TsomError) as e:
        T.Context(0)
    assert e.value.code == L.ERR_CUDA


def test_null_handles_are_rejected_not_dereferenced():
    lib = L.lib()
    assert lib.tsom_synchronize(None) != L.OK
    assert lib.tsom_mesh(None, None, 0, None, None) != L.OK
    assert lib.tsom_apply_edits(None, None, 0) != L.OK
    assert lib.tsom_halo_pull(None) != L.OK
    assert lib.tsom_halo_publish(None) != L.OK
    assert lib.tsom_box_colliders(None, None, 0, None, None) != L.OK
    assert lib.tsom_sharded_generate(None, 0, None) != L.OK
    assert lib.tsom_sharded_mesh(None, None, 0, None, None, None) != L.OK
    assert lib.tsom_sharded_part_count(None) == 0 and not lib.tsom_sharded_ctx(None, 0)
    lib.tsom_sharded_destroy(None)
    out = C.c_void_p()
    assert lib.tsom_planet_create_sharded(None, 2, None, 1.0, None, C.byref(out)) != L.OK and not out.value
    assert lib.tsom_container_chunk_count(None) == 0
    assert lib.tsom_launch_count(None) == 0
    n = C.c_uint32()
    idx = np.zeros(3, np.int32)
    assert lib.tsom_collect_dirty(None, idx.ctypes.data_as(C.POINTER(C.c_int32)), 1, C.byref(n), 0) != L.OK


def test_numpy_record_layouts_match_the_ctypes_mirror():
    """host.py passes numpy record arrays straight through the C ABI (no per-element Python work): same sizes and offsets."""
    from thisspaceofmine_b200.host import BOX_VIEW_DTYPE, EDIT_DTYPE, VIEW_DTYPE

    assert EDIT_DTYPE.itemsize == C.sizeof(L.Edit)
    assert VIEW_DTYPE.itemsize == C.sizeof(L.MeshView)
    assert BOX_VIEW_DTYPE.itemsize == C.sizeof(L.BoxView) and BOX_VIEW_DTYPE.fields["box_count"][1] == L.BoxView.box_count.offset
    for name, _ in L.Edit._fields_:
        assert EDIT_DTYPE.fields[name][1] == getattr(L.Edit, name).offset, name
    for name, _ in L.MeshView._fields_:
        assert VIEW_DTYPE.fields[name][1] == getattr(L.MeshView, name).offset, name


def test_indices_for_faces_is_the_reference_pattern():
    """tsom_mesh_indices_for_faces (host code, no device): 4k + {0, 2, 1, 1, 2, 3} per face (src/CommonLib/Chunk.cpp:95-101), also on
    the multi-threaded path (> 2^20 faces)."""
    lib = L.lib()
    for faces in (0, 1, 7, (1 << 20) + 5):
        out = np.full(6 * faces + 3, 0xDEADBEEF, np.uint32)
        assert lib.tsom_mesh_indices_for_faces(faces, out.ctypes.data_as(C.POINTER(C.c_uint32))) == L.OK
        want = (4 * np.arange(faces, dtype=np.uint32)[:, None] + np.array([0, 2, 1, 1, 2, 3], np.uint32)[None, :]).reshape(-1)
        assert np.array_equal(out[: 6 * faces], want) and np.all(out[6 * faces:] == 0xDEADBEEF)
    assert lib.tsom_mesh_indices_for_faces(1 << 31, out.ctypes.data_as(C.POINTER(C.c_uint32))) != L.OK  # vertex ordinals would not fit 32 bits
    assert lib.tsom_mesh_indices_for_faces(3, None) != L.OK
