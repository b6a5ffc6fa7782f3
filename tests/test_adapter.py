"""The C++ host layer (adapter/tsom_adapter.hpp) above the C ABI.

CPU: it compiles and links against libtsom_b200.so, and the reference's own unit test (tests/UnitTests/Common/ChunkTests.cpp:8-46,
re-stated in adapter/test_adapter.cpp) plus the BlockLibrary registration checks pass without touching CUDA.
GPU: `test_adapter --gpu` drives generate -> batched mesh -> per-chunk Chunk::BuildMesh -> collider -> edit -> dirty set through the adapter."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")
BIN = os.path.join(ADAPTER, "test_adapter")


def _build():
    from thisspaceofmine_b200 import _lib as L

    L.build()
    subprocess.check_call(["make", "-C", ADAPTER], stdout=subprocess.DEVNULL)
    assert os.path.exists(BIN)


def test_adapter_builds_and_reference_kats_pass():
    _build()
    out = subprocess.run([BIN], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK (0 failed)" in out.stdout


@pytest.mark.gpu
def test_adapter_end_to_end_on_gpu():
    if not os.path.exists(BIN):
        _build()
    out = subprocess.run([BIN, "--gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "adapter GPU pass" in out.stdout and "OK (0 failed)" in out.stdout
