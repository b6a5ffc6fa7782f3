"""314 block types (ids far beyond the built-in 14): oracle vs the reference's own code on CPU, CUDA mesher vs oracle on the GPU.
The check runs in a subprocess because it appends to the oracle's process-wide BlockLibrary (tests/big_library_check.py)."""
import os
import subprocess
import sys

import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(mode):
    out = subprocess.run([sys.executable, os.path.join(HERE, "big_library_check.py"), mode], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "big library ok" in out.stdout


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref is not built and /root/reference is absent")
def test_big_block_library_oracle_vs_reference_code():
    _run("cpu")


@pytest.mark.gpu
def test_big_block_library_gpu_vs_oracle():
    _run("gpu")
