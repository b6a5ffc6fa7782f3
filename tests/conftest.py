import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def _session_ctx():
    """One device context for the whole GPU session (tsom_create on cuda:0)."""
    import thisspaceofmine_b200 as T

    c = T.Context(0)
    yield c
    c.close()


@pytest.fixture
def ctx(_session_ctx):
    """The session context; after every test the 0xA5 guards behind the mesh arenas must be intact (tsom_debug_check_guards)."""
    yield _session_ctx
    _session_ctx.check_guards()
