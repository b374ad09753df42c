import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library; built on demand so that a fresh checkout can run the CPU suite."""
    from corrla_rs_b200 import _ffi
    if not _ffi.lib_path().exists():
        import __graft_entry__
        __graft_entry__.build()
    return _ffi.load()
