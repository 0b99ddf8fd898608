import os
import sys

import pytest

# The hash jobs switch to the radix-partitioned path above this many keys per bucket (default 2 M, measured optimum
# on B200). The tests lower it so that the partitioned code is exercised against the oracle at sizes the CPU side
# handles in seconds. Read once by libtermgpu.so, so it has to be set before the first execute.
os.environ.setdefault("TG_HASH_BUCKET_KEYS", "262144")
os.environ.setdefault("TG_HASH_GUESS_MIN_ROWS", "100000")
os.environ.setdefault("TG_STR_R2_MIN_ROWS", "1000")  # string kernel: 64-row blocks (two rows per lane) from 1000 rows on (default 2 M)  # the optimistic key-range guess of the dense path (default: 4 M rows)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library must exist (built by __graft_entry__.build()); build it if it is missing."""
    from term_b200 import _ffi
    if not os.path.exists(_ffi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _ffi.lib()


@pytest.fixture(scope="session")
def ctx(built_lib):
    """One engine on cuda:0 for the GPU tests (fails loudly without a device: no CPU fallback)."""
    import term_b200 as T
    c = T.SessionContext(0)
    yield c
    c.close()
