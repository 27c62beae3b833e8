import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full-size parity against the 1-thread reference (minutes of CPU time)")


@pytest.fixture(scope="session")
def port():
    """oracle/libnlk_port.so -- our C restatement (the checker, never the product)."""
    from oracle import oracle as O
    O.build(ref=False)
    return O.Port()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libnlkalman_ref.so -- the unmodified reference numerics, 1 thread.
    Prebuilt where /root/reference exists; travels to the GPU box with the snapshot."""
    from oracle import oracle as O
    if not os.path.exists(O.REF_SO):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return O.Ref(threads=1)


@pytest.fixture(scope="session")
def nlk():
    import bwd_nlkalman_b200 as m
    return m
