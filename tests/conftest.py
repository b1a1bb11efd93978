import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """The parity suite drives the one-thread-per-element kernels (BN254_IMPL=thread, read at context creation); the
    default context routes small pairing batches to the lane-group kernels, which test_small_batch_auto_routing and
    test_lane_group_vm_implementation cover."""
    from gopairingbasedcryptography_b200 import bn254

    old = os.environ.get("BN254_IMPL")
    os.environ["BN254_IMPL"] = "thread"
    try:
        return bn254.Engine(0)
    finally:
        if old is None:
            del os.environ["BN254_IMPL"]
        else:
            os.environ["BN254_IMPL"] = old
