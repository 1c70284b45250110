import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_libm():
    from oracle.oracle import Oracle
    return Oracle(libm=True)


@pytest.fixture(scope="session")
def mirror():
    import host_mirror
    return host_mirror.load()


@pytest.fixture(scope="session")
def ossg():
    import ossg_b200
    return ossg_b200


# OSSG_TEST_POISON=1: every floating-point CUDA tensor that torch.empty hands out during the tests is filled with
# NaN first, so an output element that no kernel writes shows as a mismatch instead of as whatever an earlier
# build left in the recycled allocation (tools/gpu_*.sh run the GPU suite once this way).
if os.environ.get("OSSG_TEST_POISON"):
    import torch as _torch
    _orig_empty = _torch.empty

    def _poisoned_empty(*args, **kwargs):
        t = _orig_empty(*args, **kwargs)
        if t.is_cuda and t.is_floating_point():
            t.fill_(float("nan"))
        return t

    _torch.empty = _poisoned_empty
