import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle_lib import Ref, have_ref

    if not have_ref():
        pytest.skip("oracle/_ref/libcuzk_ref.so not available")
    return Ref()


@pytest.fixture(scope="session")
def gpu():
    """Initialised library + torch device; fails (does not skip) when the CUDA library cannot run."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    from cuzk_b200 import api

    api.initialize(0)
    yield api
    api.cleanup()
