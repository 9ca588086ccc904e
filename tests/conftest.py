import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from _checkers import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from _checkers import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libfastc_ref.so not built (needs /root/reference at build time)")
    return Reference()


@pytest.fixture(scope="session")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box (there is no CPU fallback)")
    from fastc_b200 import lib
    return lib()
