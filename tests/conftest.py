import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py as O
    O.build(ref=True)
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_oix(oracle):
    ix = oracle.Index(os.path.join(GOLDEN, "ref.ufi"))
    yield ix
    ix.close()


@pytest.fixture(scope="session")
def built_lib():
    """liburmb.so built in-tree (nvcc cross-compiles without a GPU)."""
    from urmap_b200 import build as B
    B.build_engine()
    from urmap_b200 import engine
    return engine


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
