import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    from __graft_entry__ import load_package
    p = load_package()
    if not os.path.exists(p.LIB_PATH):
        p.build()
    return p


@pytest.fixture(scope="session")
def oracle():
    from _oracle import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref), or None when it has not been built."""
    from _oracle import load_ref
    return load_ref()


@pytest.fixture(scope="session")
def golden():
    import json
    d = os.path.join(ROOT, "tests", "golden")
    return {n[:-5]: json.load(open(os.path.join(d, n))) for n in os.listdir(d) if n.endswith(".json")}


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcudart.so")
    except OSError:
        cuda = None
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine(pkg):
    """Default (FMA-contracted) engine on cuda:0.  No GPU -> the test fails loudly (no fallback)."""
    e = pkg.PairHMMEngine(devices=[0])
    yield e
    e.close()


@pytest.fixture(scope="session")
def exact_engine(pkg):
    """exact_fp32=1: unfused arithmetic, raw FP32 bit-identical to the reference."""
    e = pkg.PairHMMEngine(devices=[0], exact_fp32=True)
    yield e
    e.close()
