import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, like __graft_entry__.build()
    lib = os.path.join(ROOT, "rapidquadcoptercollisiondetection_b200", "librqcd.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import __graft_entry__
        __graft_entry__.build()


def _have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """oracle/liboracle.so: the C restatement (test infrastructure)."""
    from oracle import binding as B
    if not os.path.exists(B.ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return B.Oracle()


@pytest.fixture(scope="session")
def reference():
    """oracle/_ref/libref.so: the unmodified reference classes; absent on a box without /root/reference
    unless the prebuilt file travelled with the snapshot."""
    from oracle import binding as B
    if not B.have_reference():
        pytest.skip("oracle/_ref/libref.so not built")
    return B.Reference()


@pytest.fixture(scope="session")
def vectors():
    return np.load(os.path.join(GOLDEN, "ref_vectors.npz"))


@pytest.fixture(scope="session")
def forest_golden():
    return np.load(os.path.join(GOLDEN, "forest_golden.npz"))


@pytest.fixture(scope="session")
def golden_counts():
    import json
    return json.load(open(os.path.join(GOLDEN, "golden_counts.json")))


@pytest.fixture(scope="session")
def ctx():
    """The product: a context of librqcd.so on cuda:0 (no fallback: raises if the library is missing)."""
    from rapidquadcoptercollisiondetection_b200.api import Context
    c = Context(0)
    yield c
    c.close()
