"""Shared fixtures. Tests that need a GPU carry @pytest.mark.gpu; everything
else must pass on a CPU-only box (python -m pytest tests -m "not gpu")."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))  # make_golden: the case list
sys.path.insert(0, os.path.join(ROOT, "tests"))

import yalla_b200 as yb  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libyalla_oracle.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _built(path, target):
    """The checker libraries travel prebuilt; (re)build them where possible."""
    if not os.path.exists(path):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), target],
                       check=False, capture_output=True)
    return os.path.exists(path)


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (test infrastructure, oracle/)."""
    if not _built(ORACLE_LIB, "oracle"):
        pytest.fail("oracle/_build/libyalla_oracle.so missing and not buildable")
    return yb.load(ORACLE_LIB)


@pytest.fixture(scope="session")
def product():
    """This repo's CUDA build. Missing library = failure, never a fallback."""
    return yb.product()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference headers compiled for sm_100a, if present."""
    if not _built(yb.REFERENCE_LIB, "ref"):
        pytest.skip("oracle/_ref/libyalla_ref.so not available here")
    return yb.reference()


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))
