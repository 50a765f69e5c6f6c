"""Header-level extensions without a C-ABI entry, exercised by the CUDA test
programs in tests/cuda/ (built into tests/_bin/ by yalla_b200/build.py):
Vtk_async_output -- asynchronous frames identical to Vtk_output's --, the
seeded generators through the header API, and Cell_division (reproducible
cell division: count statistics, placement, inheritance, order, clamping)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_extension_programs(tmp_path):
    binary = os.path.join(ROOT, "tests", "_bin", "test_extensions")
    assert os.path.exists(binary), "run __graft_entry__.build() first"
    result = subprocess.run([binary, str(tmp_path) + "/"], capture_output=True,
                            text=True, timeout=600)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    assert "all extension checks passed" in result.stdout
    assert result.stdout.count("ok ") >= 16


def test_bricks_from_cuda_cpp():
    """The decomposition driven from CUDA C++ through the header API alone: two
    bricks in one process, connected by plain pointers, identities registered as
    a travelling array, against the single-domain run (tests/cuda/test_bricks.cu)."""
    binary = os.path.join(ROOT, "tests", "_bin", "test_bricks")
    assert os.path.exists(binary), "run __graft_entry__.build() first"
    result = subprocess.run([binary], capture_output=True, text=True, timeout=600)
    assert result.returncode == 0, result.stdout[-2000:] + result.stderr[-2000:]
    assert "all brick checks passed" in result.stdout
