"""CPU-only checks of the boundary: every entry point include/yalla_b200.h
declares is exported by the product library (and by both checkers), the Python
binding mirrors the header, and the host-side helpers behave. No compute call is
made on the product library here -- that needs a GPU (tests/test_gpu_parity.py).
"""
import ctypes
import os
import re

import numpy as np
import pytest

import yalla_b200 as yb
from yalla_b200 import workloads
from conftest import ORACLE_LIB, ROOT


def declared_functions():
    header = open(os.path.join(ROOT, "include", "yalla_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(yb_[a-z0-9_]+)\s*\(", header)))


def test_header_declares_what_the_binding_binds():
    assert declared_functions() == sorted(yb.SIGNATURES)


def test_product_library_exports_every_symbol():
    # loading must work without a GPU; a missing library is an error, not a skip
    lib = ctypes.CDLL(yb.PRODUCT_LIB)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} not exported by libyalla_b200.so"
    lib.yb_build_info.restype = ctypes.c_char_p
    assert lib.yb_build_info().decode().startswith("yalla-b200")


def test_checker_libraries_export_the_same_abi(oracle):
    names = declared_functions()
    for path in (ORACLE_LIB, yb.REFERENCE_LIB):
        if not os.path.exists(path):
            continue
        lib = ctypes.CDLL(path)
        for name in names:
            assert hasattr(lib, name), f"{name} not exported by {path}"


def test_missing_library_fails_loudly():
    with pytest.raises(yb.YallaError):
        yb.load(os.path.join(ROOT, "yalla_b200", "_lib", "no_such_library.so"))


def test_product_package_does_not_touch_the_oracle():
    # the oracle is test infrastructure: no shipped code may import, load or
    # include it (the C header only names it in its explanatory comment)
    for folder in ("yalla_b200", "include"):
        for base, _, files in os.walk(os.path.join(ROOT, folder)):
            for name in files:
                if name.endswith((".py", ".cuh", ".cu")):
                    text = open(os.path.join(base, name)).read()
                    assert "libyalla_oracle" not in text, os.path.join(base, name)
                    assert "yalla_oracle.cpp" not in text, os.path.join(base, name)


def test_model_table_matches_oracle(oracle):
    for model, lanes in yb.MODEL_LANES.items():
        if model in yb.GPU_ONLY_MODELS:
            with pytest.raises(yb.YallaError, match="unknown model"):
                oracle.sim(model, 8)
            continue
        with oracle.sim(model, 8) as sim:
            assert sim.lanes == lanes
            assert sim.n_max == 8


def test_error_reporting(oracle):
    with pytest.raises(yb.YallaError, match="unknown model"):
        oracle.sim("no_such_model", 8)
    with oracle.sim("relu_grid", 4) as sim:
        with pytest.raises(yb.YallaError, match="n > n_max"):
            sim.set_state(np.zeros((5, 3), dtype=np.float32))
        with pytest.raises(yb.YallaError, match="unknown parameter"):
            sim.set_param("bogus", 1.0)


# ---- workloads --------------------------------------------------------------
def test_workloads_are_deterministic():
    a = workloads.lattice_ball(5000, 0.8, np.random.default_rng(1))
    b = workloads.lattice_ball(5000, 0.8, np.random.default_rng(1))
    assert np.array_equal(a, b)
    assert a.dtype == np.float32 and a.shape == (5000, 3)


def test_lattice_ball_is_relaxed_like():
    from scipy.spatial import cKDTree
    X = workloads.lattice_ball(20000, 0.8, np.random.default_rng(2))
    tree = cKDTree(X)
    nearest = tree.query(X, k=2)[0][:, 1]
    assert 0.7 < nearest.mean() < 0.85
    neighbours = tree.query_ball_point(X[:2000], 1.0, return_length=True) - 1
    assert 9 < neighbours.mean() < 13  # ~12 within r < 1, fewer at the surface


def test_grid_size_holds_the_ball():
    for n in (1000, 100000, 1000000):
        gs = workloads.grid_size_for(n, 0.8)
        X = workloads.lattice_ball(min(n, 100000), 0.8, np.random.default_rng(3))
        assert gs % 2 == 0
        if n <= 100000:
            assert np.abs(X).max() < gs / 2 - 1


def test_random_links_are_local():
    rng = np.random.default_rng(4)
    X = workloads.lattice_ball(2000, 0.8, rng)
    links = workloads.random_links(X, 500, 2.0, rng)
    live = links[links[:, 0] != links[:, 1]]
    assert len(live) > 400
    assert np.all(np.linalg.norm(X[live[:, 0]] - X[live[:, 1]], axis=1) <= 2.0)
