"""Helpers shared by the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import make_golden  # noqa: E402  (the generator doubles as the case list)

# north_star: positions and polarities after K noise-free steps agree within
# 1e-5 relative per step on the max-norm.
REL_TOL_PER_STEP = 1e-5


def max_norm_error(a, b):
    scale = max(float(np.max(np.abs(b))), 1.0)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / scale


def assert_states_close(got, want, steps, what, factor=1.0):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    error = max_norm_error(got, want)
    limit = REL_TOL_PER_STEP * steps * factor
    assert error <= limit, f"{what}: max-norm error {error:.3e} > {limit:.1e}"


def run_case(lib, case):
    return make_golden.run_case(lib, case)


def model_cases():
    return make_golden.cases()
