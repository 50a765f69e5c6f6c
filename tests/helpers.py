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


def polarity_vectors(X):
    theta, phi = X[:, 3].astype(np.float64), X[:, 4].astype(np.float64)
    return np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi),
                     np.cos(theta)], axis=1)


def assert_states_close(got, want, steps, what, factor=1.0):
    """Positions, polarities and further lanes each on their OWN max-norm:
    columns 0:3 relative to the largest coordinate; the polarity (theta, phi in
    columns 3:5) as a unit vector, i.e. on scale 1 -- the angles themselves are
    ill-conditioned at the coordinate poles, d phi / dt ~ 1 / sin theta; any
    further lanes relative to their own largest value."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert np.all(np.isfinite(got)), f"{what}: non-finite values"
    limit = REL_TOL_PER_STEP * steps * factor
    error = max_norm_error(got[:, :3], want[:, :3])
    assert error <= limit, f"{what}: max-norm error {error:.3e} > {limit:.1e}"
    if want.shape[1] >= 5:
        error = float(np.max(np.abs(polarity_vectors(got) - polarity_vectors(want))))
        assert error <= limit, f"{what}: polarity error {error:.3e} > {limit:.1e}"
    if want.shape[1] > 5:
        error = max_norm_error(got[:, 5:], want[:, 5:])
        assert error <= limit, f"{what}: lanes 5+ error {error:.3e} > {limit:.1e}"


def run_case(lib, case):
    return make_golden.run_case(lib, case)


def model_cases():
    return make_golden.cases()
