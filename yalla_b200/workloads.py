"""Deterministic synthetic inputs for the parity tests and the benchmark.

The reference's generators (include/inits.cuh:14-75) draw from rand() seeded by
std::random_device and relax with thousands of steps, so no shipped model is
reproducible. These generators produce the same KIND of state
from a numpy Generator:

* ``random_ball``   -- uniform ball of radius (n / 0.64)^(1/3) * d / 2, the
  formula of random_sphere (inits.cuh:41-48);
* ``lattice_ball``  -- a "relaxed" tissue without the relaxation run: a jittered
  FCC lattice with nearest-neighbour distance d, cut to the n cells nearest to
  the origin and SHUFFLED, so that memory order carries no spatial locality
  (as after random_sphere + relaxation);
* ``polarized_ball``-- Po_cell state: positions as above, polarity pointing
  outwards plus U(0, 0.5) noise on both angles (examples/epithelium.cu:39-47).
"""
import numpy as np


def ball_radius(n, dist_to_nb):
    return (n / 0.64) ** (1.0 / 3.0) * dist_to_nb / 2.0


def grid_size_for(n, dist_to_nb, cube_size=1.0, growth=1.0, lattice=True):
    """Smallest even grid_size that holds a ball of n * growth cells plus the
    one-cube margin the 27-cube sweep needs on every side."""
    # FCC packs denser than a random ball: n = sqrt(2) / d^3 * 4/3 pi R^3
    d = dist_to_nb
    if lattice:
        radius = (n * growth * d ** 3 / np.sqrt(2.0) * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    else:
        radius = ball_radius(n * growth, d)
    cubes = int(np.ceil(2.0 * (radius + d) / cube_size)) + 4
    return cubes + (cubes % 2)


def random_ball(n, dist_to_nb, rng):
    r = ball_radius(n, dist_to_nb) * rng.random(n) ** (1.0 / 3.0)
    theta = np.arccos(2.0 * rng.random(n) - 1.0)
    phi = rng.random(n) * 2.0 * np.pi
    X = np.empty((n, 3), dtype=np.float32)
    X[:, 0] = r * np.sin(theta) * np.cos(phi)
    X[:, 1] = r * np.sin(theta) * np.sin(phi)
    X[:, 2] = r * np.cos(theta)
    return X


def lattice_ball(n, dist_to_nb, rng, jitter=0.05):
    d = float(dist_to_nb)
    a = d * np.sqrt(2.0)  # conventional FCC cell edge for nn distance d
    radius = (n * d ** 3 / np.sqrt(2.0) * 3.0 / (4.0 * np.pi)) ** (1.0 / 3.0)
    half = int(np.ceil(radius / a)) + 2
    axis = np.arange(-half, half + 1, dtype=np.float64) * a
    gx, gy, gz = np.meshgrid(axis, axis, axis, indexing="ij")
    corner = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    points = (corner[:, None, :] + basis[None, :, :]).reshape(-1, 3)
    dist2 = np.einsum("ij,ij->i", points, points)
    if len(points) < n:
        raise ValueError("lattice too small")
    keep = np.argpartition(dist2, n - 1)[:n]
    points = points[keep]
    points += (rng.random(points.shape) - 0.5) * 2.0 * jitter * d
    rng.shuffle(points, axis=0)
    return points.astype(np.float32)


def polarized_ball(n, dist_to_nb, rng, lattice=False, noise=0.5):
    pos = lattice_ball(n, dist_to_nb, rng) if lattice else random_ball(
        n, dist_to_nb, rng)
    X = np.zeros((n, 5), dtype=np.float32)
    X[:, :3] = pos
    dist = np.maximum(np.linalg.norm(pos.astype(np.float64), axis=1), 1e-12)
    X[:, 3] = np.arccos(np.clip(pos[:, 2] / dist, -1.0, 1.0)) + rng.random(n) * noise
    X[:, 4] = np.arctan2(pos[:, 1], pos[:, 0]) + rng.random(n) * noise
    return X


def shell_types(X, thickness=1.0):
    """1 (epithelium) for cells within `thickness` of the surface, else 0."""
    r = np.linalg.norm(X[:, :3].astype(np.float64), axis=1)
    return (r > r.max() - thickness).astype(np.int32)


def random_links(X, n_links, max_dist, rng, k=12):
    """n_links pairs (a, b), a != b, of cells closer than max_dist -- what a
    protrusion-update kernel produces (examples/sorting_prot.cu:33-74). Each
    link joins a random cell with a random one of its k nearest neighbours
    inside max_dist; links that find none stay inactive (a == b == 0)."""
    from scipy.spatial import cKDTree
    tree = cKDTree(X[:, :3])
    a = rng.integers(0, len(X), size=n_links)
    distance, index = tree.query(X[a, :3], k=k + 1, distance_upper_bound=max_dist,
                                 workers=-1)
    pick = rng.integers(1, k + 1, size=n_links)  # column 0 is the cell itself
    b = index[np.arange(n_links), pick]
    found = np.isfinite(distance[np.arange(n_links), pick]) & (b < len(X)) & (b != a)
    links = np.zeros((n_links, 2), dtype=np.int32)
    links[found, 0] = a[found]
    links[found, 1] = b[found]
    return links
