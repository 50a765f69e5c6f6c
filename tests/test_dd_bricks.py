"""Brick decomposition over peer memory (include/b200/domain.cuh, dd.BrickDomain).

GPU (-m gpu): 2, 3 and 4 bricks of one tissue live in ONE process on one device,
each on its own stream, connected through plain device pointers -- the very
kernels, flags and epochs that run across GPUs over CUDA IPC
(scripts/dd_bricks_check.py under torchrun) -- and must reproduce the
single-domain run of the same library. CPU: the host-side layout logic.
"""
import numpy as np
import pytest

from yalla_b200 import dd, workloads


# ---- host logic (CPU) ----------------------------------------------------------
def test_brick_grids_and_ranks():
    assert dd.brick_grid_for(1) == (1, 1, 1)
    assert dd.brick_grid_for(2) == (1, 1, 2)
    assert dd.brick_grid_for(4) == (1, 2, 2)
    assert dd.brick_grid_for(8) == (2, 2, 2)
    assert dd.brick_grid_for(6) == (1, 2, 3)
    for world in (2, 4, 8, 12):
        bricks = dd.brick_grid_for(world)
        seen = set()
        for rank in range(world):
            coord = dd.brick_coord(rank, bricks)
            assert dd.brick_rank(coord, bricks) == rank
            seen.add(coord)
        assert len(seen) == world


def test_ball_brick_cuts_are_on_cube_boundaries_and_balanced():
    radius = 40.0
    cuts = dd.ball_brick_cuts(radius, (2, 2, 2))
    assert cuts == [[0.0], [0.0], [0.0]]
    cuts = dd.ball_brick_cuts(radius, (1, 1, 4))
    assert cuts[0] == [] and cuts[1] == [] and len(cuts[2]) == 3
    assert all(float(c).is_integer() for c in cuts[2])
    rng = np.random.default_rng(0)
    X = workloads.random_ball(200_000, 0.8, rng) * (radius / workloads.ball_radius(
        200_000, 0.8))
    share = np.histogram(X[:, 2], bins=[-np.inf] + cuts[2] + [np.inf])[0] / len(X)
    assert np.all(np.abs(share - 0.25) < 0.03)


class FakeSim:
    """Stands in for a library model: records what BrickDomain asks for."""
    lanes = 3

    def __init__(self):
        self.begun = None
        self.connected = {}
        self.mailboxes = {}

    def dom_register_array(self, address, width, ghosts_too):
        assert self.begun is None, "arrays are registered before dom_begin"
        self.arrays = getattr(self, "arrays", []) + [(address, width, ghosts_too)]

    def dom_begin(self, rank, world, lo, hi, halo, peers, caps, first, count):
        self.begun = dict(rank=rank, world=world, lo=np.array(lo), hi=np.array(hi),
                          halo=halo, peers=np.array(peers), caps=np.array(caps),
                          first=list(first), count=list(count))

    def dom_connect(self, direction, base, offsets):
        self.connected[direction] = (base, list(offsets))

    def dom_connect_mailbox(self, rank, base):
        self.mailboxes[rank] = base


class FakeLib:
    def sim(self, model, n_max, grid_size, cube_size):
        return FakeSim()


def test_brick_layout_of_a_corner_brick():
    # rank 5 of 2 x 2 x 2 = brick (1, 0, 1): neighbours towards -x, +y, -z
    bricks, cuts = (2, 2, 2), [[0.0], [0.0], [0.0]]
    brick = dd.BrickDomain(FakeLib(), "relu_grid", 1000, 64, 1.0, bricks, cuts, 5, 8,
                           face_capacity=1600)
    begun = brick.sim.begun
    assert brick.coord == (1, 0, 1)
    assert np.array_equal(begun["lo"], [0, -np.inf, 0])
    assert np.array_equal(begun["hi"], [np.inf, 0, np.inf])
    peers = begun["peers"]
    assert (peers >= 0).sum() == 7 and peers[13] == -1
    def index(dx, dy, dz):
        return (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1)
    assert peers[index(-1, 0, 0)] == dd.brick_rank((0, 0, 1), bricks)   # face
    assert peers[index(0, 1, 0)] == dd.brick_rank((1, 1, 1), bricks)
    assert peers[index(-1, 1, -1)] == dd.brick_rank((0, 1, 0), bricks)  # corner
    assert peers[index(1, 0, 0)] == -1 and peers[index(0, -1, 0)] == -1
    caps = begun["caps"]
    assert caps[index(-1, 0, 0)] == 1600                       # face
    assert caps[index(-1, 1, 0)] == 1600 // 16 + 2048          # edge
    assert caps[index(-1, 1, -1)] == 1600 // 256 + 2048        # corner
    # the box: from the halo (1.5) + 2 cubes of slack below the cut to the grid's end
    assert begun["first"] == [32 - 4, 0, 32 - 4] and begun["count"] == [36, 36, 36]
    # connecting: every neighbour's table is read at the OPPOSITE direction
    bases = list(range(100, 108))
    tables = [np.arange(27 * 8).reshape(27, 8) + 1000 * r for r in range(8)]
    brick.connect(bases, tables)
    base, offsets = brick.sim.connected[index(-1, 0, 0)]
    neighbour = dd.brick_rank((0, 0, 1), bricks)
    assert base == bases[neighbour]
    assert offsets == list(tables[neighbour][index(1, 0, 0)])
    assert brick.sim.mailboxes == {r: bases[r] for r in range(8)}


def test_model_arrays_are_registered_before_the_domain_begins():
    cuts = dd.ball_brick_cuts(30.0, (1, 1, 2))
    brick = dd.BrickDomain(FakeLib(), "relu_grid", 1000, 80, 1.0, (1, 1, 2), cuts, 1,
                           2, face_capacity=500, halo=2.5,
                           arrays=[(4096, 4, True), (8192, 48, False)])
    assert brick.sim.arrays == [(4096, 4, True), (8192, 48, False)]
    begun = brick.sim.begun
    assert begun["halo"] == 2.5
    # a wider halo widens the box of cubes the brick can touch
    lo_cube = int(np.floor(cuts[2][0] - 2.5)) - 2 + 40
    assert begun["first"][2] == lo_cube


def test_offset_tables_cover_four_round_kinds():
    import yalla_b200 as yb
    header = open(yb.PRODUCT_LIB.replace("yalla_b200/_lib/libyalla_b200.so",
                                         "include/yalla_b200.h")).read()
    assert f"#define YB_DOM_OFFSETS {yb.DOM_OFFSETS}" in header
    domain = open(yb.PRODUCT_LIB.replace("yalla_b200/_lib/libyalla_b200.so",
                                         "include/b200/domain.cuh")).read()
    assert f"constexpr int DD_ROUNDS = {yb.DOM_OFFSETS // 2};" in domain


def test_slab_bricks_have_two_neighbours_and_a_z_box():
    cuts = dd.ball_brick_cuts(30.0, (1, 1, 4))
    brick = dd.BrickDomain(FakeLib(), "relu_grid", 1000, 80, 1.0, (1, 1, 4), cuts, 2,
                           4, face_capacity=500)
    begun = brick.sim.begun
    assert sorted(np.nonzero(begun["peers"] >= 0)[0]) == [4, 22]  # -z and +z faces
    assert begun["first"][:2] == [0, 0] and begun["count"][:2] == [80, 80]
    lo_cube = int(np.floor(cuts[2][1] - 1.5)) - 2 + 40
    assert begun["first"][2] == lo_cube and begun["count"][2] < 80


# ---- the decomposed run on one GPU ------------------------------------------------
def match_cells(got, want, tol):
    from scipy.spatial import cKDTree
    assert got.shape == want.shape
    distance, index = cKDTree(want[:, :3]).query(got[:, :3], k=1)
    assert len(np.unique(index)) == len(want), "cells lost or duplicated"
    assert distance.max() < tol, f"max deviation {distance.max():.3e}"
    return float(np.max(np.abs(got - want[index])))


def run_bricks(product, model, X, bricks, steps, dt, gs):
    import torch
    world = bricks[0] * bricks[1] * bricks[2]
    radius = float(np.max(np.linalg.norm(X[:, :3], axis=1)))
    cuts = dd.ball_brick_cuts(radius, bricks)
    streams = [torch.cuda.Stream() for _ in range(world)]
    domains = [dd.BrickDomain(product, model, len(X), gs, 1.0, bricks, cuts, rank,
                              world, face_capacity=len(X))
               for rank in range(world)]
    for domain, stream in zip(domains, streams):
        domain.sim.set_stream(stream.cuda_stream)
    dd.connect_local(domains)
    before = []
    for domain in domains:
        mine = X[domain.owns(X)]
        before.append(len(mine))
        domain.set_cells(mine)
    for _ in range(steps):
        for domain in domains:
            domain.step(dt)
    torch.cuda.synchronize()
    parts, ghosts, after = [], 0, []
    for domain in domains:
        owned, with_ghosts, problems = domain.counts()
        assert problems == 0
        ghosts += with_ghosts - owned
        after.append(owned)
        parts.append(domain.owned_state()[0].cpu().numpy())
    for domain in domains:
        domain.close()
    return np.concatenate(parts), ghosts, before, after


@pytest.mark.gpu
@pytest.mark.parametrize("model,lanes,bricks", [
    ("relu_grid", 3, (1, 1, 2)), ("relu_grid", 3, (1, 1, 3)),
    ("relu_grid", 3, (1, 2, 2)), ("epithelium", 5, (1, 2, 2)),
    ("relu_grid", 3, (2, 2, 1))])
def test_bricks_match_single_domain(product, model, lanes, bricks):
    rng = np.random.default_rng(41)
    n, steps = 30_000, 6
    dt = 0.1 if lanes == 3 else 0.05
    if lanes == 3:
        # squeezed lattice: the tissue expands, cells cross the cuts
        X = workloads.lattice_ball(n, 0.8, rng) * 0.9
    else:
        X = workloads.polarized_ball(n, 0.8, rng, lattice=True)
        X[:, :3] *= 0.9
    X = X.astype(np.float32)
    gs = workloads.grid_size_for(n, 0.8) + 4
    with product.sim(model, n, gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(dt, steps)
        want = sim.get_state()
    got, ghosts, before, after = run_bricks(product, model, X, bricks, steps, dt, gs)
    assert len(got) == n and ghosts > 0
    assert before != after, "no cell migrated: the test is too tame"
    error = match_cells(got, want, 1e-3)
    assert error < 2e-5 * steps * max(float(np.max(np.abs(want))), 1.0)


@pytest.mark.gpu
def test_seeded_lattice_ball_is_the_same_tissue_however_it_is_cut(product):
    import torch
    radius, d = 20.0, 0.8
    gs = int(2 * radius) + 8
    tissues = []
    for bricks in ((1, 1, 1), (1, 2, 2)):
        world = bricks[0] * bricks[1] * bricks[2]
        cuts = dd.ball_brick_cuts(radius, bricks)
        domains = [dd.BrickDomain(product, "relu_grid", 120_000, gs, 1.0, bricks, cuts,
                                  rank, world, face_capacity=60_000)
                   for rank in range(world)]
        streams = [torch.cuda.Stream() for _ in domains]
        for domain, stream in zip(domains, streams):
            domain.sim.set_stream(stream.cuda_stream)
        dd.connect_local(domains)
        counts = [domain.seed_lattice_ball(radius, d, seed=7) for domain in domains]
        cells = np.concatenate([domain.owned_state()[0].cpu().numpy()
                                for domain in domains])
        assert sum(counts) == len(cells)
        for domain in domains:
            domain.close()
        tissues.append(cells[np.lexsort(cells.T[::-1])])
    assert tissues[0].shape == tissues[1].shape
    assert np.array_equal(tissues[0], tissues[1])
    # a brick that cannot hold its share says so
    small = dd.BrickDomain(product, "relu_grid", 1000, gs, 1.0, (1, 1, 1),
                           dd.ball_brick_cuts(radius, (1, 1, 1)), 0, 1, 1000)
    dd.connect_local([small])
    with pytest.raises(Exception, match="too small"):
        small.seed_lattice_ball(radius, d, seed=7)
    small.close()
    # density of an FCC lattice with nearest-neighbour distance d
    expected = np.sqrt(2.0) / d ** 3 * 4.0 / 3.0 * np.pi * radius ** 3
    assert abs(len(tissues[0]) - expected) < 0.02 * expected


# ---- per-cell arrays of the model travel with the cells ------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("bricks", [(1, 1, 2), (1, 2, 2)])
def test_registered_arrays_travel_with_their_cells(product, bricks):
    """Every brick registers an identity tag (ghosts_too) and a three-word
    payload (owned cells only). After steps with migration each cell must still
    carry ITS tag and payload: the cell tagged k sits where cell k of the
    single-domain run sits, so identity is tracked across the cuts."""
    import torch
    rng = np.random.default_rng(43)
    n, steps, dt = 30_000, 6, 0.1
    X = (workloads.lattice_ball(n, 0.8, rng) * 0.9).astype(np.float32)
    gs = workloads.grid_size_for(n, 0.8) + 4
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(dt, steps)
        want = sim.get_state()
    world = bricks[0] * bricks[1] * bricks[2]
    cuts = dd.ball_brick_cuts(float(np.max(np.linalg.norm(X, axis=1))), bricks)
    tags = [torch.full((n,), -1, dtype=torch.int32, device="cuda") for _ in range(world)]
    loads = [torch.zeros((n, 3), dtype=torch.float32, device="cuda")
             for _ in range(world)]
    domains = [dd.BrickDomain(product, "relu_grid", n, gs, 1.0, bricks, cuts, rank,
                              world, face_capacity=n,
                              arrays=[(tags[rank].data_ptr(), 4, True),
                                      (loads[rank].data_ptr(), 12, False)])
               for rank in range(world)]
    streams = [torch.cuda.Stream() for _ in domains]
    for domain, stream in zip(domains, streams):
        domain.sim.set_stream(stream.cuda_stream)
    dd.connect_local(domains)
    ids = np.arange(n, dtype=np.int32)
    before = []
    for rank, domain in enumerate(domains):
        mine = domain.owns(X)
        before.append(int(mine.sum()))
        domain.set_cells(X[mine])
        tags[rank][:before[-1]] = torch.as_tensor(ids[mine], device="cuda")
        loads[rank][:before[-1]] = torch.as_tensor(X[mine] * 2 + 1, device="cuda")
    torch.cuda.synchronize()
    for _ in range(steps):
        for domain in domains:
            domain.step(dt)
    torch.cuda.synchronize()
    seen, after = [], []
    for rank, domain in enumerate(domains):
        owned, with_ghosts, problems = domain.counts()
        assert problems == 0 and with_ghosts > owned
        after.append(owned)
        got = domain.owned_state()[0].cpu().numpy()
        tag = tags[rank][:owned].cpu().numpy()
        load = loads[rank][:owned].cpu().numpy()
        assert tag.min() >= 0
        assert np.max(np.abs(got - want[tag])) < 2e-5 * steps * np.max(np.abs(want))
        assert np.array_equal(load, X[tag] * 2 + 1)
        seen.append(tag)
    for domain in domains:
        domain.close()
    assert before != after, "no cell migrated: the test is too tame"
    assert np.array_equal(np.sort(np.concatenate(seen)), ids)


@pytest.mark.gpu
def test_growth_model_through_the_decomposed_step(product):
    """The Po_cell growth model (Property arrays, counter reset as generic force,
    curand division) run through yb_dom_step as a single brick: the migration
    pass re-stores cells, types, counters and curand states in cube order every
    step and adopts the daughters. Division off: cell by cell the same state,
    types and neighbour counters as the plain model; division on: the first
    round divides exactly the same mothers."""
    rng = np.random.default_rng(44)
    n = 20_000
    X = workloads.polarized_ball(n, 0.75, rng, lattice=True).astype(np.float32)
    types = workloads.shell_types(X)
    gs = workloads.grid_size_for(n, 0.75, growth=2.0)
    cuts = dd.ball_brick_cuts(float(np.max(np.linalg.norm(X[:, :3], axis=1))),
                              (1, 1, 1))

    def plain(rate, steps):
        with product.sim("growth", 2 * n, gs, 1.0) as sim:
            sim.set_param("prolif_rate", rate)
            sim.set_param("seed", 5)
            sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(0.1, steps)
            return (sim.get_state(), sim.get_ints("type"), sim.get_ints("mes_nbs"),
                    sim.get_ints("epi_nbs"))

    def decomposed(rate, steps):
        domain = dd.BrickDomain(product, "growth", 2 * n, gs, 1.0, (1, 1, 1), cuts,
                                0, 1, face_capacity=n)
        dd.connect_local([domain])
        domain.sim.set_param("prolif_rate", rate)
        domain.sim.set_param("seed", 5)
        domain.set_cells(X)
        domain.sim.set_ints("type", types)
        domain.step(0.1, steps)
        owned, _, problems = domain.counts()
        assert problems == 0
        out = (domain.owned_state()[0].cpu().numpy(), domain.sim.get_ints("type"),
               domain.sim.get_ints("mes_nbs"), domain.sim.get_ints("epi_nbs"))
        assert len(out[0]) == owned == len(out[1])
        domain.close()
        return out

    from scipy.spatial import cKDTree
    want, got = plain(0.0, 4), decomposed(0.0, 4)
    assert len(got[0]) == len(want[0]) == n
    distance, index = cKDTree(want[0][:, :3]).query(got[0][:, :3], k=1)
    assert len(np.unique(index)) == n and distance.max() < 1e-2
    # (a pair that sits on the cut-off within rounding may be counted on one
    # side in one run and on the other in the other: a handful of cells at most)
    assert np.sum(np.abs(got[0] - want[0][index]).max(axis=1) > 1e-4) <= 5
    assert np.array_equal(got[1], want[1][index])  # the type came along
    for k in (2, 3):  # mesenchymal and epithelial neighbour counts
        assert np.sum(got[k] != want[k][index]) <= 5
    assert 0 < got[1].sum() < n and got[2].max() > 0 and got[3].max() > 0

    want, got = plain(0.05, 1), decomposed(0.05, 1)
    assert len(got[0]) == len(want[0]) > n
    assert np.array_equal(np.sort(got[1]), np.sort(want[1]))
    a = got[0][np.lexsort(got[0][:, :3].T[::-1])]
    b = want[0][np.lexsort(want[0][:, :3].T[::-1])]
    assert np.max(np.abs(a - b)) < 1e-4
    # and it keeps growing, every cell finite, counters alive
    got = decomposed(0.05, 8)
    assert len(got[0]) > len(want[0]) and np.all(np.isfinite(got[0]))
    assert got[2].max() > 0


@pytest.mark.gpu
def test_branching_growth_through_the_decomposed_step(product):
    """configs[3] (branching cell, division, one protrusion per cell rewired
    every step) as a single brick: links are kept as cell identities
    (b200/brick_links.cuh), resolved to indices for the rewiring kernel and for
    link_forces, and travel with the cells through the migration pass. The first
    rewiring must produce exactly the links of the plain model; afterwards the
    run keeps every link resolvable and grows at the plain model's rate."""
    rng = np.random.default_rng(45)
    n = 30_000
    X = np.zeros((n, 7), dtype=np.float32)
    X[:, :5] = workloads.polarized_ball(n, 0.75, rng, lattice=True, noise=0.0)
    X[:, 5:] = rng.random((n, 2)).astype(np.float32) * 0.2
    types = workloads.shell_types(X)
    X[types == 0, 3:5] = 0
    gs = workloads.grid_size_for(n, 0.75, growth=2.0)
    cuts = dd.ball_brick_cuts(float(np.max(np.linalg.norm(X[:, :3], axis=1))),
                              (1, 1, 1))

    def plain(rate, steps):
        with product.sim("branching_growth", 2 * n, gs, 1.0) as sim:
            for name, value in (("seed", 9), ("mes_rate", rate), ("epi_rate", rate)):
                sim.set_param(name, value)
            sim.set_ints("type", types)
            sim.set_state(X)
            sim.step(0.1, steps)
            return sim.get_state(), sim.get_links()

    def decomposed(rate, steps):
        domain = dd.BrickDomain(product, "branching_growth", 2 * n, gs, 1.0,
                                (1, 1, 1), cuts, 0, 1, face_capacity=n, halo=2.5)
        dd.connect_local([domain])
        for name, value in (("seed", 9), ("mes_rate", rate), ("epi_rate", rate)):
            domain.sim.set_param(name, value)
        domain.set_cells(X)
        domain.sim.set_ints("type", types)
        domain.step(0.1, steps)
        owned, _, problems = domain.counts()
        assert problems == 0
        out = (domain.owned_state()[0].cpu().numpy(), domain.sim.get_ints("identity"),
               domain.sim.get_ints("partner"),
               int(domain.sim.get_ints("unresolved_links")[0]))
        assert len(out[0]) == owned == len(out[1]) == len(out[2])
        domain.close()
        return out

    want_X, want_links = plain(0.0, 1)
    got_X, identity, partner, unresolved = decomposed(0.0, 1)
    assert unresolved == 0 and np.array_equal(np.sort(identity), np.arange(n))
    live = want_links[:, 0] != want_links[:, 1]
    assert live.sum() > 0.02 * (types == 0).sum()
    # the plain model's link of cell a sits at index a
    assert np.array_equal(want_links[live, 0], np.nonzero(live)[0])
    expected = np.where(live, want_links[:n, 1], np.arange(n))
    assert np.array_equal(partner, expected[identity])
    assert np.max(np.abs(got_X - want_X[identity])) < 1e-4

    want_X, _ = plain(0.01, 12)
    got_X, identity, partner, unresolved = decomposed(0.01, 12)
    assert unresolved == 0
    assert len(np.unique(identity)) == len(identity) > n, (
        len(np.unique(identity)), len(identity))
    assert np.all(np.isfinite(got_X))
    assert abs(len(got_X) - len(want_X)) < 0.02 * len(want_X)
    linked = np.mean(partner != identity)
    assert linked > 0.02
    # both ends of every link are cells of the tissue, a protrusion's length apart
    where = {int(g): k for k, g in enumerate(identity)}
    ends = np.array([where[int(p)] for p in partner])
    length = np.linalg.norm(got_X[:, :3] - got_X[ends, :3], axis=1)
    assert length.max() < 3.0


@pytest.mark.gpu
def test_registering_arrays_reports_misuse_and_empty_bricks_work(product):
    """Edge cases of the travelling arrays: widths that are not whole words, too
    many arrays, registration after the domain began -- all refused with an
    error; a brick that owns no cell at all (the tissue lies entirely in its
    neighbour) still exchanges, and a tissue that drifts into it arrives with
    its arrays."""
    import torch
    import yalla_b200 as yb
    n, gs = 4000, 40
    tag = torch.zeros(n, dtype=torch.int32, device="cuda")
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        with pytest.raises(yb.YallaError, match="multiple of 4"):
            sim.dom_register_array(tag.data_ptr(), 6)
        for _ in range(8):
            sim.dom_register_array(tag.data_ptr(), 4)
        with pytest.raises(yb.YallaError, match="at most 8"):
            sim.dom_register_array(tag.data_ptr(), 4)
    with product.sim("relu_tile", 64) as sim:
        with pytest.raises(yb.YallaError, match="Grid model"):
            sim.dom_begin(0, 1, [-np.inf] * 3, [np.inf] * 3, 1.5,
                          np.full(27, -1, np.int32), np.zeros(27, np.int32),
                          [0, 0, 0], [0, 0, 0])

    # two bricks cut at z = 6: all cells start below the cut
    rng = np.random.default_rng(46)
    X = workloads.lattice_ball(n, 0.8, rng).astype(np.float32) * 0.9
    X[:, 2] += 6.0 - X[:, 2].max() - 0.05
    cuts = [[], [], [6.0]]
    tags = [torch.full((n,), -1, dtype=torch.int32, device="cuda") for _ in range(2)]
    domains = [dd.BrickDomain(product, "relu_grid", n, gs, 1.0, (1, 1, 2), cuts, rank,
                              2, face_capacity=n,
                              arrays=[(tags[rank].data_ptr(), 4, True)])
               for rank in range(2)]
    with pytest.raises(yb.YallaError, match="before yb_dom_begin"):
        domains[0].sim.dom_register_array(tag.data_ptr(), 4)
    streams = [torch.cuda.Stream() for _ in domains]
    for domain, stream in zip(domains, streams):
        domain.sim.set_stream(stream.cuda_stream)
    dd.connect_local(domains)
    owned = [domain.owns(X) for domain in domains]
    assert owned[0].all() and not owned[1].any()
    for rank, domain in enumerate(domains):
        domain.set_cells(X[owned[rank]])
    tags[0][:n] = torch.arange(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for _ in range(10):  # the squeezed ball expands across the cut
        for domain in domains:
            domain.step(0.1)
    torch.cuda.synchronize()
    counts = [domain.counts() for domain in domains]
    assert all(c[2] == 0 for c in counts)
    assert counts[0][0] + counts[1][0] == n and counts[1][0] > 0
    with product.sim("relu_grid", n, gs, 1.0) as sim:
        sim.set_state(X)
        sim.step(0.1, 10)
        want = sim.get_state()
    for rank, domain in enumerate(domains):
        got = domain.owned_state()[0].cpu().numpy()
        ids = tags[rank][:counts[rank][0]].cpu().numpy()
        assert ids.min() >= 0
        assert np.max(np.abs(got - want[ids])) < 1e-3
    for domain in domains:
        domain.close()
