"""Slab domain decomposition of one tissue across the GPUs of a node.

ya||a is single-GPU; this is the extension BASELINE.json's
north_star asks for. The tissue is cut into slabs along z, one process per GPU
(torch.distributed: NCCL on GPUs, gloo in the CPU tests), each driving one
solver through the ``yb_dd_*`` building blocks of include/yalla_b200.h. Because
interactions are strictly shorter than cube_size, a slab needs a halo of one
cube from each neighbour; per Heun stage:

    yb_slab_pack (boundary cells -> exchange buffers, on the device)
    -> send/recv whole buffers with the two neighbours
    -> yb_slab_unpack (ghosts behind the owned cells)
    -> yb_dd_forces (grid build + sweep; ghosts are partners only)
    -> all-reduce {sum dX, n} (the drift is the GLOBAL mean force,
       solvers.cuh:241-255) -> yb_slab_update (predictor / corrector)

and once per step cells that crossed a cut migrate to the neighbour the same
way. torch is used for the plumbing only (buffers, send/recv, all-reduce);
packing, grid build, sweep and updates are the library's kernels, and because
all counts stay on the device a step never synchronises with the host. The
module is device-agnostic: with the CPU oracle and gloo the very same code runs
in the CPU test-suite.

Limits (documented in DESIGN.md): Grid models without per-id property arrays
("relu_grid", "spring_grid", "epithelium"); cell identity is not tracked across
migration; centre-of-mass fixing only.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def ball_slab_cuts(radius, n_slabs, cube_size=1.0):
    """z positions of the n_slabs - 1 cuts that split a ball of the given
    radius into slabs of equal volume, snapped to cube boundaries."""
    z = np.linspace(-radius, radius, 20001)
    below = (2 * radius ** 3 / 3 + radius ** 2 * z - z ** 3 / 3) / (4 * radius ** 3 / 3)
    cuts = np.interp(np.arange(1, n_slabs) / n_slabs, below, z)
    return [float(np.round(c / cube_size) * cube_size) for c in cuts]


def lattice_ball_slab(radius, dist_to_nb, z_lo, z_hi, rng, jitter=0.05):
    """The cells of workloads.lattice_ball's jittered FCC ball that fall into
    z_lo <= z < z_hi, generated without building the whole ball."""
    d = float(dist_to_nb)
    a = d * np.sqrt(2.0)
    half = int(np.ceil(radius / a)) + 1
    k_lo = max(-half, int(np.floor(max(z_lo, -radius - a) / a)) - 1)
    k_hi = min(half, int(np.ceil(min(z_hi, radius + a) / a)) + 1)
    axis = np.arange(-half, half + 1, dtype=np.float64) * a
    z_axis = np.arange(k_lo, k_hi + 1, dtype=np.float64) * a
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    chunks = []
    for zc in z_axis:  # layer by layer keeps the temporary arrays small
        gx, gy = np.meshgrid(axis, axis, indexing="ij")
        corner = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, zc)], axis=1)
        pts = (corner[:, None, :] + basis[None, :, :]).reshape(-1, 3)
        keep = (np.einsum("ij,ij->i", pts, pts) <= radius * radius)
        keep &= (pts[:, 2] >= z_lo) & (pts[:, 2] < z_hi)
        chunks.append(pts[keep])
    points = np.concatenate(chunks) if chunks else np.zeros((0, 3))
    points += (rng.random(points.shape) - 0.5) * 2.0 * jitter * d
    # the jitter may push a cell across a cut; migration would fix it, but
    # start clean: clip into the slab
    eps = 1e-4
    if np.isfinite(z_lo):
        points[:, 2] = np.maximum(points[:, 2], z_lo + eps)
    if np.isfinite(z_hi):
        points[:, 2] = np.minimum(points[:, 2], z_hi - eps)
    rng.shuffle(points, axis=0)
    return points.astype(np.float32)


class SlabDomain:
    """One rank's slab: owns the cells with z_lo <= z < z_hi.

    State and all cell counts live inside the library; a step posts kernels,
    fixed-size neighbour exchanges and two 16-byte all-reduces and never waits
    for the device. `n_owned`, `counts()` and `gather_all()` do block and are
    meant for set-up, diagnostics and tests.
    """

    def __init__(self, lib, model, n_max, grid_size, cube_size, z_lo, z_hi,
                 device, halo=1.5, halo_capacity=None, group=None,
                 local_grid=True):
        self.lib = lib
        self.sim = lib.sim(model, n_max, grid_size, cube_size)
        self.lanes = self.sim.lanes
        self.width = self.lanes + 3  # floats per exchanged record
        self.n_max = n_max
        self.cube_size = float(cube_size)
        self.z_lo, self.z_hi = float(z_lo), float(z_hi)
        self.halo = halo * self.cube_size
        self.device = torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.lower = self.rank - 1 if self.rank > 0 else None
        self.upper = self.rank + 1 if self.rank < self.world - 1 else None
        self.capacity = int(halo_capacity or max(1024, n_max // 4))
        size = 4 + self.capacity * self.width
        self.send = [torch.zeros(size, dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.recv = [torch.zeros(size, dtype=torch.float32, device=self.device)
                     for _ in range(2)]
        self.sums = torch.zeros(4, dtype=torch.float32, device=self.device)
        # the solver's grid only needs the z layers this slab can touch: its own
        # range plus the halo and one cube of slack for cells on the move
        first_layer, n_layers = 0, 0
        if local_grid and self.world > 1:
            half = grid_size // 2
            lo = -half if not np.isfinite(z_lo) else int(
                np.floor((z_lo - self.halo) / cube_size)) - 2
            hi = half if not np.isfinite(z_hi) else int(
                np.ceil((z_hi + self.halo) / cube_size)) + 2
            lo, hi = max(lo, -half), min(hi, half)
            first_layer, n_layers = lo + half, hi - lo
        self.sim.slab_begin(self.z_lo, self.z_hi, self.halo, self.capacity,
                            first_layer, n_layers)

    def close(self):
        self.sim.close()

    # ---- state ---------------------------------------------------------------
    def set_cells(self, X, v=None):
        X = torch.as_tensor(X, dtype=torch.float32).reshape(-1, self.lanes)
        X = X.to(self.device).contiguous()
        if v is None:
            v = torch.zeros((len(X), 3), dtype=torch.float32, device=self.device)
        else:
            v = torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous()
        self.sim.slab_set_owned(X.data_ptr(), v.data_ptr(), len(X))
        self._sync()

    def _sync(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)

    def counts(self):
        """(owned, owned + ghosts, problems) -- blocks."""
        return self.sim.slab_counts()

    @property
    def n_owned(self):
        return self.counts()[0]

    def total_cells(self):
        count = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(count, group=self.group)
        return int(count.item())

    def owned_state(self):
        """(X, v) of the owned cells as new tensors -- blocks."""
        n = self.n_owned
        X = torch.empty((n, self.lanes), dtype=torch.float32, device=self.device)
        v = torch.empty((n, 3), dtype=torch.float32, device=self.device)
        self.sim.dd_read(0, X.data_ptr(), n)
        self.sim.dd_read(2, v.data_ptr(), n)
        self._sync()
        return X, v

    # ---- neighbour exchange ------------------------------------------------------
    def _exchange(self):
        """Ship send[0] to the lower and send[1] to the upper neighbour, receive
        their buffers into recv[0] / recv[1]. Whole buffers: no counts needed."""
        ops = []
        if self.lower is not None:
            ops.append(dist.P2POp(dist.isend, self.send[0], self.lower, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv[0], self.lower, self.group))
        if self.upper is not None:
            ops.append(dist.P2POp(dist.isend, self.send[1], self.upper, self.group))
            ops.append(dist.P2POp(dist.irecv, self.recv[1], self.upper, self.group))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()

    def _round(self, what):
        self.sim.slab_pack(what, self.send[0].data_ptr(), self.send[1].data_ptr())
        self._exchange()
        self.sim.slab_unpack(what, self.recv[0].data_ptr(), self.recv[1].data_ptr())

    # ---- one Heun step ---------------------------------------------------------------
    def step(self, dt):
        for stage in (0, 1):
            self._round(stage)  # halo of X resp. X1
            self.sim.dd_forces(stage, self.sums.data_ptr())
            if self.world > 1:
                dist.all_reduce(self.sums, group=self.group)
            self.sim.slab_update(stage, dt, self.sums.data_ptr())
        self._round(2)  # migration

    def gather_all(self):
        """All cells of the tissue on every rank (tests and small runs only)."""
        X, _ = self.owned_state()
        if self.world == 1:
            return X.cpu().numpy()
        counts = torch.tensor([len(X)], dtype=torch.int64, device=self.device)
        every = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(every, counts, group=self.group)
        every = every.tolist()
        pad = max(every)
        mine = torch.zeros((pad, self.lanes), dtype=torch.float32, device=self.device)
        mine[:len(X)] = X
        parts = torch.empty((self.world * pad, self.lanes), dtype=torch.float32,
                            device=self.device)
        dist.all_gather_into_tensor(parts, mine, group=self.group)
        parts = parts.view(self.world, pad, self.lanes)
        return torch.cat([parts[r, :every[r]] for r in range(self.world)]).cpu().numpy()


# ---- brick decomposition over peer memory (include/b200/domain.cuh) -----------
def brick_grid_for(world):
    """(bx, by, bz) bricks for `world` GPUs: slabs for 2, 2 x 2 x 1 for 4,
    2 x 2 x 2 for 8; otherwise the most cubic factorisation."""
    best = None
    for bz in range(1, world + 1):
        if world % bz:
            continue
        for by in range(1, world // bz + 1):
            if (world // bz) % by:
                continue
            bx = world // (bz * by)
            if bx <= by <= bz:
                shape = (bx, by, bz)
                if best is None or max(shape) < max(best):
                    best = shape
    return best


def brick_coord(rank, bricks):
    bx, by, _ = bricks
    return rank % bx, (rank // bx) % by, rank // (bx * by)


def brick_rank(coord, bricks):
    return coord[0] + bricks[0] * (coord[1] + bricks[1] * coord[2])


def ball_brick_cuts(radius, bricks, cube_size=1.0):
    """Per axis, the cut positions that split a ball into equal-volume slabs
    along that axis (the bricks are their tensor product)."""
    return [ball_slab_cuts(radius, n, cube_size) for n in bricks]


class BrickDomain:
    """One rank's brick. Set-up is host work (layout, CUDA IPC handles gathered
    with torch.distributed or plain pointers inside one process); after
    connect() a step is one call into the library, which posts kernels only:
    halo exchange, migration and the drift sum are stores into the neighbours'
    memory over NVLink, ordered by flags the kernels themselves wait on."""

    def __init__(self, lib, model, n_max, grid_size, cube_size, bricks, cuts,
                 rank, world, face_capacity, halo=1.5, local_grid=True,
                 arrays=()):
        """arrays: (device address, bytes per cell, ghosts_too) of per-cell
        arrays of the caller that travel with the cells (Sim.dom_register_array);
        the typed models register their Property arrays themselves."""
        assert world == bricks[0] * bricks[1] * bricks[2]
        self.lib = lib
        self.sim = lib.sim(model, n_max, grid_size, cube_size)
        for address, width, ghosts_too in arrays:
            self.sim.dom_register_array(address, width, ghosts_too)
        self.lanes = self.sim.lanes
        self.n_max = n_max
        self.rank, self.world = rank, world
        self.bricks = tuple(bricks)
        self.coord = brick_coord(rank, bricks)
        self.halo = halo * cube_size
        bounds = [[-np.inf] + list(cuts[a]) + [np.inf] for a in range(3)]
        self.lo = np.array([bounds[a][self.coord[a]] for a in range(3)], np.float32)
        self.hi = np.array([bounds[a][self.coord[a] + 1] for a in range(3)],
                           np.float32)
        self.peer_ranks = np.full(27, -1, dtype=np.int32)
        self.capacity = np.zeros(27, dtype=np.int32)
        for d in range(27):
            if d == 13:
                continue
            step = (d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1)
            there = tuple(self.coord[a] + step[a] for a in range(3))
            if all(0 <= there[a] < bricks[a] for a in range(3)):
                self.peer_ranks[d] = brick_rank(there, bricks)
                axes = sum(1 for s in step if s != 0)  # 1 face, 2 edge, 3 corner
                self.capacity[d] = (face_capacity if axes == 1 else
                                    face_capacity // 16 + 2048 if axes == 2 else
                                    face_capacity // 256 + 2048)
        # the cubes this brick can touch: its own range plus the halo and two
        # cubes of slack for cells on the move
        first, count = [0, 0, 0], [0, 0, 0]
        if local_grid and world > 1:
            half = grid_size // 2
            for a in range(3):
                lo = -half if not np.isfinite(self.lo[a]) else int(
                    np.floor((self.lo[a] - self.halo) / cube_size)) - 2
                hi = half if not np.isfinite(self.hi[a]) else int(
                    np.ceil((self.hi[a] + self.halo) / cube_size)) + 2
                lo, hi = max(lo, -half), min(hi, half)
                first[a], count[a] = lo + half, hi - lo
        self.sim.dom_begin(rank, world, self.lo, self.hi, self.halo,
                           self.peer_ranks, self.capacity, first, count)
        self.mapped = []

    def close(self):
        self.sim.sync()
        for base in self.mapped:
            self.lib.ipc_release(base)
        self.mapped = []
        self.sim.close()

    def connect(self, bases, offsets):
        """bases[r]: rank r's exchange allocation as an address valid in this
        process; offsets[r]: its [27, 8] table (Sim.dom_exchange)."""
        for d in range(27):
            r = int(self.peer_ranks[d])
            if r >= 0:
                self.sim.dom_connect(d, bases[r], offsets[r][26 - d])
        for r in range(self.world):
            self.sim.dom_connect_mailbox(r, bases[r])

    def connect_over_ipc(self, group=None):
        """Gather everybody's CUDA IPC handle and offsets (torch.distributed is
        the plumbing), map the other ranks' allocations, connect. Raises
        YallaError on EVERY rank if any rank cannot map a peer (no peer access,
        IPC disabled in the container, ...), so that callers can fall back
        together."""
        from . import YallaError
        base, _, offsets = self.sim.dom_exchange()
        if self.world == 1:
            return self.connect([base], [offsets])
        mine = (self.lib.ipc_export(base), offsets)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        bases, problem = [], None
        if os.environ.get("YALLA_B200_NO_IPC"):  # exercise the callers' fallback
            problem = "disabled by YALLA_B200_NO_IPC"
        for r, (handle, _) in enumerate(everyone):
            if r == self.rank or problem:
                bases.append(base)
                continue
            try:
                bases.append(self.lib.ipc_import(handle))
                self.mapped.append(bases[-1])
            except YallaError as error:
                problem = str(error)
                break
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        fine = torch.tensor([0 if problem else 1], dtype=torch.int32, device=device)
        dist.all_reduce(fine, op=dist.ReduceOp.MIN, group=group)
        if int(fine.item()) == 0:
            for mapped in self.mapped:
                self.lib.ipc_release(mapped)
            self.mapped = []
            raise YallaError("CUDA IPC between the ranks is not available: "
                             + (problem or "another rank could not map a peer"))
        self.connect(bases, [table for _, table in everyone])
        dist.barrier(group=group)

    # ---- state -------------------------------------------------------------------
    def owns(self, X):
        inside = np.ones(len(X), dtype=bool)
        for a in range(3):
            inside &= (X[:, a] >= self.lo[a]) & (X[:, a] < self.hi[a])
        return inside

    def set_cells(self, X, v=None, device="cuda"):
        X = torch.as_tensor(np.ascontiguousarray(X, np.float32)).reshape(
            -1, self.lanes).to(device)
        if v is None:
            v = torch.zeros((len(X), 3), dtype=torch.float32, device=device)
        else:
            v = torch.as_tensor(np.ascontiguousarray(v, np.float32)).to(device)
        self.sim.slab_set_owned(X.data_ptr(), v.data_ptr(), len(X))
        torch.cuda.synchronize()

    def seed_lattice_ball(self, radius, dist_to_nb, seed, jitter=0.05):
        return self.sim.dom_seed_lattice_ball(radius, dist_to_nb, jitter, seed)

    def counts(self):
        """(owned, owned + ghosts, problems) -- blocks."""
        return self.sim.slab_counts()

    def owned_state(self, device="cuda"):
        n = self.counts()[0]
        X = torch.empty((n, self.lanes), dtype=torch.float32, device=device)
        v = torch.empty((n, 3), dtype=torch.float32, device=device)
        self.sim.dd_read(0, X.data_ptr(), n)
        self.sim.dd_read(2, v.data_ptr(), n)
        torch.cuda.synchronize()
        return X, v

    def step(self, dt, n_steps=1):
        self.sim.dom_step(dt, n_steps)


def connect_local(domains):
    """All bricks live in this process (tests, one GPU): plain device pointers."""
    info = [d.sim.dom_exchange() for d in domains]
    bases = [base for base, _, _ in info]
    offsets = [table for _, _, table in info]
    for d in domains:
        d.connect(bases, offsets)
