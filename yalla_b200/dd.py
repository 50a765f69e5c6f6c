"""Slab domain decomposition of one tissue across the GPUs of a node.

ya||a is single-GPU; this is the extension BASELINE.json's
north_star asks for. The tissue is cut into slabs along z, one process per GPU
(torch.distributed: NCCL on GPUs, gloo in the CPU tests), each driving one
solver through the ``yb_dd_*`` building blocks of include/yalla_b200.h. Because
interactions are strictly shorter than cube_size, a slab needs a halo of one
cube from each neighbour; per Heun stage:

    pack boundary cells -> exchange with the two neighbours (send/recv)
    -> yb_dd_load (owned + ghosts) -> yb_dd_forces (grid build + sweep)
    -> all-reduce {sum dX, n} (the drift is the GLOBAL mean force,
       solvers.cuh:241-255) -> yb_dd_update (predictor / corrector)

and once per step cells that crossed a cut migrate to the neighbour. torch is
used for the plumbing only (masks, gathers, send/recv, all-reduce); grid build,
sweep and updates are the library's kernels. The module is device-agnostic: with
the CPU oracle and gloo the very same code runs in the CPU test-suite.

Limits (documented in DESIGN.md): Grid models without per-id property arrays
("relu_grid", "spring_grid", "epithelium"); cell identity is not tracked across
migration; centre-of-mass fixing only.
"""
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
    """One rank's slab: owns the cells with z_lo <= z < z_hi."""

    def __init__(self, lib, model, n_max, grid_size, cube_size, z_lo, z_hi,
                 device, halo=1.5, group=None):
        self.lib = lib
        self.sim = lib.sim(model, n_max, grid_size, cube_size)
        self.lanes = self.sim.lanes
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
        self.X = torch.zeros((0, self.lanes), dtype=torch.float32, device=self.device)
        self.v = torch.zeros((0, 3), dtype=torch.float32, device=self.device)
        self.sums = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.stats = {"ghosts": 0, "migrated": 0}

    def close(self):
        self.sim.close()

    # ---- state ---------------------------------------------------------------
    def set_cells(self, X, v=None):
        X = torch.as_tensor(X, dtype=torch.float32).reshape(-1, self.lanes)
        self.X = X.to(self.device).contiguous()
        if v is None:
            self.v = torch.zeros((len(self.X), 3), dtype=torch.float32,
                                 device=self.device)
        else:
            self.v = torch.as_tensor(v, dtype=torch.float32).to(self.device).contiguous()

    @property
    def n_owned(self):
        return int(self.X.shape[0])

    def total_cells(self):
        count = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(count, group=self.group)
        return int(count.item())

    # ---- neighbour exchange ------------------------------------------------------
    def _exchange(self, to_lower, to_upper):
        """Send one [m, width] tensor to each existing neighbour, receive
        theirs; returns (from_lower, from_upper), empty where there is none."""
        width = to_lower.shape[1]
        empty = torch.zeros((0, width), dtype=torch.float32, device=self.device)
        if self.world == 1:
            return empty, empty
        counts = torch.tensor([to_lower.shape[0], to_upper.shape[0]],
                              dtype=torch.int64, device=self.device)
        gathered = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(gathered, counts, group=self.group)
        gathered = gathered.view(self.world, 2).tolist()
        from_lower, from_upper = empty, empty
        ops = []
        if self.lower is not None:
            incoming = int(gathered[self.lower][1])  # what it sends upwards
            from_lower = torch.empty((incoming, width), dtype=torch.float32,
                                     device=self.device)
            if to_lower.shape[0] > 0:
                ops.append(dist.P2POp(dist.isend, to_lower.contiguous(), self.lower,
                                      self.group))
            if incoming > 0:
                ops.append(dist.P2POp(dist.irecv, from_lower, self.lower, self.group))
        if self.upper is not None:
            incoming = int(gathered[self.upper][0])  # what it sends downwards
            from_upper = torch.empty((incoming, width), dtype=torch.float32,
                                     device=self.device)
            if to_upper.shape[0] > 0:
                ops.append(dist.P2POp(dist.isend, to_upper.contiguous(), self.upper,
                                      self.group))
            if incoming > 0:
                ops.append(dist.P2POp(dist.irecv, from_upper, self.upper, self.group))
        if ops:
            for work in dist.batch_isend_irecv(ops):
                work.wait()
        return from_lower, from_upper

    def _halo(self, X):
        """Ghost cells for the positions X of the owned cells: the neighbours'
        cells within `halo` of the shared cut (positions and old velocities)."""
        payload = torch.cat([X, self.v], dim=1)
        z = X[:, 2]
        none = payload[:0]
        to_lower = payload[z < self.z_lo + self.halo] if self.lower is not None else none
        to_upper = payload[z >= self.z_hi - self.halo] if self.upper is not None else none
        from_lower, from_upper = self._exchange(to_lower, to_upper)
        ghosts = torch.cat([from_lower, from_upper], dim=0)
        return ghosts[:, :self.lanes].contiguous(), ghosts[:, self.lanes:].contiguous()

    # ---- one Heun step ---------------------------------------------------------------
    def _stage(self, stage, X_stage, dt):
        gX, gv = self._halo(X_stage)
        n, n_ghost = self.n_owned, int(gX.shape[0])
        if n + n_ghost > self.n_max:
            raise RuntimeError(f"rank {self.rank}: {n} owned + {n_ghost} ghost "
                               f"cells exceed n_max = {self.n_max}")
        self.stats["ghosts"] = n_ghost
        if stage == 0:
            self.sim.dd_load(0, self.X.data_ptr(), self.v.data_ptr(), n,
                             gX.data_ptr(), gv.data_ptr(), n_ghost)
        else:
            self.sim.dd_load(1, 0, 0, n, gX.data_ptr(), gv.data_ptr(), n_ghost)
        self.sim.dd_forces(stage, self.sums.data_ptr())
        if self.world > 1:
            dist.all_reduce(self.sums, group=self.group)
        mean = (self.sums[:3] / self.sums[3]).contiguous()
        self.sim.dd_update(stage, dt, mean.data_ptr())
        # gX, gv and mean must outlive the asynchronous copies that read them
        self._keep = (gX, gv, mean)

    def step(self, dt):
        n = self.n_owned
        self._stage(0, self.X, dt)
        X1 = torch.empty_like(self.X)
        self.sim.dd_read(1, X1.data_ptr(), n)
        self._stage(1, X1, dt)
        self.sim.dd_read(0, self.X.data_ptr(), n)
        self.sim.dd_read(2, self.v.data_ptr(), n)
        self._migrate()

    def _migrate(self):
        """Hand cells that crossed a cut to the neighbouring slab."""
        if self.world == 1:
            return
        payload = torch.cat([self.X, self.v], dim=1)
        z = self.X[:, 2]
        down = (z < self.z_lo) if self.lower is not None else torch.zeros_like(z, dtype=torch.bool)
        up = (z >= self.z_hi) if self.upper is not None else torch.zeros_like(z, dtype=torch.bool)
        from_lower, from_upper = self._exchange(payload[down], payload[up])
        stay = payload[~(down | up)]
        merged = torch.cat([stay, from_lower, from_upper], dim=0)
        self.stats["migrated"] = int(down.sum().item() + up.sum().item())
        self.X = merged[:, :self.lanes].contiguous()
        self.v = merged[:, self.lanes:].contiguous()

    def gather_all(self):
        """All cells of the tissue on every rank (tests and small runs only)."""
        if self.world == 1:
            return self.X.cpu().numpy()
        counts = torch.tensor([self.n_owned], dtype=torch.int64, device=self.device)
        every = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(every, counts, group=self.group)
        every = every.tolist()
        pad = max(every)
        mine = torch.zeros((pad, self.lanes), dtype=torch.float32, device=self.device)
        mine[:self.n_owned] = self.X
        parts = torch.empty((self.world * pad, self.lanes), dtype=torch.float32,
                            device=self.device)
        dist.all_gather_into_tensor(parts, mine, group=self.group)
        parts = parts.view(self.world, pad, self.lanes)
        return torch.cat([parts[r, :every[r]] for r in range(self.world)]).cpu().numpy()
