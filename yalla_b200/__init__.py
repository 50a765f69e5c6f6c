"""Python host side of the ya||a B200 hot path.

The product is the CUDA C++ header set in ``include/`` plus the compiled C ABI
``include/yalla_b200.h``. This package is the thin ctypes binding above that
ABI which the test-suite and ``bench.py`` use. The same binding drives three
interchangeable libraries (see ``yalla_b200.h``):

* ``product()``   -- ``yalla_b200/_lib/libyalla_b200.so`` (this repo's kernels)
* ``reference()`` -- ``oracle/_ref/libyalla_ref.so`` (unmodified reference
  headers compiled for sm_100a; the baseline arm and the A/B parity vote)
* the CPU oracle is loaded by ``tests/`` and ``bench.py`` only, via
  ``yalla_b200.load(path)``; nothing in this package imports ``oracle/``.

There is no CPU fallback: ``product()`` raises if the CUDA library is missing.
"""
import ctypes
import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_LIB = os.environ.get("YALLA_B200_LIB") or os.path.join(
    _ROOT, "yalla_b200", "_lib", "libyalla_b200.so")
REFERENCE_LIB = os.path.join(_ROOT, "oracle", "_ref", "libyalla_ref.so")
REFERENCE_NDEBUG_LIB = os.path.join(_ROOT, "oracle", "_ref",
                                    "libyalla_ref_ndebug.so")

YB_OK, YB_EINVAL, YB_ECUDA, YB_ENOSYS = 0, -1, -2, -3

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_float_p = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); mirrors include/yalla_b200.h one to one
DOM_OFFSETS = 8  # YB_DOM_OFFSETS of include/yalla_b200.h

SIGNATURES = {
    "yb_build_info": (ctypes.c_char_p, []),
    "yb_last_error": (ctypes.c_char_p, []),
    "yb_sim_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_float,
                                     ctypes.POINTER(ctypes.c_void_p)]),
    "yb_sim_destroy": (None, [ctypes.c_void_p]),
    "yb_sim_lanes": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_sim_n_max": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_sim_set_param": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p,
                                        ctypes.c_double]),
    "yb_sim_set_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_int]),
    "yb_sim_get_state": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, _c_int_p]),
    "yb_sim_get_velocities": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int]),
    "yb_sim_set_velocities": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int]),
    "yb_sim_host_drain": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_sim_set_ints": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p,
                                       ctypes.c_void_p, ctypes.c_int]),
    "yb_sim_get_ints": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p,
                                       ctypes.c_void_p, ctypes.c_int]),
    "yb_sim_seed_sphere": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_ulonglong,
                                          ctypes.c_int]),
    "yb_sim_set_links": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int]),
    "yb_sim_get_links": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, _c_int_p]),
    "yb_sim_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float,
                                   ctypes.c_int]),
    "yb_sim_step_timed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float,
                                         ctypes.c_int, _c_float_p,
                                         ctypes.POINTER(ctypes.c_longlong)]),
    "yb_sim_step_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_float,
                                        ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_int, _c_int_p]),
    "yb_dd_load": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_int]),
    "yb_dd_forces": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_void_p]),
    "yb_dd_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                    ctypes.c_void_p]),
    "yb_dd_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_int]),
    "yb_slab_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float,
                                     ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int]),
    "yb_slab_set_owned": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int]),
    "yb_slab_pack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "yb_slab_unpack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "yb_slab_update": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_void_p]),
    "yb_slab_counts": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, _c_int_p,
                                      _c_int_p]),
    "yb_dom_begin": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_float, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]),
    "yb_dom_register_array": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_int]),
    "yb_dom_exchange": (ctypes.c_int, [ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_void_p),
                                       ctypes.POINTER(ctypes.c_longlong),
                                       ctypes.c_void_p]),
    "yb_dom_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "yb_dom_connect_mailbox": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int,
                                              ctypes.c_void_p]),
    "yb_dom_seed_lattice_ball": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float,
                                                ctypes.c_float, ctypes.c_float,
                                                ctypes.c_ulonglong, _c_int_p]),
    "yb_dom_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_int]),
    "yb_dom_read_profile": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_ipc_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_ipc_import": (ctypes.c_int, [ctypes.c_void_p,
                                     ctypes.POINTER(ctypes.c_void_p)]),
    "yb_ipc_release": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_sim_set_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "yb_sim_step_host_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_float,
                                              ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_void_p]),
    "yb_sim_profile_sweeps": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "yb_sim_read_sweep_profile": (ctypes.c_int, [ctypes.c_void_p, _c_float_p,
                                                 _c_int_p]),
    "yb_sim_n": (ctypes.c_int, [ctypes.c_void_p, _c_int_p]),
    "yb_sim_sync": (ctypes.c_int, [ctypes.c_void_p]),
    "yb_grid_build": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "yb_nhood": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "yb_link_forces": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_float]),
    "yb_bending_force": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_int, ctypes.c_void_p]),
    "yb_polarization_force": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int, ctypes.c_void_p]),
}

MODEL_LANES = {
    "springs": 3, "spring_tile": 3, "spring_grid": 3, "relu_tile": 3,
    "relu_grid": 3, "relu_gabriel": 3, "protrusions": 3, "epithelium": 5, "growth": 5,
    "branching": 7, "branching_growth": 7,
}
# models without a CPU oracle (curand; the Gabriel solver)
GPU_ONLY_MODELS = {"branching_growth", "relu_gabriel"}


class YallaError(RuntimeError):
    pass


def _f32(array):
    return np.ascontiguousarray(array, dtype=np.float32)


def _i32(array):
    return np.ascontiguousarray(array, dtype=np.int32)


class Library:
    """One loaded implementation of the C ABI."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise YallaError(
                f"{path} is missing -- build it first "
                "(python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(self.cdll, name)  # AttributeError if not exported
            fn.restype = restype
            fn.argtypes = argtypes

    @property
    def build_info(self):
        return self.cdll.yb_build_info().decode()

    def check(self, status, what):
        if status != YB_OK:
            raise YallaError(
                f"{what} failed ({status}): "
                f"{self.cdll.yb_last_error().decode()} [{self.build_info}]")

    def sim(self, model, n_max, grid_size=50, cube_size=1.0):
        return Sim(self, model, n_max, grid_size, cube_size)

    # ---- stateless entry points -------------------------------------------
    def nhood(self, grid_size):
        out = np.zeros(27, dtype=np.int32)
        self.check(self.cdll.yb_nhood(grid_size, out.ctypes.data), "yb_nhood")
        return out

    def grid_build(self, d_X, n, lanes, grid_size, cube_size, d_cube_id,
                   d_point_id, d_cube_start, d_cube_end):
        """All array arguments are device pointers (ints)."""
        self.check(self.cdll.yb_grid_build(
            d_X, lanes, n, grid_size, cube_size, d_cube_id, d_point_id,
            d_cube_start, d_cube_end), "yb_grid_build")

    def ipc_export(self, d_base):
        handle = (ctypes.c_ubyte * 64)()
        self.check(self.cdll.yb_ipc_export(d_base, handle), "yb_ipc_export")
        return bytes(handle)

    def ipc_import(self, handle_bytes):
        handle = (ctypes.c_ubyte * 64).from_buffer_copy(handle_bytes)
        base = ctypes.c_void_p()
        self.check(self.cdll.yb_ipc_import(handle, ctypes.byref(base)),
                   "yb_ipc_import")
        return base.value

    def ipc_release(self, d_base):
        self.check(self.cdll.yb_ipc_release(d_base), "yb_ipc_release")

    def link_forces(self, d_X, d_dX, lanes, n, d_links, n_links, strength):
        self.check(self.cdll.yb_link_forces(
            d_X, d_dX, lanes, n, d_links, n_links, strength), "yb_link_forces")

    def bending_force(self, Xi, Xj):
        Xi, Xj = _f32(Xi), _f32(Xj)
        out = np.zeros_like(Xi)
        self.check(self.cdll.yb_bending_force(
            Xi.ctypes.data, Xj.ctypes.data, len(Xi), out.ctypes.data),
            "yb_bending_force")
        return out

    def polarization_force(self, Xi, Xj):
        Xi, Xj = _f32(Xi), _f32(Xj)
        out = np.zeros_like(Xi)
        self.check(self.cdll.yb_polarization_force(
            Xi.ctypes.data, Xj.ctypes.data, len(Xi), out.ctypes.data),
            "yb_polarization_force")
        return out


class Sim:
    """A named model: Solution<Pt, Solver> plus its user code, behind the ABI."""

    def __init__(self, lib, model, n_max, grid_size=50, cube_size=1.0):
        self.lib = lib
        self.model = model
        handle = ctypes.c_void_p()
        lib.check(lib.cdll.yb_sim_create(
            model.encode(), n_max, grid_size, cube_size, ctypes.byref(handle)),
            f"yb_sim_create({model})")
        self.handle = handle
        self.lanes = lib.cdll.yb_sim_lanes(handle)
        self.n_max = lib.cdll.yb_sim_n_max(handle)

    def close(self):
        if self.handle:
            self.lib.cdll.yb_sim_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_param(self, name, value):
        self.lib.check(self.lib.cdll.yb_sim_set_param(
            self.handle, name.encode(), float(value)), f"set_param({name})")

    def set_state(self, X, reset_v=True):
        X = _f32(X).reshape(-1, self.lanes)
        self.lib.check(self.lib.cdll.yb_sim_set_state(
            self.handle, X.ctypes.data, len(X), int(reset_v)), "set_state")

    def get_state(self):
        out = np.zeros((self.n_max, self.lanes), dtype=np.float32)
        n = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_sim_get_state(
            self.handle, out.ctypes.data, self.n_max, ctypes.byref(n)),
            "get_state")
        return out[:n.value].copy()

    def get_velocities(self):
        out = np.zeros((self.n_max, 3), dtype=np.float32)
        self.lib.check(self.lib.cdll.yb_sim_get_velocities(
            self.handle, out.ctypes.data, self.n_max), "get_velocities")
        return out[:self.n()].copy()

    def set_velocities(self, v):
        v = _f32(v).reshape(-1, 3)
        self.lib.check(self.lib.cdll.yb_sim_set_velocities(
            self.handle, v.ctypes.data, len(v)), "set_velocities")

    def set_ints(self, name, values):
        values = _i32(values)
        self.lib.check(self.lib.cdll.yb_sim_set_ints(
            self.handle, name.encode(), values.ctypes.data, len(values)),
            f"set_ints({name})")

    def get_ints(self, name):
        out = np.zeros(self.n_max, dtype=np.int32)
        self.lib.check(self.lib.cdll.yb_sim_get_ints(
            self.handle, name.encode(), out.ctypes.data, self.n_max),
            f"get_ints({name})")
        return out[:self.n()].copy()

    def seed_sphere(self, n, dist_to_nb, seed, relax_steps=0):
        """Seeded ball generated (and optionally relaxed) on the device."""
        self.lib.check(self.lib.cdll.yb_sim_seed_sphere(
            self.handle, n, dist_to_nb, seed, relax_steps), "seed_sphere")

    def set_links(self, links):
        links = _i32(links).reshape(-1, 2)
        self.lib.check(self.lib.cdll.yb_sim_set_links(
            self.handle, links.ctypes.data, len(links)), "set_links")

    def get_links(self, capacity=None):
        capacity = capacity or 4 * self.n_max
        out = np.zeros((capacity, 2), dtype=np.int32)
        n = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_sim_get_links(
            self.handle, out.ctypes.data, capacity, ctypes.byref(n)), "get_links")
        return out[:n.value].copy()

    def step(self, dt, n_steps=1):
        self.lib.check(self.lib.cdll.yb_sim_step(self.handle, dt, n_steps),
                       "step")

    def step_timed(self, dt, n_steps):
        """-> (milliseconds on the device, cell updates done)"""
        ms = ctypes.c_float()
        updates = ctypes.c_longlong()
        self.lib.check(self.lib.cdll.yb_sim_step_timed(
            self.handle, dt, n_steps, ctypes.byref(ms), ctypes.byref(updates)),
            "step_timed")
        return ms.value, updates.value

    def step_host(self, X_in, dt, n_steps, X_out):
        """Host buffers in, host buffers out; X_out must hold n_max cells."""
        n = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_sim_step_host(
            self.handle, X_in.ctypes.data, len(X_in), dt, n_steps,
            X_out.ctypes.data, len(X_out), ctypes.byref(n)), "step_host")
        return n.value

    # ---- domain decomposition building blocks (pointers as ints) -----------
    def dd_load(self, stage, X_owned, v_owned, n_owned, X_ghost, v_ghost,
                n_ghost):
        self.lib.check(self.lib.cdll.yb_dd_load(
            self.handle, stage, X_owned, v_owned, n_owned, X_ghost, v_ghost,
            n_ghost), "dd_load")

    def dd_forces(self, stage, sums4):
        self.lib.check(self.lib.cdll.yb_dd_forces(self.handle, stage, sums4),
                       "dd_forces")

    def dd_update(self, stage, dt, mean3):
        self.lib.check(self.lib.cdll.yb_dd_update(self.handle, stage, dt, mean3),
                       "dd_update")

    def dd_read(self, which, out, n):
        self.lib.check(self.lib.cdll.yb_dd_read(self.handle, which, out, n),
                       "dd_read")

    def slab_begin(self, z_lo, z_hi, halo, capacity, first_layer=0, n_layers=0):
        self.lib.check(self.lib.cdll.yb_slab_begin(
            self.handle, z_lo, z_hi, halo, capacity, first_layer, n_layers),
            "slab_begin")

    def slab_set_owned(self, X, v, n_owned):
        self.lib.check(self.lib.cdll.yb_slab_set_owned(self.handle, X, v, n_owned),
                       "slab_set_owned")

    def slab_pack(self, what, send_lo, send_hi):
        self.lib.check(self.lib.cdll.yb_slab_pack(self.handle, what, send_lo,
                                                  send_hi), "slab_pack")

    def slab_unpack(self, what, recv_lo, recv_hi):
        self.lib.check(self.lib.cdll.yb_slab_unpack(self.handle, what, recv_lo,
                                                    recv_hi), "slab_unpack")

    def slab_update(self, stage, dt, sums4):
        self.lib.check(self.lib.cdll.yb_slab_update(self.handle, stage, dt, sums4),
                       "slab_update")

    def slab_counts(self):
        """-> (owned cells, owned + ghost cells, problems); blocks"""
        owned, total, problems = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_slab_counts(
            self.handle, ctypes.byref(owned), ctypes.byref(total),
            ctypes.byref(problems)), "slab_counts")
        return owned.value, total.value, problems.value

    # ---- brick decomposition over peer memory ----------------------------------
    def dom_begin(self, rank, world, lo, hi, halo, peer_ranks27, capacity27,
                  box_first=(0, 0, 0), box_n=(0, 0, 0)):
        lo, hi = _f32(lo), _f32(hi)
        peers, caps = _i32(peer_ranks27), _i32(capacity27)
        first, count = _i32(box_first), _i32(box_n)
        self.lib.check(self.lib.cdll.yb_dom_begin(
            self.handle, rank, world, lo.ctypes.data, hi.ctypes.data, halo,
            peers.ctypes.data, caps.ctypes.data, first.ctypes.data,
            count.ctypes.data), "dom_begin")

    def dom_register_array(self, d_array, bytes_per_cell, ghosts_too=False):
        """A per-cell device array that travels with the cells (before dom_begin)."""
        self.lib.check(self.lib.cdll.yb_dom_register_array(
            self.handle, d_array, bytes_per_cell, 1 if ghosts_too else 0),
            "dom_register_array")

    def dom_exchange(self):
        """-> (device address, bytes, offsets[27, DOM_OFFSETS]) of this rank's exchange
        allocation: per direction the offsets of 4 inboxes and 4 flag words."""
        base, size = ctypes.c_void_p(), ctypes.c_longlong()
        offsets = np.zeros((27, DOM_OFFSETS), dtype=np.int64)
        self.lib.check(self.lib.cdll.yb_dom_exchange(
            self.handle, ctypes.byref(base), ctypes.byref(size),
            offsets.ctypes.data), "dom_exchange")
        return base.value, size.value, offsets

    def dom_connect(self, direction, peer_base, peer_offsets):
        offsets = np.ascontiguousarray(peer_offsets, dtype=np.int64)
        assert offsets.size == DOM_OFFSETS
        self.lib.check(self.lib.cdll.yb_dom_connect(
            self.handle, direction, peer_base, offsets.ctypes.data), "dom_connect")

    def dom_connect_mailbox(self, rank, peer_base):
        self.lib.check(self.lib.cdll.yb_dom_connect_mailbox(
            self.handle, rank, peer_base), "dom_connect_mailbox")

    def dom_seed_lattice_ball(self, radius, dist_to_nb, jitter, seed):
        n = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_dom_seed_lattice_ball(
            self.handle, radius, dist_to_nb, jitter, seed, ctypes.byref(n)),
            "dom_seed_lattice_ball")
        return n.value

    def dom_step(self, dt, n_steps=1):
        self.lib.check(self.lib.cdll.yb_dom_step(self.handle, dt, n_steps),
                       "dom_step")

    def dom_read_profile(self):
        """-> dict of device milliseconds per phase of the decomposed steps
        taken since the last read (while profile_sweeps is on)."""
        ms = np.zeros(8, dtype=np.float32)
        self.lib.check(self.lib.cdll.yb_dom_read_profile(
            self.handle, ms.ctypes.data), "dom_read_profile")
        names = ("select_halo", "wait", "unpack", "forces", "drift_sum", "update",
                 "push", "select_migration")
        return {name: float(value) for name, value in zip(names, ms)}

    def profile_sweeps(self, enable=True):
        self.lib.check(self.lib.cdll.yb_sim_profile_sweeps(
            self.handle, int(enable)), "profile_sweeps")

    def read_sweep_profile(self):
        """-> (total milliseconds in sweep kernels, number of sweep launches)"""
        ms = ctypes.c_float()
        launches = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_sim_read_sweep_profile(
            self.handle, ctypes.byref(ms), ctypes.byref(launches)),
            "read_sweep_profile")
        return ms.value, launches.value

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t as int (e.g. torch.cuda.Stream().cuda_stream)"""
        self.lib.check(self.lib.cdll.yb_sim_set_stream(self.handle, cuda_stream),
                       "set_stream")

    def step_host_async(self, X_in, dt, n_steps, X_out, out_cells, n_out_ptr):
        """Enqueue upload, steps and download (copy streams + the model's
        stream); X_in, X_out and the int at n_out_ptr must be pinned host memory;
        host_drain() waits."""
        self.lib.check(self.lib.cdll.yb_sim_step_host_async(
            self.handle, X_in.ctypes.data, len(X_in), dt, n_steps,
            X_out.ctypes.data, out_cells, n_out_ptr), "step_host_async")

    def host_drain(self):
        """Wait for every batch enqueued with step_host_async."""
        self.lib.check(self.lib.cdll.yb_sim_host_drain(self.handle), "host_drain")

    def n(self):
        n = ctypes.c_int()
        self.lib.check(self.lib.cdll.yb_sim_n(self.handle, ctypes.byref(n)), "n")
        return n.value

    def sync(self):
        self.lib.check(self.lib.cdll.yb_sim_sync(self.handle), "sync")


def load(path):
    return Library(path)


_cache = {}


def product():
    """This repo's CUDA build. No fallback: fails loudly if it is missing."""
    if "product" not in _cache:
        _cache["product"] = Library(PRODUCT_LIB)
    return _cache["product"]


def reference():
    """The unmodified reference headers compiled for sm_100a (baseline arm)."""
    if "reference" not in _cache:
        _cache["reference"] = Library(REFERENCE_LIB)
    return _cache["reference"]
