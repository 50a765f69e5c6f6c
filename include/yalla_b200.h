/* C ABI of the ya||a hot path: grid build, link forces and whole model steps.
 *
 * ya||a itself has no FFI: a model is a .cu file that instantiates
 * Solution<Pt, Solver>::take_step<pairwise_fn> (reference include/solvers.cuh:
 * 60-106, 226-275), so the pairwise interaction is compiled INTO the force
 * kernel. The drop-in boundary for user models is therefore the header set in
 * this directory (solvers.cuh, links.cuh, ...). This C ABI is the second
 * boundary: it exposes the path for fixed, named models ("plugins") with plain
 * pointers and sizes, so that a host program in any language -- here the Python
 * test-suite and bench.py through ctypes -- can drive it, and so that the SAME
 * entry points can be served by three interchangeable libraries:
 *
 *   yalla_b200/_lib/libyalla_b200.so   this repo's headers        (the product)
 *   oracle/_ref/libyalla_ref.so        the unmodified reference headers,
 *                                      compiled from /root/reference (GPU)
 *   oracle/_build/libyalla_oracle.so   the CPU restatement in oracle/ (tests
 *                                      and cpu_baseline only)
 *
 * The first two are built from the one source yalla_b200/csrc/capi.cu, which
 * uses nothing but the public ya||a API -- that it compiles against both header
 * sets is the API-compatibility check.
 *
 * Conventions: every function returns 0 on success and a negative YB_E* code
 * otherwise; yb_last_error() describes the last failure of the calling
 * thread. State is array-of-structs float32: `lanes` floats per cell in the
 * member order of the model's point type (x, y, z, then extras). Pointers
 * named h_* are host memory, d_* device memory. Nothing here synchronises the
 * device unless it says so.
 */
#ifndef YALLA_B200_H
#define YALLA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define YB_OK 0
#define YB_EINVAL -1   /* bad argument (unknown model, n > n_max, ...)     */
#define YB_ECUDA -2    /* a CUDA call failed                                */
#define YB_ENOSYS -3   /* not available in this library (e.g. GPU entry in
                          the CPU oracle)                                   */

typedef struct yb_sim yb_sim;

/* "yalla-b200", "yalla-reference" or "yalla-oracle", plus build details. */
const char* yb_build_info(void);
const char* yb_last_error(void);

/* ---- whole-model steps ---------------------------------------------------
 * Each model is one Solution<Pt, Solver> plus the user code of a reference
 * example; one "step" is what that example does per iteration on the device
 * (take_step, and for growing models the division kernel) -- no host I/O.
 *
 *  name           Pt (lanes)  solver  follows
 *  "springs"      float3 (3)  Tile    examples/springs.cu:14-21 (L0 = 0.5)
 *  "spring_tile"  float3 (3)  Tile    tests/test_solvers.cu:44-53 clipped_spring
 *  "spring_grid"  float3 (3)  Grid    the same, Grid_solver
 *  "relu_tile"    float3 (3)  Tile    include/inits.cuh:78-93 relu_force
 *  "relu_grid"    float3 (3)  Grid    the same, Grid_solver
 *  "relu_gabriel" float3 (3)  Gabriel the same, Gabriel_solver (solvers.cuh:
 *                                     505-644; GPU libraries only)
 *  "epithelium"   Po_cell (5) Grid    examples/epithelium.cu:16-31 layer_force,
 *                                     friction_on_background
 *  "growth"       Po_cell (5) Grid    examples/passive_growth.cu: relu_w_epithelium
 *                                     :29-57, reset_nbs :107-113, proliferate
 *                                     :59-91 (curand, dynamic n)
 *  "protrusions"  float3 (3)  Grid    examples/sorting_prot.cu-style Links +
 *                                     link_forces as generic force, relu_force
 *  "branching"    Cell (7)    Grid    examples/branching.cu:57-110
 *                                     epi_turing_mes_noturing + counters
 *  "branching_growth" Cell (7) Grid   the same cell with division
 *                                     (branching.cu:113-138) and one protrusion
 *                                     per cell rewired every step
 *                                     (intercalation_w_gradient.cu:119-173,
 *                                     Grid::build + curand) pulling through
 *                                     link_forces: BASELINE.json configs[3].
 *                                     GPU libraries only (needs curand).
 *
 * grid_size / cube_size are the Grid_solver constructor arguments
 * (solvers.cuh:469) and are ignored by Tile models. */
int yb_sim_create(const char* model, int n_max, int grid_size,
    float cube_size, yb_sim** out);
void yb_sim_destroy(yb_sim* sim);
int yb_sim_lanes(const yb_sim* sim);
int yb_sim_n_max(const yb_sim* sim);

/* Model parameters by name (e.g. "prolif_rate", "mean_dist", "seed",
 * "link_strength", "fix_point"); must be set before the first step that uses
 * them. Unknown names return YB_EINVAL. */
int yb_sim_set_param(yb_sim* sim, const char* name, double value);

/* Host <-> device, blocking: Solution::copy_to_device / copy_to_host
 * (solvers.cuh:80-91) restricted to the first n cells of the caller's
 * buffer. set_state also resets old velocities to zero when reset_v != 0. */
int yb_sim_set_state(yb_sim* sim, const float* h_X, int n, int reset_v);
int yb_sim_get_state(yb_sim* sim, float* h_X, int capacity, int* n_out);
int yb_sim_get_velocities(yb_sim* sim, float* h_v, int capacity);
/* Replace the old velocities (d_old_v, 3 floats per cell) of the first n cells;
 * with yb_sim_set_state(..., reset_v = 0) this restarts a model from a state
 * another library produced (the step-by-step parity tests do). Blocks. */
int yb_sim_set_velocities(yb_sim* sim, const float* h_v, int n);

/* Integer per-cell properties of the model by name ("type", "mes_nbs",
 * "epi_nbs", ...): Property<int>::copy_to_device / copy_to_host. */
int yb_sim_set_ints(yb_sim* sim, const char* name, const int* h_values, int n);
int yb_sim_get_ints(yb_sim* sim, const char* name, int* h_values, int capacity);

/* Extension (product library only; YB_ENOSYS elsewhere): fill the model with
 * n cells of a seeded uniform ball at neighbour distance dist_to_nb, generated
 * on the device (b200/seeded_inits.cuh; radius formula of random_sphere,
 * inits.cuh:41-48), extra lanes and old velocities zero. With relax_steps != 0
 * the ball is generated at distance 0.6, relaxed with that many relu_force
 * steps (< 0: the reference's step count, inits.cuh:96-112) and rescaled like
 * relaxed_sphere. float3 Grid/Tile models only. */
int yb_sim_seed_sphere(yb_sim* sim, int n, float dist_to_nb,
    unsigned long long seed, int relax_steps);

/* Links of models that have them: pairs (a, b) as 2 * n_links ints. get_links
 * waits for the model's stream and copies the current links back
 * (Links::copy_to_host). */
int yb_sim_set_links(yb_sim* sim, const int* h_links, int n_links);
int yb_sim_get_links(yb_sim* sim, int* h_links, int capacity, int* n_out);

/* Enqueue n_steps model steps of size dt. Does not wait for them. */
int yb_sim_step(yb_sim* sim, float dt, int n_steps);
/* Same, bracketed by CUDA events on the launching stream; waits, and returns
 * the device time in milliseconds and the sum over steps of the cell count
 * at the start of each step (the numerator of cell-updates/s). */
int yb_sim_step_timed(yb_sim* sim, float dt, int n_steps, float* ms_out,
    long long* cell_updates_out);
/* Host-buffer round trip, timed on the host clock by the caller: copies the
 * n cells at h_in to the device, takes n_steps steps, copies the state back
 * to h_out and waits. */
int yb_sim_step_host(yb_sim* sim, const float* h_in, int n, float dt,
    int n_steps, float* h_out, int capacity, int* n_out);

/* Extensions for pipelining independent batches (product library only).
 * yb_sim_set_stream gives a model its own CUDA stream (a cudaStream_t;
 * default: the legacy default stream like every ya||a launch); it waits for
 * the work already issued to the old one.
 * yb_sim_step_host_async only ENQUEUES: the upload of n cells from h_in
 * (pinned) on a copy stream, n_steps steps on the model's stream, and the
 * download of out_cells cells to h_out plus the final cell count to *h_n_out
 * (both pinned) on a second copy stream, through two staging slots in device
 * memory -- so the copies of consecutive calls overlap the steps of the batch
 * in between, with ONE model instance. At most two batches are in flight; the
 * third call waits for the first one's buffers. yb_sim_host_drain waits until
 * everything enqueued so far has landed in the host buffers.
 *
 * The "growth" and "branching" models reach their Property arrays through
 * process-global __device__ pointers (as the reference examples do), so only
 * one instance of them is active at a time: a step of another instance first
 * waits for the stream of the one that ran before. */
int yb_sim_set_stream(yb_sim* sim, void* stream);
int yb_sim_step_host_async(yb_sim* sim, const float* h_in, int n, float dt,
    int n_steps, float* h_out, int out_cells, int* h_n_out);
int yb_sim_host_drain(yb_sim* sim);

/* Time the dominant kernel (the pairwise sweep) with CUDA events on the
 * launching stream: enable, run steps, then read the accumulated milliseconds
 * and the number of sweep launches since the last read. Product library only
 * (YB_ENOSYS elsewhere); used by bench.py for the roofline line. */
int yb_sim_profile_sweeps(yb_sim* sim, int enable);
int yb_sim_read_sweep_profile(yb_sim* sim, float* total_ms, int* launches);

/* ---- domain decomposition building blocks ------------------------------------
 * For tissues larger than one GPU the domain is cut into slabs, one solver per
 * slab and GPU (yalla_b200/dd.py; not part of the reference, which is single-
 * GPU). A solver then holds its n_owned cells followed by n_ghost GHOST cells --
 * copies of the neighbouring slabs' boundary cells, interaction partners only.
 * Pointers are in the library's memory space (device memory for the CUDA
 * libraries, host memory for the CPU oracle). Grid models without per-id
 * property arrays only ("relu_grid", "spring_grid", "epithelium").
 *
 * yb_dd_load    stage 0: replace the owned state (positions/extras X, old
 *               velocities v) and append the ghosts; stage 1: keep the owned
 *               predictor positions, append the ghosts' predictor positions.
 * yb_dd_forces  grid build + pairwise sweep of the stage; writes
 *               {sum dX.x, sum dX.y, sum dX.z, n_owned} of the OWNED cells to
 *               sums4 -- to be summed over all domains by the caller.
 * yb_dd_update  predictor (stage 0) / corrector (stage 1) with the global mean
 *               force mean3 = sum / n removed (solvers.cuh:241-255, 266-274).
 * yb_dd_read    copy n owned cells of X (0), X1 (1) or old velocities (2). */
int yb_dd_load(yb_sim* sim, int stage, const float* X_owned,
    const float* v_owned, int n_owned, const float* X_ghost,
    const float* v_ghost, int n_ghost);
int yb_dd_forces(yb_sim* sim, int stage, float* sums4);
int yb_dd_update(yb_sim* sim, int stage, float dt, const float* mean3);
int yb_dd_read(yb_sim* sim, int which, float* out, int n);

/* Slab decomposition without host round trips: packing/unpacking of halo and
 * migrating cells happens inside the library and all counts stay in its memory
 * space. Exchange buffers hold 4 header floats (header[0] = bits of the record
 * count) followed by `capacity` records of lanes + 3 floats (cell, old
 * velocity); the caller ships whole buffers to the neighbouring ranks.
 *
 * yb_slab_begin   this solver owns z_lo <= z < z_hi (+-INFINITY at the ends);
 *                 halo = width of the strip sent to a neighbour; the grid is
 *                 restricted to n_layers z layers starting at first_layer of
 *                 the cubic grid (n_layers <= 0: keep the whole grid).
 * yb_slab_pack    what 0 / 1: halo of X / X1 -> send_lo, send_hi;
 *                 what 2: cells that left the slab (and compaction of the rest)
 * yb_slab_unpack  what 0 / 1: append received ghosts; 2: append arrivals
 * yb_slab_update  predictor / corrector, drift = sums4[0..2] / sums4[3] with
 *                 sums4 already reduced over all slabs
 * yb_slab_counts  blocking diagnostics: owned, owned + ghosts, problems */
int yb_slab_begin(yb_sim* sim, float z_lo, float z_hi, float halo,
    int capacity, int first_layer, int n_layers);
int yb_slab_set_owned(yb_sim* sim, const float* X, const float* v, int n_owned);
int yb_slab_pack(yb_sim* sim, int what, float* send_lo, float* send_hi);
int yb_slab_unpack(yb_sim* sim, int what, const float* recv_lo,
    const float* recv_hi);
int yb_slab_update(yb_sim* sim, int stage, float dt, const float* sums4);
int yb_slab_counts(yb_sim* sim, int* n_owned, int* n_total, int* problems);

/* ---- brick decomposition over peer memory (product library only) ---------------
 * The tissue is cut into bricks, one model instance per brick and GPU; the
 * halo exchange, the migration of cells and the global drift sum are done by
 * the library's kernels storing straight into the neighbours' memory over
 * NVLink (include/b200/domain.cuh) -- yb_dom_step never waits for the host and
 * calls no communication library. Set-up, once:
 *
 *  yb_dom_begin     this instance is rank `rank` of `world`; it owns
 *                   lo3 <= (x, y, z) < hi3 (+-INFINITY at the tissue's ends) and
 *                   copies cells within `halo` of a face to the neighbour
 *                   behind it. peer_ranks27[d] is the rank of the neighbour in
 *                   direction d = (dx + 1) + 3 (dy + 1) + 9 (dz + 1), or -1;
 *                   capacity27[d] the records its inbox for that neighbour
 *                   holds. box_first3 / box_n3: the cubes of the global grid
 *                   this brick can touch (box_n3[0] <= 0: the whole grid).
 *  yb_dom_exchange  device address and size of this rank's exchange allocation
 *                   and, per direction d, the byte offsets of its four
 *                   inboxes (halo of either Heun stage, migration, survey) and
 *                   four flag words: offsets_out[YB_DOM_OFFSETS d + q], -1
 *                   without a neighbour. Other processes map the allocation
 *                   with yb_ipc_export / yb_ipc_import.
 *  yb_dom_connect   the neighbour in `direction` keeps this rank's records at
 *                   d_peer_base (its allocation as mapped here) + the offsets
 *                   IT reported for the opposite direction, 26 - direction.
 *  yb_dom_connect_mailbox   every rank's allocation (its mailbox is at offset
 *                   0), this rank's own included, for the drift sum.
 *
 * yb_dom_seed_lattice_ball fills the brick with its share of a jittered FCC
 * ball generated on the device (same tissue however it is cut);
 * yb_slab_set_owned / yb_dd_read / yb_slab_counts load, read and count cells. */
int yb_dom_begin(yb_sim* sim, int rank, int world, const float* lo3,
    const float* hi3, float halo, const int* peer_ranks27,
    const int* capacity27, const int* box_first3, const int* box_n3);
/* Before yb_dom_begin: a per-cell device array of the caller (n_max entries of
 * bytes_per_cell bytes, a multiple of 4, indexed like the cells) that travels
 * with the cells -- it migrates with its cell and is re-stored with it, and
 * with ghosts_too != 0 it also comes along with the ghost copies (entries
 * n_owned.. of the array then hold the ghosts' values during a step). This is
 * how Property<T> arrays, curand states and cell identities survive the
 * decomposition (the typed models register their own; at most 8 arrays). */
int yb_dom_register_array(yb_sim* sim, void* d_array, int bytes_per_cell,
    int ghosts_too);
#define YB_DOM_OFFSETS 8 /* entries per direction in the offset tables */
int yb_dom_exchange(yb_sim* sim, void** d_base_out, long long* bytes_out,
    long long* offsets_out /* [27 * YB_DOM_OFFSETS] */);
int yb_dom_connect(yb_sim* sim, int direction, void* d_peer_base,
    const long long* peer_offsets /* [YB_DOM_OFFSETS] */);
int yb_dom_connect_mailbox(yb_sim* sim, int rank, void* d_peer_base);
int yb_dom_seed_lattice_ball(yb_sim* sim, float radius, float dist_to_nb,
    float jitter, unsigned long long seed, int* n_out);
int yb_dom_step(yb_sim* sim, float dt, int n_steps);
/* While yb_sim_profile_sweeps is on: device milliseconds since the last read
 * spent in {packing the halo rounds, waiting for the neighbours' flags,
 * unpacking, grid build + sweep, drift sum, update, pushing the outboxes to the
 * neighbours, the migration round's packing and re-store} (CUDA events; 8
 * floats). */
int yb_dom_read_profile(yb_sim* sim, float* ms8);
/* CUDA IPC plumbing for the above: 64-byte handle of a device allocation,
 * mapping of another process's handle, unmapping. */
int yb_ipc_export(const void* d_base, unsigned char* handle64);
int yb_ipc_import(const unsigned char* handle64, void** d_base_out);
int yb_ipc_release(void* d_base);

/* Current cell count (blocking read of d_n: Solution::get_d_n). */
int yb_sim_n(yb_sim* sim, int* n_out);
int yb_sim_sync(yb_sim* sim);

/* ---- grid build ------------------------------------------------------------
 * Grid::build (solvers.cuh:380-425) on n points of `lanes` floats each
 * (3, 4, 5 or 7) resident at d_X. Writes the four public arrays:
 * d_cube_id[n] sorted cube ids, d_point_id[n] original index per sorted slot,
 * d_cube_start / d_cube_end [grid_size^3], inclusive, -1 / -2 when empty.
 * Blocks until done. */
int yb_grid_build(const float* d_X, int lanes, int n, int grid_size,
    float cube_size, int* d_cube_id, int* d_point_id, int* d_cube_start,
    int* d_cube_end);
/* The 27 neighbour-cube offsets (d_nhood, solvers.cuh:428, :472-484) for a
 * grid size, in the order user kernels index them. */
int yb_nhood(int grid_size, int* h_nhood27);

/* ---- link forces -------------------------------------------------------------
 * link_forces<Pt> (links.cuh:128-133) with linear_force on n cells of
 * `lanes` floats: adds -/+ strength * r / |r| to d_dX for every link (a, b)
 * with a != b. d_links holds 2 * n_links ints. Blocks until done. */
int yb_link_forces(const float* d_X, float* d_dX, int lanes, int n,
    const int* d_links, int n_links, float strength);

/* ---- polarity forces, evaluated once on the device ----------------------------
 * bending_force (polarity.cuh:74-94) and bidirectional_polarization_force
 * (:65-69) for n_pairs independent (Xi, Xj) pairs of Po_cell (5 floats each):
 * out[k] = f(Xi[k], r = Xi[k] - Xj[k], |r|)  resp.  f(Xi[k], pol(Xj[k])). */
int yb_bending_force(const float* h_Xi, const float* h_Xj, int n_pairs,
    float* h_out);
int yb_polarization_force(const float* h_Xi, const float* h_Xj, int n_pairs,
    float* h_out);

#ifdef __cplusplus
}
#endif
#endif /* YALLA_B200_H */
