// Grid solver with Gabriel-graph pruning of the neighbourhood (Delile et al.
// 2017, Nat. Commun.; Marin-Riera et al. 2016, Bioinformatics): the pair i-j
// only interacts if no third cell lies inside the sphere whose diameter is
// gabriel_coefficient * |r_ij| around their midpoint.
//
// Reference: compute_cube_gabriel / Gabriel_computer, solvers.cuh:505-644.
// Included at the end of solvers.cuh. It reuses the grid solver's bucket sort
// and cube-ordered planes; the kernel itself stays close to the reference's
// structure (collect candidates, order them by distance with the same
// selection sort so ties break identically, prune from the far end) because
// it is used on small sheets only and is not a throughput target.
#pragma once

namespace yb {

constexpr int GABRIEL_THREADS = 64;
constexpr int GABRIEL_MAX_NEIGHBOURS = 100;

template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED>
__global__ void __launch_bounds__(GABRIEL_THREADS) sweep_gabriel(
    const int* __restrict__ d_n, int n_max, const float4* __restrict__ pos4,
    const float4* __restrict__ aux, const int* __restrict__ cube_sorted,
    const int* __restrict__ offset, float cube_size, Grid_box box,
    float gabriel_coefficient, Pt* d_dX, float* __restrict__ partials,
    int stage, int drift_mode, int fix_point, Step_ctl* ctl, int only_if_overflow)
{
    using L = Layout<Pt>;
    __shared__ float s_red[3][GABRIEL_THREADS / 32];
    // behind list_cubes + gabriel_lists: only if a neighbour list overflowed
    if (only_if_overflow && *(volatile int*)&ctl->list_overflow == 0) return;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, GABRIEL_THREADS);
    float3 cta_sum{0.f, 0.f, 0.f};

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int k = chunk * GABRIEL_THREADS + threadIdx.x;
        float3 mine{0.f, 0.f, 0.f};
        if (k < n) {
            const float4 me = __ldg(pos4 + k);
            const int my_id = __float_as_int(me.w);
            const int my_cube = __ldg(cube_sorted + k);
            const Pt Xi = assemble_pt<Pt>(me, aux + size_t(k) * L::aux_vec4);

            int nb_slot[GABRIEL_MAX_NEIGHBOURS];
            float nb_dist[GABRIEL_MAX_NEIGHBOURS];
            int n_nbs = 0;

            // 1. every cell within cube_size, in the reference's sweep order
            for (int r = 0; r < SWEEP_ROWS; r++) {
                const int c = my_cube + row_shift(r, box);
                const int lo = __ldg(offset + clamp_cube(c - 1, box.n_cubes));
                const int hi = __ldg(offset + clamp_cube(c + 2, box.n_cubes));
                for (int q = lo; q < hi; q++) {
                    const float4 pj = __ldg(pos4 + q);
                    // r = Xi - Xj lane-wise: x - x', as in operator-
                    const float dist =
                        norm3df(me.x - pj.x, me.y - pj.y, me.z - pj.z);
                    if (dist >= cube_size) continue;
                    if (n_nbs == GABRIEL_MAX_NEIGHBOURS) continue;
                    nb_slot[n_nbs] = q;
                    nb_dist[n_nbs] = dist;
                    n_nbs++;
                }
            }

            // 2. ascending distance (selection sort: same tie-breaking)
            for (int m = 0; m < n_nbs - 1; m++) {
                float least = nb_dist[m];
                int at = m;
                for (int q = m + 1; q < n_nbs; q++) {
                    if (nb_dist[q] < least) {
                        least = nb_dist[q];
                        at = q;
                    }
                }
                if (at != m) {
                    const int slot = nb_slot[at];
                    nb_slot[at] = nb_slot[m];
                    nb_slot[m] = slot;
                    nb_dist[at] = nb_dist[m];
                    nb_dist[m] = least;
                }
            }

            // 3. farthest first: keep i-j unless a closer cell sits inside the
            //    (shrunken) sphere around the midpoint of i and j
            Pt F{0};
            float3 sum_v{0.f, 0.f, 0.f};
            float sum_friction = 0.f;
            for (int m = n_nbs - 1; m >= 0; m--) {
                const int kj = nb_slot[m];
                const float4 pj = __ldg(pos4 + kj);
                const int j_id = __float_as_int(pj.w);
                const float4* aux_j = aux + size_t(kj) * L::aux_vec4;
                const Pt Xj = assemble_pt<Pt>(pj, aux_j);
                const float dist = nb_dist[m];
                bool keep = true;
                if (j_id != my_id) {
                    const float radius = 0.5f * dist * gabriel_coefficient;
                    const Pt mid_point = 0.5f * (Xi + Xj);
                    for (int q = m - 1; q >= 0; q--) {
                        const float4 pk = __ldg(pos4 + nb_slot[q]);
                        const float dist_mk = norm3df(mid_point.x - pk.x,
                            mid_point.y - pk.y, mid_point.z - pk.z);
                        if (dist_mk < radius) {
                            keep = false;
                            break;
                        }
                    }
                }
                if (!keep) continue;
                const Pt rij = Xi - Xj;
                F += pw_int(Xi, rij, dist, my_id, j_id);
                const float friction = pw_friction(Xi, rij, dist, my_id, j_id);
                sum_friction += friction;
                if (friction != 0.f) sum_v += friction * velocity_of<Pt>(aux_j);
            }

            Pt dX = F;
            if (SEEDED) {
                dX = load_pt_rw(d_dX, my_id);
                dX += F;
            }
            if (sum_friction > 0) {
                dX.x += sum_v.x / sum_friction;
                dX.y += sum_v.y / sum_friction;
                dX.z += sum_v.z / sum_friction;
            }
            store_pt(d_dX, my_id, dX);
            mine = float3{dX.x, dX.y, dX.z};
        }
        __syncthreads();
        const float3 chunk_sum =
            block_sum3<GABRIEL_THREADS>(mine.x, mine.y, mine.z, s_red);
        cta_sum.x += chunk_sum.x, cta_sum.y += chunk_sum.y, cta_sum.z += chunk_sum.z;
    }
    finish_drift<GABRIEL_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}

// The same sum fed from the neighbour lists of list_cubes (pair_sweep.cuh): the
// candidate scan is the Grid solver's -- spans staged by bulk copies, cheap
// squared-distance filter, survivors in the reference's sweep order -- instead
// of 27 cubes' worth of norm3df over global memory per cell; what remains per
// cell is the exact cut-off for the ~15 listed candidates, the reference's
// selection sort, and the pruning. The pruning test dist(mid, k) < radius is
// decided on squared distances wherever that is safe and by the reference's
// own expression (norm3df) within a relative 1e-5 of the threshold, so the
// decisions are the reference's. A cell with more than LIST_MAX listed
// candidates sends the whole stage to sweep_gabriel (launched behind).
// Measured at 1 M cells (profiles/r02_gabriel_ab.log), 8 CTAs per SM: 1.64 ms
// per step against 1.77 with sweep_gabriel alone and 8.1 with the reference's
// build; with grids sized by occupancy 1.42 against 1.31 -- see use_lists.
// Keeping the per-cell arrays in shared memory instead of thread-local memory
// was tried: 40 KB per 64-thread CTA leave 10 warps per SM, 2.21 ms.
template<typename Pt, Pt (*pw_int)(Pt, Pt, float, int, int),
    float (*pw_friction)(Pt, Pt, float, int, int), bool SEEDED>
__global__ void __launch_bounds__(GABRIEL_THREADS) gabriel_lists(
    const int* __restrict__ d_n, int n_max, const float4* __restrict__ pos4,
    const float4* __restrict__ aux, const int* __restrict__ nb,
    const int* __restrict__ nb_count, int nb_stride, float cube_size,
    float gabriel_coefficient, Pt* d_dX, float* __restrict__ partials,
    int stage, int drift_mode, int fix_point, Step_ctl* ctl)
{
    using L = Layout<Pt>;
    __shared__ float s_red[3][GABRIEL_THREADS / 32];
    if (*(volatile int*)&ctl->list_overflow) return;
    const int n = live_cells(d_n, n_max);
    const int n_chunks = ceil_div(n, GABRIEL_THREADS);
    const int n_owned = ctl->external_drift ? ctl->n_owned : n_max;
    float3 cta_sum{0.f, 0.f, 0.f};

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int k = chunk * GABRIEL_THREADS + threadIdx.x;
        float3 mine{0.f, 0.f, 0.f};
        float4 me = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < n) me = __ldg(pos4 + k);
        const int my_id = __float_as_int(me.w);
        if (k < n && my_id < n_owned) {
            const Pt Xi = assemble_pt<Pt>(me, aux + size_t(k) * L::aux_vec4);
            const int listed = min(__ldg(nb_count + k), LIST_MAX);

            int nb_slot[LIST_MAX];
            float nb_dist[LIST_MAX];
            int n_nbs = 0;
            // 1. exact cut-off, in the order of the list (= the reference's sweep)
            for (int e = 0; e < listed; e++) {
                const int q = __ldg(nb + size_t(e) * nb_stride + k);
                const float4 pj = __ldg(pos4 + q);
                const float dist = pair_distance(
                    me.x - pj.x, me.y - pj.y, me.z - pj.z, q == k);
                if (dist >= cube_size) continue;
                nb_slot[n_nbs] = q;
                nb_dist[n_nbs] = dist;
                n_nbs++;
            }

            // 2. ascending distance (selection sort: same tie-breaking)
            for (int m = 0; m < n_nbs - 1; m++) {
                float least = nb_dist[m];
                int at = m;
                for (int q = m + 1; q < n_nbs; q++) {
                    if (nb_dist[q] < least) {
                        least = nb_dist[q];
                        at = q;
                    }
                }
                if (at != m) {
                    const int slot = nb_slot[at];
                    nb_slot[at] = nb_slot[m];
                    nb_slot[m] = slot;
                    nb_dist[at] = nb_dist[m];
                    nb_dist[m] = least;
                }
            }

            // 3. farthest first: keep i-j unless a closer cell sits inside the
            //    (shrunken) sphere around the midpoint of i and j
            Pt F{0};
            float3 sum_v{0.f, 0.f, 0.f};
            float sum_friction = 0.f;
            for (int m = n_nbs - 1; m >= 0; m--) {
                const int kj = nb_slot[m];
                const float4 pj = __ldg(pos4 + kj);
                const int j_id = __float_as_int(pj.w);
                const float4* aux_j = aux + size_t(kj) * L::aux_vec4;
                const Pt Xj = assemble_pt<Pt>(pj, aux_j);
                const float dist = nb_dist[m];
                bool keep = true;
                if (j_id != my_id) {
                    const float radius = 0.5f * dist * gabriel_coefficient;
                    const Pt mid_point = 0.5f * (Xi + Xj);
                    const float radius2 = radius * radius;
                    const float band = 1e-5f * radius2;
                    for (int q = m - 1; q >= 0; q--) {
                        const float4 pk = __ldg(pos4 + nb_slot[q]);
                        const float dx = mid_point.x - pk.x, dy = mid_point.y - pk.y,
                                    dz = mid_point.z - pk.z;
                        const float d2 = dx * dx + dy * dy + dz * dz;
                        bool inside = d2 < radius2;
                        if (fabsf(d2 - radius2) <= band)  // too close to call
                            inside = norm3df(dx, dy, dz) < radius;
                        if (inside) {
                            keep = false;
                            break;
                        }
                    }
                }
                if (!keep) continue;
                const Pt rij = Xi - Xj;
                F += pw_int(Xi, rij, dist, my_id, j_id);
                const float friction = pw_friction(Xi, rij, dist, my_id, j_id);
                sum_friction += friction;
                if (friction != 0.f) sum_v += friction * velocity_of<Pt>(aux_j);
            }

            Pt dX = F;
            if (SEEDED) {
                dX = load_pt_rw(d_dX, my_id);
                dX += F;
            }
            if (sum_friction > 0) {
                dX.x += sum_v.x / sum_friction;
                dX.y += sum_v.y / sum_friction;
                dX.z += sum_v.z / sum_friction;
            }
            store_pt(d_dX, my_id, dX);
            mine = float3{dX.x, dX.y, dX.z};
        }
        __syncthreads();
        const float3 chunk_sum =
            block_sum3<GABRIEL_THREADS>(mine.x, mine.y, mine.z, s_red);
        cta_sum.x += chunk_sum.x, cta_sum.y += chunk_sum.y, cta_sum.z += chunk_sum.z;
    }
    finish_drift<GABRIEL_THREADS>(cta_sum, partials, n, stage, drift_mode,
        fix_point, d_dX, ctl, s_red);
}

}  // namespace yb


template<typename Pt>
class Gabriel_computer : public Grid_computer<Pt> {
public:
    float gabriel_coefficient;

    Gabriel_computer(int n_max, int grid_size = 50, float cube_size = 1,
        float gabriel_coefficient = 0.8)
        : Grid_computer<Pt>{n_max, grid_size, cube_size},
          gabriel_coefficient{gabriel_coefficient}
    {
        this->allocate_lists();
    }

    // Extension: true = take the candidates from the staged neighbour lists
    // (list_cubes + gabriel_lists) instead of scanning the 27 cubes in
    // sweep_gabriel; YALLA_B200_GABRIEL_LISTS=1 sets it for every solver. Off by
    // default: once both kernels got grids sized by occupancy (20 CTAs of 64
    // threads per SM instead of 8) the plain kernel was the faster one at 1 M
    // cells, 1.31 against 1.42 ms per step (profiles/r02_gabriel_ab.log).
    bool use_lists = [] {
        const char* env = getenv("YALLA_B200_GABRIEL_LISTS");
        return env && env[0] == '1';
    }();

protected:
    // both knobs are baked into captured graphs
    yb::Graph_key graph_key() const
    {
        return yb::Graph_key{this->cube_size, gabriel_coefficient,
            (this->box.z_half * 1024 + this->box.y_half) * 1024 + this->box.x_half,
            this->box.n_cubes * 2 + (use_lists ? 1 : 0)};
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    int prepare()
    {
        Grid_computer<Pt>::prepare_list();  // attributes: not inside a capture
        // The per-cell arrays live in thread-local memory, so the kernels wait
        // for the L1 most of the time: as many 64-thread CTAs as the registers
        // allow (the ncu capture of the first version, sized for 8 CTAs = 16
        // warps per SM, showed 24 % issue-slot use with long-scoreboard stalls).
        static const int ctas_per_sm = [] {
            int lists = 0, plain = 0;
            YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&lists,
                yb::gabriel_lists<Pt, pw_int, pw_friction, SEEDED>,
                yb::GABRIEL_THREADS, 0));
            YB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&plain,
                yb::sweep_gabriel<Pt, pw_int, pw_friction, SEEDED>,
                yb::GABRIEL_THREADS, 0));
            int ctas = lists < plain ? lists : plain;
            const char* cap = getenv("YALLA_B200_GABRIEL_CTAS");
            if (cap && cap[0] && atoi(cap) > 0) ctas = atoi(cap);
            return ctas < 1 ? 1 : (ctas > 32 ? 32 : ctas);
        }();
        return ctas_per_sm;
    }

    template<Pairwise_interaction<Pt> pw_int, Pairwise_friction<Pt> pw_friction,
        bool SEEDED>
    void pwints(cudaStream_t s, const int* d_n, const Pt* d_X,
        const float3* d_old_v, Pt* d_dX, float* d_partials, int max_ctas,
        int stage, int drift_mode, int fix_point, yb::Step_ctl* d_ctl,
        bool binned_by_predictor, cudaEvent_t before_sweep = nullptr)
    {
        if (!this->take_index_ahead())
            this->build_index(s, d_n, d_X, d_old_v, d_ctl, binned_by_predictor);
        const int ctas = this->persistent_ctas(
            prepare<pw_int, pw_friction, SEEDED>(), yb::GABRIEL_THREADS, max_ctas);
        if (before_sweep) YB_CUDA(cudaEventRecord(before_sweep, s));
        if (use_lists) {
            yb::list_cubes<<<this->persistent_ctas(Grid_computer<Pt>::prepare_list(),
                                 yb::SWEEP_THREADS, max_ctas),
                yb::SWEEP_THREADS, yb::List_config::smem, s>>>(d_n, this->n_max,
                this->pos4, this->cube_sorted, this->sort.offset, this->cube_size,
                this->box, this->nb, this->nb_count, this->nb_order,
                this->nb_stride, d_ctl, yb::LIST_MAX);
            yb::gabriel_lists<Pt, pw_int, pw_friction, SEEDED>
                <<<ctas, yb::GABRIEL_THREADS, 0, s>>>(d_n, this->n_max, this->pos4,
                    this->aux, this->nb, this->nb_count, this->nb_stride,
                    this->cube_size, gabriel_coefficient, d_dX, d_partials, stage,
                    drift_mode, fix_point, d_ctl);
        }
        yb::sweep_gabriel<Pt, pw_int, pw_friction, SEEDED>
            <<<ctas, yb::GABRIEL_THREADS, 0, s>>>(d_n, this->n_max, this->pos4,
                this->aux, this->cube_sorted, this->sort.offset, this->cube_size,
                this->box, gabriel_coefficient, d_dX,
                d_partials, stage, drift_mode, fix_point, d_ctl, use_lists ? 1 : 0);
    }
};

template<typename Pt>
using Gabriel_solver = Heun_solver<Pt, Gabriel_computer>;
