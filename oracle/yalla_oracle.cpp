// CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain C++ restatement of the reference's per-step hot path, exporting the
// C ABI of include/yalla_b200.h so that tests can drive it exactly like the
// two GPU libraries. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / reference legs may load this library; nothing under
// yalla_b200/ or include/ does.
//
// It follows the reference line by line, NOT this repo's kernels: the grid is
// built with a float cube id and a stable sort and swept through
// cube_start/cube_end and the 27-entry neighbourhood table, state stays
// array-of-structs in original order, the drift is a plain sum / n.
//   cube id                 /root/reference/include/solvers.cuh:350-365
//   sort + start/end        solvers.cuh:367-378, 406-417
//   neighbourhood table     solvers.cuh:472-484
//   grid sweep              solvers.cuh:432-463
//   all-pairs sweep         solvers.cuh:286-322
//   friction term           solvers.cuh:147-161, 27-41
//   Heun step, drift        solvers.cuh:226-275, 114-144
//   vector arithmetic       dtypes.cuh:151-217 (a /= b is a *= float(1. / b))
//   links                   links.cuh:99-125
//   polarity forces         polarity.cuh:13-94
//   models                  the examples cited in include/yalla_b200.h
//
// Pinning (see tests/test_oracle.py): the known-answer vectors of
// tests/test_polarity.cu:20-34, 78-94, the lattice of tests/test_solvers.cu:
// 247-315, the behavioural pins of test_solvers.cu / test_links.cu, and golden
// outputs of the reference's own sm_100a build (tests/golden/, generated on a
// B200 by scripts/make_golden.py).
//
// Differences from the device that tests allow for: libm sinf/cosf/acosf/atan2f
// instead of CUDA's, a correctly rounded sqrt instead of norm3df (<= 1 ulp), no
// FMA contraction (built with -ffp-contract=off). No curand: the growth model
// runs with proliferation switched off only.
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <functional>
#include <numeric>
#include <string>
#include <vector>

#include "../include/yalla_b200.h"

namespace {

thread_local std::string last_error;

int fail(int code, const std::string& what)
{
    last_error = what;
    return code;
}

// ---- vector space over L floats (dtypes.cuh) ------------------------------------
template<int L>
struct Pt {
    float v[L];
    float& x() { return v[0]; }
    float& y() { return v[1]; }
    float& z() { return v[2]; }
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
};

template<int L>
Pt<L> zero()
{
    Pt<L> a;
    for (int k = 0; k < L; k++) a.v[k] = 0.f;
    return a;
}
template<int L>
Pt<L>& operator+=(Pt<L>& a, const Pt<L>& b)
{
    for (int k = 0; k < L; k++) a.v[k] += b.v[k];
    return a;
}
template<int L>
Pt<L>& operator*=(Pt<L>& a, float b)
{
    for (int k = 0; k < L; k++) a.v[k] *= b;
    return a;
}
template<int L>
Pt<L> operator*(Pt<L> a, float b)
{
    a *= b;
    return a;
}
template<int L>
Pt<L> operator+(Pt<L> a, const Pt<L>& b)
{
    a += b;
    return a;
}
// a - b is a + (-1 * b)   (dtypes.cuh:160-175)
template<int L>
Pt<L> operator-(Pt<L> a, const Pt<L>& b)
{
    a += b * -1.f;
    return a;
}
// a / b is a * float(1. / b)   (dtypes.cuh:202-217)
template<int L>
Pt<L> operator/(Pt<L> a, float b)
{
    a *= static_cast<float>(1. / b);
    return a;
}

// norm3df stand-in: correctly rounded Euclidean norm
inline float norm3(float x, float y, float z)
{
    return static_cast<float>(
        sqrt(double(x) * x + double(y) * y + double(z) * z));
}

// ---- polarity (polarity.cuh) ------------------------------------------------------
struct Polarity {
    float theta, phi;
};
struct F3 {
    float x, y, z;
};

// polarity.cuh:13-21
inline F3 pol_to_float3(Polarity p)
{
    return F3{sinf(p.theta) * cosf(p.phi), sinf(p.theta) * sinf(p.phi),
        cosf(p.theta)};
}
// polarity.cuh:23-28
inline Polarity pt_to_pol(float rx, float ry, float rz, float dist)
{
    return Polarity{acosf(rz / dist), atan2f(ry, rx)};
}
// polarity.cuh:41-46
inline float pol_dot_product(Polarity a, Polarity p)
{
    return sinf(a.theta) * sinf(p.theta) * cosf(a.phi - p.phi) +
           cosf(a.theta) * cosf(p.theta);
}
// polarity.cuh:50-60; returns (d theta, d phi)
inline Polarity unidirectional_polarization_force(Polarity Xi, Polarity p)
{
    Polarity dF{0.f, 0.f};
    dF.theta = cosf(Xi.theta) * sinf(p.theta) * cosf(Xi.phi - p.phi) -
               sinf(Xi.theta) * cosf(p.theta);
    const float sin_Xi_theta = sinf(Xi.theta);
    if (fabs(sin_Xi_theta) > 1e-10)
        dF.phi = -sinf(p.theta) * sinf(Xi.phi - p.phi) / sin_Xi_theta;
    return dF;
}
// polarity.cuh:64-69
inline Polarity bidirectional_polarization_force(Polarity Xi, Polarity p)
{
    const float prod = pol_dot_product(Xi, p);
    const Polarity uni = unidirectional_polarization_force(Xi, p);
    return Polarity{prod * uni.theta, prod * uni.phi};
}
// polarity.cuh:73-94 for a point with theta, phi in lanes 3, 4. Returns the
// five components (x, y, z, theta, phi); further lanes of the point get 0.
template<int L>
Pt<L> bending_force(const Pt<L>& Xi, const Pt<L>& r, float dist)
{
    const Polarity pol_i{Xi.v[3], Xi.v[4]};
    const F3 pi = pol_to_float3(pol_i);
    const float prodi = (pi.x * r.v[0] + pi.y * r.v[1] + pi.z * r.v[2]) / dist;
    const Polarity r_hat = pt_to_pol(r.v[0], r.v[1], r.v[2], dist);
    const Polarity uni = unidirectional_polarization_force(pol_i, r_hat);
    Pt<L> dF = zero<L>();
    dF.v[3] = -prodi * uni.theta;
    dF.v[4] = -prodi * uni.phi;

    dF.v[0] = -prodi / dist * pi.x + powf(prodi, 2) / powf(dist, 2) * r.v[0];
    dF.v[1] = -prodi / dist * pi.y + powf(prodi, 2) / powf(dist, 2) * r.v[1];
    dF.v[2] = -prodi / dist * pi.z + powf(prodi, 2) / powf(dist, 2) * r.v[2];

    const Polarity pol_j{Xi.v[3] - r.v[3], Xi.v[4] - r.v[4]};
    const F3 pj = pol_to_float3(pol_j);
    const float prodj = (pj.x * r.v[0] + pj.y * r.v[1] + pj.z * r.v[2]) / dist;
    dF.v[0] += -prodj / dist * pj.x + powf(prodj, 2) / powf(dist, 2) * r.v[0];
    dF.v[1] += -prodj / dist * pj.y + powf(prodj, 2) / powf(dist, 2) * r.v[1];
    dF.v[2] += -prodj / dist * pj.z + powf(prodj, 2) / powf(dist, 2) * r.v[2];
    return dF;
}

// ---- the neighbour grid (solvers.cuh:350-425) ------------------------------------------
struct Grid {
    int n = 0, grid_size = 0, n_cubes = 0;
    std::vector<int> cube_id, point_id, cube_start, cube_end;
    int nhood[27];

    void init(int gs)
    {
        grid_size = gs;
        n_cubes = gs * gs * gs;
        cube_start.assign(n_cubes, -1);
        cube_end.assign(n_cubes, -2);
        // solvers.cuh:472-484
        nhood[0] = -1;
        nhood[1] = 0;
        nhood[2] = 1;
        for (int i = 0; i < 3; i++) {
            nhood[i + 3] = nhood[i % 3] - gs;
            nhood[i + 6] = nhood[i % 3] + gs;
        }
        for (int i = 0; i < 9; i++) {
            nhood[i + 9] = nhood[i % 9] - gs * gs;
            nhood[i + 18] = nhood[i % 9] + gs * gs;
        }
    }

    // X: n points of `stride` floats, x y z first
    void build(const float* X, int stride, int n_points, float cube_size)
    {
        n = n_points;
        cube_id.resize(n);
        point_id.resize(n);
        std::vector<int> key(n);
        const int gs = grid_size;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            const float* p = X + size_t(i) * stride;
            // FP32 evaluation, as compute_cube_id does (solvers.cuh:357-360)
            const float fx = floorf(p[0] / cube_size) + gs / 2;
            const float fy = floorf(p[1] / cube_size) + gs / 2;
            const float fz = floorf(p[2] / cube_size) + gs / 2;
            const float fid = fx + fy * gs + fz * gs * gs;
            int id = static_cast<int>(fid);
            if (id < 0) id = 0;  // the reference asserts here
            if (id >= n_cubes) id = n_cubes - 1;
            key[i] = id;
        }
        // thrust::sort_by_key on int keys is a STABLE sort of (key, identity
        // permutation): equal keys keep ascending original index. A counting
        // sort that places cells in index order produces exactly that.
        std::vector<int> cursor(n_cubes + 1, 0);
        for (int i = 0; i < n; i++) cursor[key[i] + 1]++;
        for (int c = 0; c < n_cubes; c++) cursor[c + 1] += cursor[c];
        for (int i = 0; i < n; i++) point_id[cursor[key[i]]++] = i;
        for (int k = 0; k < n; k++) cube_id[k] = key[point_id[k]];
        std::fill(cube_start.begin(), cube_start.end(), -1);
        std::fill(cube_end.begin(), cube_end.end(), -2);
        // solvers.cuh:367-378
        for (int k = 0; k < n; k++) {
            const int cube = cube_id[k];
            const int prev = k > 0 ? cube_id[k - 1] : -1;
            if (cube != prev) cube_start[cube] = k;
            const int next = k < n - 1 ? cube_id[k + 1] : cube_id[k] + 1;
            if (cube != next) cube_end[cube] = k;
        }
    }
};

// ---- models ------------------------------------------------------------------------------
enum Friction { FRICTION_W_NEIGHBOUR, FRICTION_ON_BACKGROUND };
enum Solver { TILE, GRID };

struct Params {
    float spring_length = 0.5f;
    float link_strength = 0.2f;
    float prolif_rate = 0.006f;
    std::vector<int> type, mes_nbs, epi_nbs;
};

template<int L>
using Force = Pt<L> (*)(
    const Pt<L>& Xi, const Pt<L>& r, float dist, int i, int j, Params& p);

// examples/springs.cu:14-21
Pt<3> spring(const Pt<3>& Xi, const Pt<3>& r, float dist, int i, int j, Params& p)
{
    if (i == j) return zero<3>();
    return r * (p.spring_length - dist) / dist;
}
// tests/test_solvers.cu:44-53
Pt<3> clipped_spring(
    const Pt<3>& Xi, const Pt<3>& r, float dist, int i, int j, Params& p)
{
    if (i == j) return zero<3>();
    if (dist >= 1) return zero<3>();
    return r * (p.spring_length - dist) / dist;
}
// include/inits.cuh:78-93
Pt<3> relu_force(
    const Pt<3>& Xi, const Pt<3>& r, float dist, int i, int j, Params& p)
{
    Pt<3> dF = zero<3>();
    if (i == j) return dF;
    if (dist > 1.f) return dF;
    const float F = fmaxf(0.8f - dist, 0) * 2.f - fmaxf(dist - 0.8f, 0);
    dF.v[0] = r.v[0] * F / dist;
    dF.v[1] = r.v[1] * F / dist;
    dF.v[2] = r.v[2] * F / dist;
    return dF;
}
// The examples write the ReLU with double literals: fmaxf(0.7 - dist, 0) narrows
// the double difference to float (examples/epithelium.cu:24).
inline float relu_d(double repulsive_below, double attractive_above, float dist)
{
    return fmaxf(repulsive_below - dist, 0) * 2 -
           fmaxf(dist - attractive_above, 0);
}
// examples/epithelium.cu:16-31
Pt<5> layer_force(
    const Pt<5>& Xi, const Pt<5>& r, float dist, int i, int j, Params& p)
{
    Pt<5> dF = zero<5>();
    if (i == j) return dF;
    if (dist > 1) return dF;
    const float F = relu_d(0.7, 0.8, dist);
    dF.v[0] = r.v[0] * F / dist;
    dF.v[1] = r.v[1] * F / dist;
    dF.v[2] = r.v[2] * F / dist;
    dF += bending_force(Xi, r, dist) * 0.2f;
    return dF;
}
// examples/passive_growth.cu:29-57
Pt<5> relu_w_epithelium(
    const Pt<5>& Xi, const Pt<5>& r, float dist, int i, int j, Params& p)
{
    Pt<5> dF = zero<5>();
    if (i == j) return dF;
    if (dist > 1) return dF;
    const float F = p.type[i] == p.type[j] ? relu_d(0.7, 0.8, dist)
                                           : relu_d(0.8, 0.9, dist);
    dF.v[0] = r.v[0] * F / dist;
    dF.v[1] = r.v[1] * F / dist;
    dF.v[2] = r.v[2] * F / dist;
    if (p.type[j] == 0)
        p.mes_nbs[i] += 1;
    else
        p.epi_nbs[i] += 1;
    if (p.type[i] == 0 || p.type[j] == 0) return dF;
    dF += bending_force(Xi, r, dist) * 0.15f;
    return dF;
}
// examples/branching.cu:60-110 (constants :21-31; they are doubles there)
Pt<7> epi_turing_mes_noturing(
    const Pt<7>& Xi, const Pt<7>& r, float dist, int i, int j, Params& p)
{
    const double lambda = 0.0075, D_u = 0.001, D_v = 0.2, f_v = 1.0, f_u = 80.0,
                 g_u = 80.0, m_u = 0.25, m_v = 0.75, s_u = 0.05;
    Pt<7> dF = zero<7>();
    const float u = Xi.v[5], v = Xi.v[6];
    if (i == j) {
        if (p.type[i] == 1) {
            dF.v[5] = lambda * ((f_u * u * u) / (1 + f_v * v) - m_u * u + s_u);
            dF.v[6] = lambda * (g_u * u * u - m_v * v);
            if (-dF.v[5] > u) dF.v[5] = 0.0f;
            if (-dF.v[6] > v) dF.v[6] = 0.0f;
        }
        return dF;
    }
    if (dist > 1.0f) return dF;
    const float F = p.type[i] == p.type[j] ? relu_d(0.7, 0.8, dist)
                                           : relu_d(0.8, 0.9, dist);
    dF.v[0] = r.v[0] * F / dist;
    dF.v[1] = r.v[1] * F / dist;
    dF.v[2] = r.v[2] * F / dist;
    if (p.type[i] == 1 && p.type[j] == 1) {
        dF.v[5] = -D_u * r.v[5];
        dF.v[6] = -D_v * r.v[6];
        if (-dF.v[5] > u) dF.v[5] = 0.0f;
        if (-dF.v[6] > v) dF.v[6] = 0.0f;
        dF += bending_force(Xi, r, dist) * 0.2f;
    } else {
        dF.v[6] = -D_v * r.v[6];
    }
    if (p.type[j] == 1)
        p.epi_nbs[i] += 1;
    else
        p.mes_nbs[i] += 1;
    return dF;
}

inline float friction_of(Friction f, float dist, int i, int j)
{
    if (f == FRICTION_ON_BACKGROUND) return 0.f;
    if (i == j) return 0.f;  // solvers.cuh:27-35
    return dist < 1 ? 1.f : 0.f;
}

}  // namespace


struct yb_sim {
    virtual ~yb_sim() {}
    virtual int lanes() const = 0;
    virtual int n_max() const = 0;
    virtual int set_param(const std::string&, double) = 0;
    virtual int set_state(const float*, int, int) = 0;
    virtual int get_state(float*, int, int*) = 0;
    virtual int get_velocities(float*, int) = 0;
    virtual int set_velocities(const float*, int) = 0;
    virtual int set_ints(const std::string&, const int*, int) = 0;
    virtual int get_ints(const std::string&, int*, int) = 0;
    virtual int set_links(const int*, int) = 0;
    virtual int step(float dt) = 0;
    virtual int current_n() const = 0;
    virtual int dd_load(int, const float*, const float*, int, const float*,
        const float*, int) = 0;
    virtual int dd_forces(int, float*) = 0;
    virtual int dd_update(int, float, const float*) = 0;
    virtual int dd_read(int, float*, int) = 0;
    virtual int slab_begin(float, float, float, int) = 0;
    virtual int slab_set_owned(const float*, const float*, int) = 0;
    virtual int slab_pack(int, float*, float*) = 0;
    virtual int slab_unpack(int, const float*, const float*) = 0;
    virtual int slab_update(int, float, const float*) = 0;
    virtual int slab_counts(int*, int*, int*) = 0;
};

namespace {

template<int L>
struct Sim : yb_sim {
    int capacity, n = 0;
    Solver solver;
    Friction friction;
    Force<L> force;
    bool counts_neighbours = false, has_links = false, grows = false;
    float cube_size;
    Grid grid;
    Params params;
    std::vector<Pt<L>> X, X1, dX, dX1;
    std::vector<F3> old_v, sum_v;
    std::vector<float> sum_friction;
    std::vector<int> links;  // pairs
    bool fix_com = true, fix_com_z = false;
    int fix_point = 0;
    int n_owned = 0;  // domain decomposition: cells >= n_owned are ghosts
    float slab_z_lo = 0, slab_z_hi = 0, slab_halo = 0;
    int slab_capacity = 0;
    bool has_lower = false, has_upper = false;
    std::vector<int> stayers;

    Sim(int n_max, Solver solver, Force<L> force, Friction friction,
        int grid_size, float cube_size)
        : capacity{n_max}, solver{solver}, friction{friction}, force{force},
          cube_size{cube_size}, X(n_max), X1(n_max), dX(n_max), dX1(n_max),
          old_v(n_max, F3{0, 0, 0}), sum_v(n_max), sum_friction(n_max)
    {
        if (solver == GRID) grid.init(grid_size);
        params.type.assign(n_max, 0);
        params.mes_nbs.assign(n_max, 0);
        params.epi_nbs.assign(n_max, 0);
    }
    int lanes() const override { return L; }
    int n_max() const override { return capacity; }
    int current_n() const override { return n; }

    int set_param(const std::string& name, double value) override
    {
        if (name == "spring_length")
            params.spring_length = float(value);
        else if (name == "link_strength")
            params.link_strength = float(value);
        else if (name == "prolif_rate")
            params.prolif_rate = float(value);
        else if (name == "mean_dist" || name == "seed")
            ;
        else if (name == "fix_point") {  // solvers.cuh:197-201
            fix_com = false;
            fix_point = int(value);
        } else if (name == "fix_point_xy") {  // :203-208
            fix_com = false;
            fix_com_z = true;
            fix_point = int(value);
        } else if (name == "fix_com")
            fix_com = true;
        else
            return fail(YB_EINVAL, "unknown parameter " + name);
        return YB_OK;
    }
    int set_state(const float* h_X, int n_new, int reset_v) override
    {
        if (n_new < 0 || n_new > capacity) return fail(YB_EINVAL, "n > n_max");
        memcpy(X.data(), h_X, sizeof(Pt<L>) * size_t(n_new));
        n = n_new;
        if (reset_v) std::fill(old_v.begin(), old_v.end(), F3{0, 0, 0});
        return YB_OK;
    }
    int get_state(float* h_X, int cap, int* n_out) override
    {
        if (n > cap) return fail(YB_EINVAL, "capacity < n");
        memcpy(h_X, X.data(), sizeof(Pt<L>) * size_t(n));
        if (n_out) *n_out = n;
        return YB_OK;
    }
    int get_velocities(float* h_v, int cap) override
    {
        if (n > cap) return fail(YB_EINVAL, "capacity < n");
        memcpy(h_v, old_v.data(), sizeof(F3) * size_t(n));
        return YB_OK;
    }
    int set_velocities(const float* h_v, int count) override
    {
        if (count < 0 || count > capacity) return fail(YB_EINVAL, "n > n_max");
        memcpy(old_v.data(), h_v, sizeof(F3) * size_t(count));
        return YB_OK;
    }
    int set_ints(const std::string& name, const int* values, int count) override
    {
        if (name != "type" || !counts_neighbours)
            return fail(YB_EINVAL, "model has no int property " + name);
        if (count > capacity) return fail(YB_EINVAL, "n > n_max");
        std::copy(values, values + count, params.type.begin());
        return YB_OK;
    }
    int get_ints(const std::string& name, int* values, int cap) override
    {
        if (!counts_neighbours)
            return fail(YB_EINVAL, "model has no int property " + name);
        if (n > cap) return fail(YB_EINVAL, "capacity < n");
        const std::vector<int>* src = nullptr;
        if (name == "type") src = &params.type;
        if (name == "mes_nbs") src = &params.mes_nbs;
        if (name == "epi_nbs") src = &params.epi_nbs;
        if (!src) return fail(YB_EINVAL, "model has no int property " + name);
        std::copy(src->begin(), src->begin() + n, values);
        return YB_OK;
    }
    int set_links(const int* h_links, int n_links) override
    {
        if (!has_links) return fail(YB_EINVAL, "model has no links");
        links.assign(h_links, h_links + 2 * size_t(n_links));
        return YB_OK;
    }

    // links.cuh:99-125, sequentially in link order
    void link_forces(const std::vector<Pt<L>>& P, std::vector<Pt<L>>& dP)
    {
        const float strength = params.link_strength;
        for (size_t l = 0; l + 1 < links.size(); l += 2) {
            const int a = links[l], b = links[l + 1];
            if (a == b) continue;
            const Pt<L> r = P[a] - P[b];
            const float dist = norm3(r.v[0], r.v[1], r.v[2]);
            for (int c = 0; c < 3; c++) {
                dP[a].v[c] += -strength * r.v[c] / dist;
                dP[b].v[c] += strength * r.v[c] / dist;
            }
        }
    }

    // One evaluation of the right-hand side: solvers.cuh:232-238 (258-264)
    void derivative(const std::vector<Pt<L>>& P, std::vector<Pt<L>>& dP)
    {
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            dP[i] = zero<L>();
            sum_friction[i] = 0;
            sum_v[i] = F3{0, 0, 0};
        }
        // generic forces
        if (counts_neighbours) {
            std::fill(params.mes_nbs.begin(), params.mes_nbs.begin() + n, 0);
            std::fill(params.epi_nbs.begin(), params.epi_nbs.begin() + n, 0);
        }
        if (has_links) link_forces(P, dP);

        if (solver == TILE) {
            // solvers.cuh:286-322: all pairs incl. self, j ascending
#pragma omp parallel for schedule(static)
            for (int i = 0; i < n; i++) {
                Pt<L> F = zero<L>();
                F3 sv{0, 0, 0};
                float sf = 0;
                for (int j = 0; j < n; j++) {
                    const Pt<L> r = P[i] - P[j];
                    const float dist = norm3(r.v[0], r.v[1], r.v[2]);
                    F += force(P[i], r, dist, i, j, params);
                    const float fr = friction_of(friction, dist, i, j);
                    sf += fr;
                    sv.x += fr * old_v[j].x;
                    sv.y += fr * old_v[j].y;
                    sv.z += fr * old_v[j].z;
                }
                dP[i] += F;
                sum_friction[i] = sf;
                sum_v[i] = sv;
            }
        } else {
            grid.build(&P[0].v[0], L, n, cube_size);
            // solvers.cuh:432-463: one sorted slot per thread
#pragma omp parallel for schedule(dynamic, 256)
            for (int s = 0; s < n; s++) {
                const int i = grid.point_id[s];
                if (n_owned > 0 && i >= n_owned) continue;  // ghost
                const Pt<L> Xi = P[i];
                Pt<L> F = zero<L>();
                F3 sv{0, 0, 0};
                float sf = 0;
                for (int q = 0; q < 27; q++) {
                    const int cube = grid.cube_id[s] + grid.nhood[q];
                    if (cube < 0 || cube >= grid.n_cubes) continue;  // UB in ref
                    for (int k = grid.cube_start[cube]; k <= grid.cube_end[cube];
                         k++) {
                        const int j = grid.point_id[k];
                        const Pt<L> r = Xi - P[j];
                        const float dist = norm3(r.v[0], r.v[1], r.v[2]);
                        if (dist >= cube_size) continue;
                        F += force(Xi, r, dist, i, j, params);
                        const float fr = friction_of(friction, dist, i, j);
                        sf += fr;
                        sv.x += fr * old_v[j].x;
                        sv.y += fr * old_v[j].y;
                        sv.z += fr * old_v[j].z;
                    }
                }
                dP[i] += F;
                sum_v[i] = sv;
                sum_friction[i] = sf;
            }
        }
        // add_rhs, solvers.cuh:147-161
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            if (sum_friction[i] > 0) {
                dP[i].v[0] += sum_v[i].x / sum_friction[i];
                dP[i].v[1] += sum_v[i].y / sum_friction[i];
                dP[i].v[2] += sum_v[i].z / sum_friction[i];
            }
        }
    }

    // thrust::reduce(dX) / n over all members (solvers.cuh:242); the summation
    // order of the device reduction is unspecified, double accumulation here
    Pt<L> mean(const std::vector<Pt<L>>& dP)
    {
        double acc[L] = {0};
        for (int i = 0; i < n; i++)
            for (int k = 0; k < L; k++) acc[k] += dP[i].v[k];
        Pt<L> total;
        for (int k = 0; k < L; k++) total.v[k] = float(acc[k]);
        return total / float(n);
    }

    // ---- domain decomposition building blocks (host pointers) ---------------
    int dd_load(int stage, const float* X_owned, const float* v_owned,
        int owned, const float* X_ghost, const float* v_ghost,
        int n_ghost) override
    {
        if (solver != GRID || counts_neighbours || has_links)
            return fail(YB_ENOSYS, "domain decomposition needs a plain Grid model");
        if (owned + n_ghost > capacity) return fail(YB_EINVAL, "n > n_max");
        std::vector<Pt<L>>& P = stage == 0 ? X : X1;
        if (stage == 0) {
            memcpy(P.data(), X_owned, sizeof(Pt<L>) * size_t(owned));
            memcpy(old_v.data(), v_owned, sizeof(F3) * size_t(owned));
        }
        if (n_ghost > 0) {
            memcpy(P.data() + owned, X_ghost, sizeof(Pt<L>) * size_t(n_ghost));
            memcpy(old_v.data() + owned, v_ghost, sizeof(F3) * size_t(n_ghost));
        }
        n_owned = owned;
        n = owned + n_ghost;
        return YB_OK;
    }
    int dd_forces(int stage, float* sums4) override
    {
        std::vector<Pt<L>>& dP = stage == 0 ? dX : dX1;
        derivative(stage == 0 ? X : X1, dP);
        double acc[3] = {0, 0, 0};
        for (int i = 0; i < n_owned; i++)
            for (int k = 0; k < 3; k++) acc[k] += dP[i].v[k];
        for (int k = 0; k < 3; k++) sums4[k] = float(acc[k]);
        sums4[3] = float(n_owned);
        return YB_OK;
    }
    int dd_update(int stage, float dt, const float* mean3) override
    {
        if (stage == 0) {
            for (int i = 0; i < n_owned; i++) {
                for (int k = 0; k < 3; k++) dX[i].v[k] -= mean3[k];
                X1[i] = X[i] + dX[i] * dt;
            }
        } else {
            for (int i = 0; i < n_owned; i++) {
                for (int k = 0; k < 3; k++) dX1[i].v[k] -= mean3[k];
                X[i] += (dX[i] + dX1[i]) * 0.5f * dt;
                old_v[i].x = (dX[i].v[0] + dX1[i].v[0]) * 0.5f;
                old_v[i].y = (dX[i].v[1] + dX1[i].v[1]) * 0.5f;
                old_v[i].z = (dX[i].v[2] + dX1[i].v[2]) * 0.5f;
            }
        }
        return YB_OK;
    }
    int dd_read(int which, float* out, int count) override
    {
        if (count > capacity) return fail(YB_EINVAL, "n > n_max");
        if (which == 0)
            memcpy(out, X.data(), sizeof(Pt<L>) * size_t(count));
        else if (which == 1)
            memcpy(out, X1.data(), sizeof(Pt<L>) * size_t(count));
        else
            memcpy(out, old_v.data(), sizeof(F3) * size_t(count));
        return YB_OK;
    }

    // ---- slab API: same semantics as the device kernels in b200/slab.cuh,
    //      written as plain index-order loops -----------------------------------
    static int header_count(const float* buffer)
    {
        int count;
        memcpy(&count, buffer, sizeof(int));
        return count;
    }
    static void set_header(float* buffer, int count)
    {
        memcpy(buffer, &count, sizeof(int));
    }
    void put_record(float* record, const std::vector<Pt<L>>& P, int i) const
    {
        memcpy(record, P[i].v, sizeof(float) * L);
        record[L + 0] = old_v[i].x;
        record[L + 1] = old_v[i].y;
        record[L + 2] = old_v[i].z;
    }
    void get_record(const float* record, std::vector<Pt<L>>& P, int i)
    {
        memcpy(P[i].v, record, sizeof(float) * L);
        old_v[i] = F3{record[L + 0], record[L + 1], record[L + 2]};
    }
    int slab_begin(float z_lo, float z_hi, float halo, int cap) override
    {
        if (solver != GRID || counts_neighbours || has_links)
            return fail(YB_ENOSYS, "domain decomposition needs a plain Grid model");
        slab_z_lo = z_lo, slab_z_hi = z_hi, slab_halo = halo, slab_capacity = cap;
        has_lower = std::isfinite(z_lo), has_upper = std::isfinite(z_hi);
        n_owned = n = 0;
        return YB_OK;
    }
    int slab_set_owned(const float* X_new, const float* v_new, int owned) override
    {
        if (owned > capacity) return fail(YB_EINVAL, "n_owned > n_max");
        memcpy(X.data(), X_new, sizeof(Pt<L>) * size_t(owned));
        memcpy(old_v.data(), v_new, sizeof(F3) * size_t(owned));
        n_owned = n = owned;
        return YB_OK;
    }
    int slab_pack(int what, float* send_lo, float* send_hi) override
    {
        const bool migration = what == 2;
        const std::vector<Pt<L>>& P = what == 1 ? X1 : X;
        const float lo_edge = migration ? slab_z_lo : slab_z_lo + slab_halo;
        const float hi_edge = migration ? slab_z_hi : slab_z_hi - slab_halo;
        const int W = L + 3;
        int n_lo = 0, n_hi = 0;
        stayers.clear();
        for (int i = 0; i < n_owned; i++) {
            const float z = P[i].v[2];
            const bool lo = has_lower && z < lo_edge;
            const bool hi = has_upper && z >= hi_edge;
            if (lo && n_lo < slab_capacity)
                put_record(send_lo + 4 + size_t(n_lo++) * W, P, i);
            if (hi && n_hi < slab_capacity)
                put_record(send_hi + 4 + size_t(n_hi++) * W, P, i);
            if (migration && !(lo || hi)) stayers.push_back(i);
        }
        set_header(send_lo, n_lo);
        set_header(send_hi, n_hi);
        return YB_OK;
    }
    int slab_unpack(int what, const float* recv_lo, const float* recv_hi) override
    {
        const int W = L + 3;
        const int n_lo = has_lower ? header_count(recv_lo) : 0;
        const int n_hi = has_upper ? header_count(recv_hi) : 0;
        int base = n_owned;
        std::vector<Pt<L>>& P = what == 1 ? X1 : X;
        if (what == 2) {
            for (size_t k = 0; k < stayers.size(); k++) {  // stable, in place
                X[k] = X[stayers[k]];
                old_v[k] = old_v[stayers[k]];
            }
            base = int(stayers.size());
        }
        if (base + n_lo + n_hi > capacity) return fail(YB_EINVAL, "n > n_max");
        for (int r = 0; r < n_lo; r++)
            get_record(recv_lo + 4 + size_t(r) * W, P, base + r);
        for (int r = 0; r < n_hi; r++)
            get_record(recv_hi + 4 + size_t(r) * W, P, base + n_lo + r);
        n = base + n_lo + n_hi;
        if (what == 2) n_owned = n;
        return YB_OK;
    }
    int slab_update(int stage, float dt, const float* sums4) override
    {
        const float inv_n = static_cast<float>(1. / sums4[3]);
        const float mean[3] = {sums4[0] * inv_n, sums4[1] * inv_n, sums4[2] * inv_n};
        return dd_update(stage, dt, mean);
    }
    int slab_counts(int* owned, int* total, int* problems) override
    {
        if (owned) *owned = n_owned;
        if (total) *total = n;
        if (problems) *problems = 0;
        return YB_OK;
    }

    int step(float dt) override
    {
        n_owned = 0;
        if (grows && params.prolif_rate > 0)
            return fail(YB_ENOSYS, "the oracle has no curand: set prolif_rate 0");
        if (n == 0) return YB_OK;
        // 1st stage, solvers.cuh:231-255
        derivative(X, dX);
        Pt<L> fix;
        if (fix_com || fix_com_z) {
            fix = mean(dX);
            if (fix_com_z) {
                fix.v[0] = dX[fix_point].v[0];
                fix.v[1] = dX[fix_point].v[1];
            }
        } else {
            fix = dX[fix_point];
        }
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            dX[i].v[0] -= fix.v[0];
            dX[i].v[1] -= fix.v[1];
            dX[i].v[2] -= fix.v[2];
            X1[i] = X[i] + dX[i] * dt;
        }
        // 2nd stage, solvers.cuh:257-274
        derivative(X1, dX1);
        const Pt<L> fix1 = fix_com ? mean(dX1) : dX1[fix_point];
#pragma omp parallel for schedule(static)
        for (int i = 0; i < n; i++) {
            dX1[i].v[0] -= fix1.v[0];
            dX1[i].v[1] -= fix1.v[1];
            dX1[i].v[2] -= fix1.v[2];
            X[i] += (dX[i] + dX1[i]) * 0.5f * dt;
            old_v[i].x = (dX[i].v[0] + dX1[i].v[0]) * 0.5f;
            old_v[i].y = (dX[i].v[1] + dX1[i].v[1]) * 0.5f;
            old_v[i].z = (dX[i].v[2] + dX1[i].v[2]) * 0.5f;
        }
        return YB_OK;
    }
};

template<int L>
int grid_build_host(const float* X, int n, int grid_size, float cube_size,
    int* cube_id, int* point_id, int* cube_start, int* cube_end)
{
    Grid grid;
    grid.init(grid_size);
    grid.build(X, L, n, cube_size);
    std::copy(grid.cube_id.begin(), grid.cube_id.end(), cube_id);
    std::copy(grid.point_id.begin(), grid.point_id.end(), point_id);
    std::copy(grid.cube_start.begin(), grid.cube_start.end(), cube_start);
    std::copy(grid.cube_end.begin(), grid.cube_end.end(), cube_end);
    return YB_OK;
}

}  // namespace


extern "C" {

const char* yb_build_info(void)
{
    return "yalla-oracle (CPU restatement of the reference; test infrastructure)";
}
const char* yb_last_error(void) { return last_error.c_str(); }

int yb_sim_create(const char* model, int n_max, int grid_size, float cube_size,
    yb_sim** out)
{
    if (!model || !out || n_max <= 0) return fail(YB_EINVAL, "bad argument");
    const std::string name = model;
    if (grid_size <= 0) grid_size = 50;
    if (!(cube_size > 0)) cube_size = 1.f;
    const Friction w_nb = FRICTION_W_NEIGHBOUR;
    if (name == "springs") {
        *out = new Sim<3>(n_max, TILE, spring, w_nb, grid_size, cube_size);
    } else if (name == "spring_tile") {
        *out = new Sim<3>(n_max, TILE, clipped_spring, w_nb, grid_size, cube_size);
    } else if (name == "spring_grid") {
        *out = new Sim<3>(n_max, GRID, clipped_spring, w_nb, grid_size, cube_size);
    } else if (name == "relu_tile") {
        *out = new Sim<3>(n_max, TILE, relu_force, w_nb, grid_size, cube_size);
    } else if (name == "relu_grid") {
        *out = new Sim<3>(n_max, GRID, relu_force, w_nb, grid_size, cube_size);
    } else if (name == "protrusions") {
        auto* sim = new Sim<3>(n_max, GRID, relu_force, w_nb, grid_size, cube_size);
        sim->has_links = true;
        *out = sim;
    } else if (name == "epithelium") {
        *out = new Sim<5>(n_max, GRID, layer_force, FRICTION_ON_BACKGROUND,
            grid_size, cube_size);
    } else if (name == "growth") {
        auto* sim =
            new Sim<5>(n_max, GRID, relu_w_epithelium, w_nb, grid_size, cube_size);
        sim->counts_neighbours = true;
        sim->grows = true;
        *out = sim;
    } else if (name == "branching") {
        auto* sim = new Sim<7>(
            n_max, GRID, epi_turing_mes_noturing, w_nb, grid_size, cube_size);
        sim->counts_neighbours = true;
        *out = sim;
    } else {
        return fail(YB_EINVAL, "unknown model " + name);
    }
    return YB_OK;
}

void yb_sim_destroy(yb_sim* sim) { delete sim; }
int yb_sim_lanes(const yb_sim* sim) { return sim->lanes(); }
int yb_sim_n_max(const yb_sim* sim) { return sim->n_max(); }
int yb_sim_set_param(yb_sim* sim, const char* name, double value)
{
    return sim->set_param(name, value);
}
int yb_sim_set_state(yb_sim* sim, const float* h_X, int n, int reset_v)
{
    return sim->set_state(h_X, n, reset_v);
}
int yb_sim_get_state(yb_sim* sim, float* h_X, int capacity, int* n_out)
{
    return sim->get_state(h_X, capacity, n_out);
}
int yb_sim_get_velocities(yb_sim* sim, float* h_v, int capacity)
{
    return sim->get_velocities(h_v, capacity);
}
int yb_sim_set_velocities(yb_sim* sim, const float* h_v, int n)
{
    return sim->set_velocities(h_v, n);
}
int yb_sim_set_ints(yb_sim* sim, const char* name, const int* h_values, int n)
{
    return sim->set_ints(name, h_values, n);
}
int yb_sim_get_ints(yb_sim* sim, const char* name, int* h_values, int capacity)
{
    return sim->get_ints(name, h_values, capacity);
}
int yb_sim_seed_sphere(yb_sim*, int, float, unsigned long long, int)
{
    return fail(YB_ENOSYS, "seeded device generators exist in the product library only");
}

int yb_sim_set_links(yb_sim* sim, const int* h_links, int n_links)
{
    return sim->set_links(h_links, n_links);
}
int yb_sim_step(yb_sim* sim, float dt, int n_steps)
{
    for (int k = 0; k < n_steps; k++) {
        const int status = sim->step(dt);
        if (status != YB_OK) return status;
    }
    return YB_OK;
}
int yb_sim_step_timed(yb_sim* sim, float dt, int n_steps, float* ms_out,
    long long* cell_updates_out)
{
    const auto start = std::chrono::steady_clock::now();
    long long updates = 0;
    for (int k = 0; k < n_steps; k++) {
        updates += sim->current_n();
        const int status = sim->step(dt);
        if (status != YB_OK) return status;
    }
    const auto stop = std::chrono::steady_clock::now();
    if (ms_out)
        *ms_out = std::chrono::duration<float, std::milli>(stop - start).count();
    if (cell_updates_out) *cell_updates_out = updates;
    return YB_OK;
}
int yb_sim_step_host(yb_sim* sim, const float* h_in, int n, float dt,
    int n_steps, float* h_out, int capacity, int* n_out)
{
    int status = sim->set_state(h_in, n, 0);
    if (status != YB_OK) return status;
    status = yb_sim_step(sim, dt, n_steps);
    if (status != YB_OK) return status;
    return sim->get_state(h_out, capacity, n_out);
}
int yb_dd_load(yb_sim* sim, int stage, const float* X_owned,
    const float* v_owned, int n_owned, const float* X_ghost,
    const float* v_ghost, int n_ghost)
{
    return sim->dd_load(stage, X_owned, v_owned, n_owned, X_ghost, v_ghost, n_ghost);
}
int yb_dd_forces(yb_sim* sim, int stage, float* sums4)
{
    return sim->dd_forces(stage, sums4);
}
int yb_dd_update(yb_sim* sim, int stage, float dt, const float* mean3)
{
    return sim->dd_update(stage, dt, mean3);
}
int yb_dd_read(yb_sim* sim, int which, float* out, int n)
{
    return sim->dd_read(which, out, n);
}
int yb_slab_begin(yb_sim* sim, float z_lo, float z_hi, float halo,
    int capacity, int, int)
{
    return sim->slab_begin(z_lo, z_hi, halo, capacity);
}
int yb_slab_set_owned(yb_sim* sim, const float* X, const float* v, int n_owned)
{
    return sim->slab_set_owned(X, v, n_owned);
}
int yb_slab_pack(yb_sim* sim, int what, float* send_lo, float* send_hi)
{
    return sim->slab_pack(what, send_lo, send_hi);
}
int yb_slab_unpack(yb_sim* sim, int what, const float* recv_lo,
    const float* recv_hi)
{
    return sim->slab_unpack(what, recv_lo, recv_hi);
}
int yb_slab_update(yb_sim* sim, int stage, float dt, const float* sums4)
{
    return sim->slab_update(stage, dt, sums4);
}
int yb_slab_counts(yb_sim* sim, int* n_owned, int* n_total, int* problems)
{
    return sim->slab_counts(n_owned, n_total, problems);
}
int yb_sim_set_stream(yb_sim*, void*)
{
    return fail(YB_ENOSYS, "streams need the product library");
}
int yb_sim_step_host_async(
    yb_sim*, const float*, int, float, int, float*, int, int*)
{
    return fail(YB_ENOSYS, "asynchronous steps need the product library");
}
int yb_sim_get_links(yb_sim*, int*, int, int*)
{
    return fail(YB_ENOSYS, "reading links back: GPU libraries only");
}
int yb_dom_begin(yb_sim*, int, int, const float*, const float*, float,
    const int*, const int*, const int*, const int*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_register_array(yb_sim*, void*, int, int)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_exchange(yb_sim*, void**, long long*, long long*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_connect(yb_sim*, int, void*, const long long*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_connect_mailbox(yb_sim*, int, void*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_seed_lattice_ball(yb_sim*, float, float, float, unsigned long long, int*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_step(yb_sim*, float, int)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_dom_read_profile(yb_sim*, float*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_ipc_export(const void*, unsigned char*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_ipc_import(const unsigned char*, void**)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_ipc_release(void*)
{
    return fail(YB_ENOSYS, "peer-memory decomposition needs the product library");
}
int yb_sim_host_drain(yb_sim*)
{
    return fail(YB_ENOSYS, "asynchronous steps need the product library");
}
int yb_sim_profile_sweeps(yb_sim*, int)
{
    return fail(YB_ENOSYS, "sweep profiling needs the product library");
}
int yb_sim_read_sweep_profile(yb_sim*, float*, int*)
{
    return fail(YB_ENOSYS, "sweep profiling needs the product library");
}
int yb_sim_n(yb_sim* sim, int* n_out)
{
    *n_out = sim->current_n();
    return YB_OK;
}
int yb_sim_sync(yb_sim*) { return YB_OK; }

// In the oracle every "device" pointer is a host pointer.
int yb_grid_build(const float* X, int lanes, int n, int grid_size,
    float cube_size, int* cube_id, int* point_id, int* cube_start, int* cube_end)
{
    if (n <= 0 || grid_size <= 0) return fail(YB_EINVAL, "bad size");
    switch (lanes) {
        case 3:
            return grid_build_host<3>(X, n, grid_size, cube_size, cube_id,
                point_id, cube_start, cube_end);
        case 4:
            return grid_build_host<4>(X, n, grid_size, cube_size, cube_id,
                point_id, cube_start, cube_end);
        case 5:
            return grid_build_host<5>(X, n, grid_size, cube_size, cube_id,
                point_id, cube_start, cube_end);
        case 7:
            return grid_build_host<7>(X, n, grid_size, cube_size, cube_id,
                point_id, cube_start, cube_end);
    }
    return fail(YB_EINVAL, "lanes must be 3, 4, 5 or 7");
}

int yb_nhood(int grid_size, int* h_nhood27)
{
    Grid grid;
    grid.init(grid_size);
    std::copy(grid.nhood, grid.nhood + 27, h_nhood27);
    return YB_OK;
}

int yb_link_forces(const float* X, float* dX, int lanes, int n,
    const int* links, int n_links, float strength)
{
    if (n <= 0 || n_links <= 0) return fail(YB_EINVAL, "bad size");
    for (int l = 0; l < n_links; l++) {
        const int a = links[2 * l], b = links[2 * l + 1];
        if (a == b) continue;
        float r[3];
        for (int c = 0; c < 3; c++)
            r[c] = X[size_t(a) * lanes + c] + -1.f * X[size_t(b) * lanes + c];
        const float dist = norm3(r[0], r[1], r[2]);
        for (int c = 0; c < 3; c++) {
            dX[size_t(a) * lanes + c] += -strength * r[c] / dist;
            dX[size_t(b) * lanes + c] += strength * r[c] / dist;
        }
    }
    return YB_OK;
}

int yb_bending_force(
    const float* h_Xi, const float* h_Xj, int n_pairs, float* h_out)
{
    for (int k = 0; k < n_pairs; k++) {
        Pt<5> Xi, Xj;
        memcpy(Xi.v, h_Xi + 5 * size_t(k), sizeof(Xi.v));
        memcpy(Xj.v, h_Xj + 5 * size_t(k), sizeof(Xj.v));
        const Pt<5> r = Xi - Xj;
        const float dist = norm3(r.v[0], r.v[1], r.v[2]);
        const Pt<5> dF = bending_force(Xi, r, dist);
        memcpy(h_out + 5 * size_t(k), dF.v, sizeof(dF.v));
    }
    return YB_OK;
}

int yb_polarization_force(
    const float* h_Xi, const float* h_Xj, int n_pairs, float* h_out)
{
    for (int k = 0; k < n_pairs; k++) {
        const float* Xi = h_Xi + 5 * size_t(k);
        const float* Xj = h_Xj + 5 * size_t(k);
        const Polarity dF = bidirectional_polarization_force(
            Polarity{Xi[3], Xi[4]}, Polarity{Xj[3], Xj[4]});
        float* out = h_out + 5 * size_t(k);
        out[0] = out[1] = out[2] = 0.f;
        out[3] = dF.theta;
        out[4] = dF.phi;
    }
    return YB_OK;
}

}  // extern "C"
