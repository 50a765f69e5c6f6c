// GPU test program for header-level extensions that have no C-ABI entry:
// Vtk_async_output against the synchronous Vtk_output, the seeded generators
// through the header API, and Cell_division. Prints one "ok <name>" line per check
// and exits non-zero on the first failure (run by tests/test_extensions_gpu.py).
#include <stdio.h>
#include <stdlib.h>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>
#include <string.h>

#include "../../include/dtypes.cuh"
#include "../../include/inits.cuh"
#include "../../include/polarity.cuh"
#include "../../include/property.cuh"
#include "../../include/solvers.cuh"
#include "../../include/vtk.cuh"
#include "../../include/b200/division.cuh"
#include "../../include/b200/protrusions.cuh"

#define CHECK(cond, name)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            printf("FAILED %s (%s:%d)\n", name, __FILE__, __LINE__); \
            exit(1);                                        \
        }                                                   \
        printf("ok %s\n", name);                            \
    } while (0)

static std::string slurp(const std::string& path)
{
    std::ifstream file(path, std::ios::binary);
    return std::string(
        std::istreambuf_iterator<char>(file), std::istreambuf_iterator<char>());
}

__device__ Po_cell layer(Po_cell Xi, Po_cell r, float dist, int i, int j)
{
    Po_cell dF{0};
    if (i == j or dist > 1) return dF;
    const float F = fmaxf(0.7f - dist, 0.f) * 2.f - fmaxf(dist - 0.8f, 0.f);
    dF.x = r.x * F / dist, dF.y = r.y * F / dist, dF.z = r.z * F / dist;
    dF += bending_force(Xi, r, dist) * 0.2f;
    return dF;
}

// ---- Cell_division: a rate rule with a per-cell property to inherit -----------
__device__ int* d_lineage;

__device__ float divide_sometimes(int i, const Po_cell& X)
{
    return X.z > 0 ? 0.25f : 0.f;  // only the upper half of the tissue divides
}

__device__ void inherit_lineage(int mother, int daughter)
{
    d_lineage[daughter] = d_lineage[mother];
}

static void check_division()
{
    const int n = 50000, n_max = 80000;
    std::vector<Po_cell> first_run;
    std::vector<int> first_lineage;
    int first_n = 0;
    for (int run = 0; run < 2; run++) {
        Solution<Po_cell, Grid_solver> cells{n_max, 64, 1.f};
        Property<int> lineage{n_max, "lineage"};
        cudaMemcpyToSymbol(d_lineage, &lineage.d_prop, sizeof(d_lineage));
        *cells.h_n = n;
        for (int i = 0; i < n_max; i++) {
            cells.h_X[i] = Po_cell{0};
            lineage.h_prop[i] = i < n ? i : -1;
        }
        cells.copy_to_device();
        lineage.copy_to_device();
        seeded_sphere(0.8f, cells, 9);
        Cell_division<Po_cell> division{n_max, 1234};
        int upper = 0;
        for (int i = 0; i < n; i++) upper += cells.h_X[i].z > 0;

        division.divide<divide_sometimes, inherit_lineage>(cells, 0.8f);
        cells.copy_to_host();
        lineage.copy_to_host();
        const int n_1 = *cells.h_n;
        if (run == 0) {
            const double expected = 0.25 * upper, sigma = sqrt(0.25 * 0.75 * upper);
            CHECK(fabs((n_1 - n) - expected) < 5 * sigma, "division count");
            bool placed = true, inherited = true;
            for (int j = n; j < n_1; j++) {
                const int mother = lineage.h_prop[j];
                inherited = inherited && mother >= 0 && mother < n &&
                            cells.h_X[mother].z > 0;
                if (!inherited) break;
                const float dx = cells.h_X[j].x - cells.h_X[mother].x,
                            dy = cells.h_X[j].y - cells.h_X[mother].y,
                            dz = cells.h_X[j].z - cells.h_X[mother].z;
                placed = placed && fabsf(sqrtf(dx * dx + dy * dy + dz * dz) - 0.2f) < 1e-4f;
                // stable: daughters appear in the order of their mothers
                if (j > n) inherited = inherited && lineage.h_prop[j - 1] < mother;
            }
            CHECK(inherited, "daughters inherit and are appended in mother order");
            CHECK(placed, "daughters sit mean_dist / 4 from their mothers");
        }
        // second call: new draws (other mothers), and the tissue fills up
        division.divide<divide_sometimes, inherit_lineage>(cells, 0.8f);
        for (int k = 0; k < 6; k++)
            division.divide<divide_sometimes, inherit_lineage>(cells, 0.8f);
        cells.copy_to_host();
        lineage.copy_to_host();
        int before = 0, dropped = 0;
        division.last_call(&before, &dropped);
        if (run == 0) {
            CHECK(*cells.h_n == n_max && dropped > 0, "full tissue drops divisions");
            first_n = *cells.h_n;
            first_run.assign(cells.h_X, cells.h_X + first_n);
            first_lineage.assign(lineage.h_prop, lineage.h_prop + first_n);
        } else {
            bool same = *cells.h_n == first_n;
            for (int i = 0; same && i < first_n; i++)
                same = memcmp(&first_run[i], &cells.h_X[i], sizeof(Po_cell)) == 0 &&
                       first_lineage[i] == lineage.h_prop[i];
            CHECK(same, "division is reproducible bit for bit");
        }
    }
}

// ---- Protrusion_update: rewiring as a library operation ---------------------------
// rule: take a partner if there is none yet, or if it is closer than the current
__device__ bool closer_partner(
    const float3* __restrict__ d_X, int a, int b, Link current, float noise)
{
    if (current.a == current.b) return true;
    const float3 now = d_X[current.a] - d_X[current.b];
    const float3 then = d_X[a] - d_X[b];
    return norm3df(then.x, then.y, then.z) < norm3df(now.x, now.y, now.z);
}

static void check_protrusions()
{
    const int n = 30000, n_max = 40000, prots = 2;
    const float r_protrusion = 2.f;
    std::vector<Link> first_run;
    for (int run = 0; run < 2; run++) {
        Solution<float3, Grid_solver> cells{n_max, 64, 1.f};
        *cells.h_n = n;
        for (int i = 0; i < n_max; i++) cells.h_X[i] = float3{0};
        cells.copy_to_device();
        seeded_sphere(0.8f, cells, 5);
        Links protrusions{n_max * prots, 0.2f};
        Protrusion_update<float3> update{n_max, prots, 77, 40};
        for (int round = 0; round < 4; round++)
            update.rewire<closer_partner>(cells, protrusions, r_protrusion);
        cudaDeviceSynchronize();
        protrusions.copy_to_host();
        cells.copy_to_host();
        if (run == 0) {
            CHECK(*protrusions.h_n == n * prots, "the links' count follows the cells'");
            int live = 0;
            bool sane = true;
            for (int l = 0; l < n * prots; l++) {
                const Link link = protrusions.h_link[l];
                if (link.a == link.b) continue;
                live++;
                sane = sane && link.a == l / prots && link.b >= 0 && link.b < n;
                if (!sane) break;
                const float3 r = cells.h_X[link.a] - cells.h_X[link.b];
                sane = sane && sqrtf(r.x * r.x + r.y * r.y + r.z * r.z) <=
                                   r_protrusion * (1 + 1e-6f);
            }
            CHECK(sane, "links belong to their cell and span at most r_protrusion");
            CHECK(live > n * prots / 4, "most cells found partners in four rounds");
            first_run.assign(protrusions.h_link, protrusions.h_link + n * prots);
            // pulled through link_forces inside a step: action = reaction
            cells.take_step<relu_force<float3>>(0.05f,
                [&](const int, const float3* d_X, float3* d_dX) {
                    link_forces(protrusions, d_X, d_dX);
                });
            CHECK(cudaDeviceSynchronize() == cudaSuccess, "step with rewired links");
        } else {
            CHECK(memcmp(first_run.data(), protrusions.h_link,
                      sizeof(Link) * n * prots) == 0,
                "rewiring is reproducible bit for bit");
        }
    }
}

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp/yb_ext_test/";
    const int n = 3000, n_max = 4000;

    // ---- seeded generators through the header API ------------------------
    Solution<Po_cell, Grid_solver> cells{n_max, 50, 1.f};
    *cells.h_n = n;
    for (int i = 0; i < n_max; i++) cells.h_X[i] = Po_cell{0};
    cells.copy_to_device();
    seeded_sphere(0.8f, cells, 42);
    CHECK(*cells.h_n == n && cells.get_d_n() == n, "seeded_sphere keeps n");
    float r_max = 0;
    for (int i = 0; i < n; i++) {
        const Po_cell& X = cells.h_X[i];
        r_max = fmaxf(r_max, sqrtf(X.x * X.x + X.y * X.y + X.z * X.z));
        cells.h_X[i].theta = acosf(X.z / fmaxf(1e-6f, sqrtf(X.x * X.x + X.y * X.y + X.z * X.z)));
        cells.h_X[i].phi = atan2f(X.y, X.x);
    }
    const float radius = powf(n / 0.64f, 1.f / 3) * 0.8f / 2;
    CHECK(r_max <= radius * 1.00001f && r_max > 0.9f * radius,
        "seeded_sphere radius");
    cells.copy_to_device();

    // ---- frames: synchronous ASCII writer vs asynchronous ASCII writer -----
    {
        // same base name (it is part of the file header), different folders
        Vtk_output sync_out{"frame", dir + "sync/", false};
        Vtk_async_output<Po_cell> async_out{n_max, "frame", dir + "async/", false};
        async_out.add_field("theta", &Po_cell::theta);
        async_out.add_polarity(&Po_cell::theta, &Po_cell::phi);
        Vtk_async_output<Po_cell> binary_out{n_max, "binary", dir, true, 3};
        binary_out.add_field("theta", &Po_cell::theta);
        for (int frame = 0; frame < 4; frame++) {
            async_out.write(cells);   // returns at once
            binary_out.write(cells);
            cells.copy_to_host();     // the reference's blocking path
            sync_out.write_positions(cells);
            sync_out.write_field(cells, "theta", &Po_cell::theta);
            sync_out.write_polarity(cells);
            for (int k = 0; k < 5; k++) cells.take_step<layer>(0.05f);
        }
        async_out.wait();
        binary_out.wait();
        CHECK(async_out.frames_written() == 4, "async frames written");
        bool same = true;
        for (int frame = 0; frame < 4; frame++) {
            const std::string a =
                slurp(dir + "sync/frame_" + std::to_string(frame) + ".vtk");
            const std::string b = slurp(async_out.frame_path(frame));
            same = same && !a.empty() && a == b;
        }
        CHECK(same, "async ASCII frames identical to Vtk_output");
        const std::string first = slurp(binary_out.frame_path(3));
        CHECK(first.find("BINARY") != std::string::npos &&
                  first.size() > size_t(n) * 12,
            "binary frame written");
        // the binary frame reads back bit for bit; the ASCII one to 6 digits
        Vtk_input binary_in{binary_out.frame_path(3)};
        Vtk_input ascii_in{async_out.frame_path(3)};
        Solution<Po_cell, Grid_solver> from_binary{n_max, 50, 1.f};
        Solution<Po_cell, Grid_solver> from_ascii{n_max, 50, 1.f};
        binary_in.read_positions(from_binary);
        binary_in.read_field(from_binary, "theta", &Po_cell::theta);
        ascii_in.read_positions(from_ascii);
        ascii_in.read_polarity(from_ascii);
        CHECK(binary_in.n_points == n && ascii_in.n_points == n, "frames hold n points");
        bool close = true;
        for (int i = 0; i < n; i++) {
            // ASCII carries 6 significant digits; acos is ill-conditioned
            // near the poles, so compare theta away from them
            const float theta = from_binary.h_X[i].theta;
            close = close &&
                    fabsf(from_binary.h_X[i].x - from_ascii.h_X[i].x) <=
                        1e-5f * fmaxf(1.f, fabsf(from_binary.h_X[i].x));
            if (theta > 0.3f && theta < 2.8f)
                close = close && fabsf(theta - from_ascii.h_X[i].theta) < 1e-4f;
        }
        CHECK(close, "binary and ASCII frames read back the same state");
    }

    // ---- links and properties in the frame ---------------------------------
    {
        Links protrusions{n_max, 0.2f};
        Property<int> kind{n_max, "kind"};
        Property<float> weight{n_max, "weight"};
        for (int i = 0; i < n_max; i++) {
            protrusions.h_link[i] = Link{i, (i * 7 + 1) % n};
            kind.h_prop[i] = i % 3;
            weight.h_prop[i] = 0.25f * (i % 9);
        }
        *protrusions.h_n = n / 2;
        protrusions.copy_to_device();
        kind.copy_to_device();
        weight.copy_to_device();
        Vtk_output sync_out{"linked", dir + "sync/", false};
        Vtk_async_output<Po_cell> async_out{n_max, "linked", dir + "async/", false};
        async_out.add_links(protrusions);
        async_out.add_polarity(&Po_cell::theta, &Po_cell::phi);
        async_out.add_property(kind);
        async_out.add_property(weight);
        Vtk_async_output<Po_cell> binary_out{n_max, "linked_binary", dir, true};
        binary_out.add_links(protrusions);
        binary_out.add_property(kind);
        async_out.write(cells);
        binary_out.write(cells);
        cells.copy_to_host();
        sync_out.write_positions(cells);
        sync_out.write_links(protrusions);
        sync_out.write_polarity(cells);
        sync_out.write_property(kind);
        sync_out.write_property(weight);
        async_out.wait();
        binary_out.wait();
        const std::string a = slurp(dir + "sync/linked_0.vtk");
        CHECK(!a.empty() && a == slurp(async_out.frame_path(0)),
            "async frame with links and properties identical to Vtk_output");
        Vtk_input binary_in{binary_out.frame_path(0)};
        Property<int> kind_back{n_max, "kind"};
        binary_in.read_property(kind_back, "kind");
        bool same = true;
        for (int i = 0; i < n; i++) same = same && kind_back.h_prop[i] == kind.h_prop[i];
        CHECK(same, "binary frame with links: the property reads back");
    }

    // ---- a growing tissue: the frame holds the device-side count -----------
    {
        Vtk_async_output<Po_cell> out{n_max, "grown", dir, false};
        const int more = n + 100;
        cudaMemcpy(cells.d_n, &more, sizeof(int), cudaMemcpyHostToDevice);
        out.write(cells);
        out.wait();
        const std::string text = slurp(out.frame_path(0));
        CHECK(text.find("POINTS " + std::to_string(more) + " float") !=
                  std::string::npos,
            "frame uses the device-side cell count");
    }
    // ---- seeded box, plain and relaxed --------------------------------------
    {
        Solution<float3, Grid_solver> box{20000, 50, 1.f};
        const float3 lo{-3.f, -2.f, -1.f}, hi{5.f, 4.f, 3.f};
        seeded_cuboid(0.8f, lo, hi, box, 77);
        const double expected = 8. * 6. * 4. / (4. / 3 * M_PI * 0.4 * 0.4 * 0.4) * 0.64;
        CHECK(*box.h_n == int(expected) && box.get_d_n() == *box.h_n,
            "seeded_cuboid places the packing-fraction count");
        bool inside = true;
        double mean_x = 0;
        for (int i = 0; i < *box.h_n; i++) {
            const float3 X = box.h_X[i];
            inside = inside && X.x > lo.x && X.x <= hi.x && X.y > lo.y && X.y <= hi.y &&
                     X.z > lo.z && X.z <= hi.z;
            mean_x += X.x / *box.h_n;
        }
        CHECK(inside && fabs(mean_x - 1.0) < 0.2, "seeded_cuboid fills the box uniformly");

        Solution<float3, Grid_solver> relaxed{20000, 50, 1.f};
        relaxed_seeded_cuboid(0.75f, lo, hi, relaxed, 78, 0, 300);
        float nearest_min = 1e9f;
        const int m = *relaxed.h_n;
        for (int i = 0; i < 200; i++) {  // a sample of cells
            float nearest = 1e9f;
            for (int j = 0; j < m; j++) {
                if (j == i) continue;
                const float dx = relaxed.h_X[i].x - relaxed.h_X[j].x,
                            dy = relaxed.h_X[i].y - relaxed.h_X[j].y,
                            dz = relaxed.h_X[i].z - relaxed.h_X[j].z;
                nearest = fminf(nearest, sqrtf(dx * dx + dy * dy + dz * dz));
            }
            nearest_min = fminf(nearest_min, nearest);
        }
        CHECK(m > 400 && nearest_min > 0.4f * 0.75f,
            "relaxed_seeded_cuboid pushes overlapping cells apart");
    }

    check_division();
    check_protrusions();
    printf("all extension checks passed\n");
    return 0;
}
