// Initial conditions (reference: include/inits.cuh).
//
// Host-side generators that fill points.h_X[n_0 .. *h_n) and copy to the
// device, plus relu_force, the interaction used to relax random packings (and
// the benchmark's spring). The random generators draw from rand() after
// seeding it from std::random_device like the reference does; set the
// environment variable YALLA_B200_SEED for repeatable runs.
#pragma once

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <iostream>
#include <random>


template<typename Pt, template<typename> class Solver>
class Solution;


namespace yb_inits {

inline void seed_rand()
{
    const char* fixed = getenv("YALLA_B200_SEED");
    if (fixed && fixed[0]) {
        srand(static_cast<unsigned>(atoi(fixed)));
        return;
    }
    std::random_device entropy;
    srand(entropy());
}

// uniform in [0, 1)
inline double uniform() { return rand() / (RAND_MAX + 1.); }

// Number of relu_force steps used to relax n points.
inline int relaxation_steps(int n, const int (&limits)[3], const int (&steps)[4])
{
    for (int k = 0; k < 3; k++)
        if (limits[k] > 0 && n <= limits[k]) return steps[k];
    return steps[3];
}

inline void warn_if_large(int n, int limit)
{
    if (n > limit)
        std::cout << "Warning: The system is quite large, it may "
                  << "not be completely relaxed." << std::endl;
}

template<typename Pt, template<typename> class Solver>
void rescale_positions(Solution<Pt, Solver>& points, double scale)
{
    for (int i = 0; i < *points.h_n; i++) {
        points.h_X[i].x *= scale;
        points.h_X[i].y *= scale;
        points.h_X[i].z *= scale;
    }
}

}  // namespace yb_inits


// Uniformly filled disk in the y-z plane; radius from hexagonal packing.
template<typename Pt, template<typename> class Solver>
void random_disk(
    float dist_to_nb, Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    assert(n_0 < *points.h_n);
    yb_inits::seed_rand();
    const auto r_max =
        pow((*points.h_n - n_0) / 0.9069, 1. / 2) * dist_to_nb / 2;
    for (auto i = n_0; i < *points.h_n; i++) {
        const auto r = r_max * pow(yb_inits::uniform(), 1. / 2);
        const auto phi = yb_inits::uniform() * 2 * M_PI;
        points.h_X[i].x = 0;
        points.h_X[i].y = r * sin(phi);
        points.h_X[i].z = r * cos(phi);
    }
    points.copy_to_device();
}

// Uniformly filled ball; radius from random sphere packing (fraction 0.64).
template<typename Pt, template<typename> class Solver>
void random_sphere(
    float dist_to_nb, Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    assert(n_0 < *points.h_n);
    yb_inits::seed_rand();
    const auto r_max =
        pow((*points.h_n - n_0) / 0.64, 1. / 3) * dist_to_nb / 2;
    for (auto i = n_0; i < *points.h_n; i++) {
        const auto r = r_max * pow(yb_inits::uniform(), 1. / 3);
        const auto theta = acos(2. * yb_inits::uniform() - 1);
        const auto phi = yb_inits::uniform() * 2 * M_PI;
        points.h_X[i].x = r * sin(theta) * cos(phi);
        points.h_X[i].y = r * sin(theta) * sin(phi);
        points.h_X[i].z = r * cos(theta);
    }
    points.copy_to_device();
}

// Uniformly filled box; sets *h_n to the number of cells that fit.
template<typename Pt, template<typename> class Solver>
void random_cuboid(float dist_to_nb, float3 minimum, float3 maximum,
    Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    assert(n_0 < *points.h_n);

    const auto dimension = maximum - minimum;
    const auto cube_volume = dimension.x * dimension.y * dimension.z;
    const auto sphere_volume = 4. / 3 * M_PI * pow(dist_to_nb / 2, 3);
    const auto n = cube_volume / sphere_volume * 0.64;  // sphere packing

    assert(n_0 + n < *points.h_n);
    *points.h_n = n_0 + n;

    yb_inits::seed_rand();
    for (auto i = n_0; i < *points.h_n; i++) {
        points.h_X[i].x = minimum.x + dimension.x * yb_inits::uniform();
        points.h_X[i].y = minimum.y + dimension.y * yb_inits::uniform();
        points.h_X[i].z = minimum.z + dimension.z * yb_inits::uniform();
    }
    points.copy_to_device();
}


// Repulsion below 0.8, attraction between 0.8 and 1, nothing beyond:
// F(d) = 2 max(0.8 - d, 0) - max(d - 0.8, 0) along r for d <= 1.
template<typename Pt>
__device__ Pt relu_force(Pt Xi, Pt r, float dist, int i, int j)
{
    Pt dF{0};
    if (i == j) return dF;

    if (dist > 1.f) return dF;

    const auto F = fmaxf(0.8f - dist, 0) * 2.f - fmaxf(dist - 0.8f, 0);
    dF.x = r.x * F / dist;
    dF.y = r.y * F / dist;
    dF.z = r.z * F / dist;
    return dF;
}

// Ball relaxed with relu_force, then scaled to the wanted neighbour distance.
template<typename Pt, template<typename> class Solver>
void relaxed_sphere(
    float dist_to_nb, Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    random_sphere(0.6, points, n_0);

    const int n = *points.h_n;
    const int relax_steps =
        yb_inits::relaxation_steps(n, {100, 1000, 6000}, {500, 1000, 2000, 3000});
    yb_inits::warn_if_large(n, 10000);

    for (int i = 0; i < relax_steps; i++)
        points.template take_step<relu_force>(0.1f);
    points.copy_to_host();

    yb_inits::rescale_positions(points, dist_to_nb / 0.8);
    points.copy_to_device();
}

// Box relaxed with relu_force, then scaled to the wanted neighbour distance.
template<typename Pt, template<typename> class Solver>
void relaxed_cuboid(float dist_to_nb, float3 minimum, float3 maximum,
    Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    const auto scale = dist_to_nb / 0.8;
    random_cuboid(0.8, minimum / scale, maximum / scale, points, n_0);

    const int n = *points.h_n;
    const int relax_steps =
        yb_inits::relaxation_steps(n, {3000, 12000, 0}, {1000, 2000, 3000, 3000});
    yb_inits::warn_if_large(n, 15000);

    for (int i = 0; i < relax_steps; i++)
        points.template take_step<relu_force>(0.1f);
    points.copy_to_host();

    yb_inits::rescale_positions(points, scale);
    points.copy_to_device();
}


// Flat hexagonal lattice in the x-y plane, filled ring by ring from the
// centre: ring i has 6 corner cells at distance i * dist_to_nb and i - 1 cells
// spread evenly along each edge. Stops as soon as *h_n cells are placed.
template<typename Pt, template<typename> class Solver>
void regular_hexagon(
    float dist_to_nb, Solution<Pt, Solver>& points, unsigned int n_0 = 0)
{
    assert(n_0 < *points.h_n);

    unsigned int placed = n_0;
    const auto place = [&](float x, float y) {
        points.h_X[placed].x = x;
        points.h_X[placed].y = y;
        points.h_X[placed].z = 0.f;
        placed++;
        return placed == static_cast<unsigned int>(*points.h_n);
    };
    const auto beta = M_PI / 3.f;

    bool full = place(0.f, 0.f);
    for (int ring = 1; !full; ring++) {
        for (int corner = 0; corner < 6 && !full; corner++) {
            const auto angle = beta * corner;
            const float3 p{-dist_to_nb * ring * sinf(angle),
                dist_to_nb * ring * cosf(angle), 0.f};
            full = place(p.x, p.y);
            if (full || ring < 2) continue;

            const auto next_angle = beta * (corner + 1);
            const float3 q{-dist_to_nb * ring * sinf(next_angle),
                dist_to_nb * ring * cosf(next_angle), 0.f};
            auto edge = q - p;
            const auto length = sqrt(pow(edge.x, 2) + pow(edge.y, 2));
            edge = edge * (1.f / length);
            for (int k = 1; k <= ring - 1 && !full; k++) {
                const auto along = edge * length * (float(k) / float(ring));
                full = place(p.x + along.x, p.y + along.y);
            }
        }
    }
    points.copy_to_device();
}

// Flat triangular lattice in the x-y plane, rows of nx cells, odd rows shifted
// by half a spacing.
template<typename Pt, template<typename> class Solver>
void regular_rectangle(float dist_to_nb, int nx, Solution<Pt, Solver>& points,
    unsigned int n_0 = 0)
{
    assert(n_0 < *points.h_n);

    const float row_height =
        sqrt(pow(dist_to_nb, 2) - pow(dist_to_nb / 2.f, 2));
    unsigned int placed = n_0;
    for (int row = 0; placed < static_cast<unsigned int>(*points.h_n); row++) {
        const float py = row * row_height;
        const float shift = (row % 2 != 0) ? dist_to_nb / 2.f : 0.0f;
        for (int col = 0;
             col < nx && placed < static_cast<unsigned int>(*points.h_n);
             col++) {
            points.h_X[placed].x = shift + col * dist_to_nb;
            points.h_X[placed].y = py;
            points.h_X[placed].z = 0.0f;
            placed++;
        }
    }
    points.copy_to_device();
}


// Extension: the same distributions from a counter-based generator, generated
// and relaxed on the device (seeded_sphere, relaxed_seeded_sphere, ...).
#include "b200/seeded_inits.cuh"
