// Closed triangle meshes for image-based models (reference: include/mesh.cuh):
// load a legacy-VTK surface, move/rotate/scale/inflate it, test whether a point
// lies inside (ray casting), write it back, and compare shapes -- the mean
// nearest-neighbour distance between two point sets, both ways. Host-side
// tooling plus one tiled device kernel; not part of the per-step hot path.
#pragma once

#include <assert.h>
#include <math.h>
#include <algorithm>
#include <array>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "dtypes.cuh"
#include "solvers.cuh"
#include "utils.cuh"


// ---- shape comparison --------------------------------------------------------
// d_min_dist[i] = distance from point i of set 1 to the nearest point of set
// 2. One thread per point of set 1, set 2 staged through shared memory
// TILE_SIZE points at a time (reference: mesh.cuh:27-56).
template<typename Pt1, typename Pt2>
__global__ void compute_minimum_distance(const int n1, const int n2,
    const Pt1* __restrict__ d_X1, const Pt2* __restrict__ d_X2,
    float* d_min_dist)
{
    __shared__ float s_x[TILE_SIZE], s_y[TILE_SIZE], s_z[TILE_SIZE];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float xi = 0.f, yi = 0.f, zi = 0.f;
    if (i < n1) xi = d_X1[i].x, yi = d_X1[i].y, zi = d_X1[i].z;

    float nearest = 0.f;
    for (int tile_start = 0; tile_start < n2; tile_start += TILE_SIZE) {
        const int mine = tile_start + threadIdx.x;
        __syncthreads();
        if (mine < n2) {
            s_x[threadIdx.x] = d_X2[mine].x;
            s_y[threadIdx.x] = d_X2[mine].y;
            s_z[threadIdx.x] = d_X2[mine].z;
        }
        __syncthreads();
        const int in_tile = min(TILE_SIZE, n2 - tile_start);
        for (int k = 0; k < in_tile; k++) {
            const float dist = norm3df(xi - s_x[k], yi - s_y[k], zi - s_z[k]);
            nearest = (tile_start + k == 0) ? dist : fminf(dist, nearest);
        }
    }
    if (i < n1) d_min_dist[i] = nearest;
}

namespace yb_mesh {
// Sum of n floats, one block, fixed order (replaces thrust::reduce).
__global__ void sum_floats(const float* __restrict__ values, int n, float* total)
{
    __shared__ float partial[256];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) s += values[i];
    partial[threadIdx.x] = s;
    __syncthreads();
    for (int width = 128; width > 0; width >>= 1) {
        if (threadIdx.x < width) partial[threadIdx.x] += partial[threadIdx.x + width];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = partial[0];
}

template<typename Pt1, typename Pt2>
float mean_nearest_distance(
    int n1, int n2, const Pt1* d_X1, const Pt2* d_X2)
{
    float *d_dist, *d_total, total = 0.f;
    YB_CUDA(cudaMalloc(&d_dist, n1 * sizeof(float)));
    YB_CUDA(cudaMalloc(&d_total, sizeof(float)));
    compute_minimum_distance<<<(n1 + TILE_SIZE - 1) / TILE_SIZE, TILE_SIZE>>>(
        n1, n2, d_X1, d_X2, d_dist);
    sum_floats<<<1, 256>>>(d_dist, n1, d_total);
    YB_CUDA(cudaMemcpy(&total, d_total, sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(d_total);
    cudaFree(d_dist);
    return total / n1;
}
}  // namespace yb_mesh

// Mean distance from each point of one set to the nearest point of the other,
// averaged over both directions.
template<typename Pt1, typename Pt2>
float shape_comparison(const int n1, const int n2, const Pt1* __restrict__ d_X1,
    const Pt2* __restrict__ d_X2)
{
    const float mean_12 = yb_mesh::mean_nearest_distance(n1, n2, d_X1, d_X2);
    const float mean_21 = yb_mesh::mean_nearest_distance(n2, n1, d_X2, d_X1);
    return (mean_12 + mean_21) / 2;
}

template<typename Pt1, typename Pt2, template<typename> class Solver1,
    template<typename> class Solver2>
float shape_comparison_points_to_points(
    Solution<Pt1, Solver1>& points1, Solution<Pt2, Solver2>& points2)
{
    return shape_comparison(
        points1.get_d_n(), points2.get_d_n(), points1.d_X, points2.d_X);
}


// ---- meshes ---------------------------------------------------------------------
struct Ray {
    float3 P0;
    float3 P1;
    Ray(float3 P0, float3 P1) : P0{P0}, P1{P1} {}
};

struct Triangle {
    float3 V0;
    float3 V1;
    float3 V2;
    float3 C;  // centroid
    float3 n;  // unit normal
    Triangle() : Triangle(float3{0}, float3{0}, float3{0}) {}
    Triangle(float3 V0, float3 V1, float3 V2) : V0{V0}, V1{V1}, V2{V2}
    {
        calculate_centroid();
        calculate_normal();
    }
    void calculate_centroid() { C = (V0 + V1 + V2) / 3.f; }
    void calculate_normal()
    {
        const auto v = V2 - V0;
        const auto u = V1 - V0;
        n = float3{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z,
            u.x * v.y - u.y * v.x};
        n /= sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    }
};

class Mesh {
public:
    std::vector<float3> vertices;
    std::vector<Triangle> facets;
    float3* d_vertices;
    std::vector<std::array<int, 3>> triangle_to_vertices;
    std::vector<std::vector<int>> vertex_to_triangles;
    Mesh();
    Mesh(std::string file_name);
    Mesh(const Mesh& copy);
    ~Mesh();
    Mesh& operator=(const Mesh& other);
    float3 get_minimum();
    float3 get_maximum();
    void translate(float3 offset);
    void rotate(float around_z, float around_y, float around_x);
    void rescale(float factor);
    void grow_normally(float amount, bool boundary);
    template<typename Pt>
    bool test_exclusion(const Pt point);
    void write_vtk(std::string);
    void copy_to_device();
    template<typename Pt, template<typename> class Solver>
    float shape_comparison_mesh_to_points(Solution<Pt, Solver>& points);

private:
    template<typename Function>
    void for_every_point(Function apply)
    {
        for (auto& vertex : vertices) apply(vertex);
        for (auto& facet : facets) {
            apply(facet.V0);
            apply(facet.V1);
            apply(facet.V2);
            apply(facet.C);
        }
    }
    void allocate_device_vertices()
    {
        const size_t count = vertices.size() > 0 ? vertices.size() : 1;
        YB_CUDA(cudaMalloc(&d_vertices, count * sizeof(float3)));
    }
};

inline Mesh::Mesh() : d_vertices{nullptr} {}

// Legacy VTK: "POINTS n type" followed by coordinates (any number of points
// per line), then "POLYGONS m ..." or "CELLS m ..." with "3 a b c" per line.
inline Mesh::Mesh(std::string file_name)
{
    std::ifstream input_file(file_name);
    assert(input_file.is_open());

    std::string line;
    std::vector<std::string> items;
    const auto skip_to = [&](const char* keyword, const char* alternative) {
        while (getline(input_file, line)) {
            items = split(line);
            if (items.empty()) continue;
            if (items[0] == keyword || items[0] == alternative) return;
        }
        assert(false && "keyword not found in mesh file");
    };

    skip_to("POINTS", "POINTS");
    const int n_vertices = stoi(items[1]);
    while (static_cast<int>(vertices.size()) < n_vertices) {
        getline(input_file, line);
        items = split(line);
        for (size_t k = 0; k + 2 < items.size(); k += 3)
            vertices.push_back(
                float3{stof(items[k]), stof(items[k + 1]), stof(items[k + 2])});
    }
    allocate_device_vertices();

    skip_to("POLYGONS", "CELLS");
    const int n_facets = stoi(items[1]);
    assert(n_facets % 2 == 0);  // otherwise the mesh cannot be closed

    vertex_to_triangles = std::vector<std::vector<int>>(n_vertices);
    for (int i = 0; i < n_facets; i++) {
        getline(input_file, line);
        items = split(line);
        const std::array<int, 3> corner{
            stoi(items[1]), stoi(items[2]), stoi(items[3])};
        triangle_to_vertices.push_back(corner);
        facets.push_back(Triangle(
            vertices[corner[0]], vertices[corner[1]], vertices[corner[2]]));
        for (int c = 0; c < 3; c++) vertex_to_triangles[corner[c]].push_back(i);
    }
}

inline Mesh::Mesh(const Mesh& copy)
    : vertices{copy.vertices}, facets{copy.facets},
      triangle_to_vertices{copy.triangle_to_vertices},
      vertex_to_triangles{copy.vertex_to_triangles}
{
    allocate_device_vertices();
}

inline Mesh::~Mesh() { cudaFree(d_vertices); }

inline Mesh& Mesh::operator=(const Mesh& other)
{
    if (this == &other) return *this;
    vertices = other.vertices;
    facets = other.facets;
    triangle_to_vertices = other.triangle_to_vertices;
    vertex_to_triangles = other.vertex_to_triangles;
    cudaFree(d_vertices);
    allocate_device_vertices();
    return *this;
}

inline float3 Mesh::get_minimum()
{
    float3 minimum = vertices[0];
    for (const auto& v : vertices) {
        minimum.x = std::min(minimum.x, v.x);
        minimum.y = std::min(minimum.y, v.y);
        minimum.z = std::min(minimum.z, v.z);
    }
    return minimum;
}

inline float3 Mesh::get_maximum()
{
    float3 maximum = vertices[0];
    for (const auto& v : vertices) {
        maximum.x = std::max(maximum.x, v.x);
        maximum.y = std::max(maximum.y, v.y);
        maximum.z = std::max(maximum.z, v.z);
    }
    return maximum;
}

inline void Mesh::translate(float3 offset)
{
    for_every_point([&](float3& p) { p = p + offset; });
}

// Rotations about z, then y, then x, each as a plane rotation of the other two
// coordinates (same conventions and order as the reference, mesh.cuh:257-333).
inline void Mesh::rotate(float around_z, float around_y, float around_x)
{
    const auto turn = [](float& a, float& b, float angle) {
        const float old_a = a, old_b = b;
        a = old_a * cos(angle) - old_b * sin(angle);
        b = old_a * sin(angle) + old_b * cos(angle);
    };
    for_every_point([&](float3& p) { turn(p.x, p.y, around_z); });
    for_every_point([&](float3& p) { turn(p.x, p.z, around_y); });
    for_every_point([&](float3& p) { turn(p.y, p.z, around_x); });
    for (auto& facet : facets) facet.calculate_normal();
}

inline void Mesh::rescale(float factor)
{
    for_every_point([&](float3& p) { p = p * factor; });
}

// Move every vertex by `amount` along the average normal of its facets; with
// boundary = true, vertices in the plane x = 0 stay put.
inline void Mesh::grow_normally(float amount, bool boundary = false)
{
    for (size_t i = 0; i < vertices.size(); i++) {
        if (boundary && vertices[i].x == 0.f) continue;

        float3 average_normal{0};
        for (const int triangle : vertex_to_triangles[i])
            average_normal = average_normal + facets[triangle].n;
        const float length =
            sqrt(pow(average_normal.x, 2) + pow(average_normal.y, 2) +
                 pow(average_normal.z, 2));
        vertices[i] = vertices[i] + average_normal * (amount / length);
    }
    for (size_t i = 0; i < facets.size(); i++) {
        facets[i].V0 = vertices[triangle_to_vertices[i][0]];
        facets[i].V1 = vertices[triangle_to_vertices[i][1]];
        facets[i].V2 = vertices[triangle_to_vertices[i][2]];
        facets[i].calculate_centroid();
        facets[i].calculate_normal();
    }
}

// Does the ray from R.P0 through R.P1 hit the triangle? (parametric plane
// intersection + barycentric test, http://geomalgorithms.com/a06-_intersect-2.html)
inline bool intersect(Ray R, Triangle T)
{
    const auto direction = R.P1 - R.P0;
    const auto r = dot_product(T.n, T.V0 - R.P0) / dot_product(T.n, direction);
    if (r < 0) return false;  // the plane is behind the ray's origin

    const auto hit = R.P0 + direction * r;
    const auto u = T.V1 - T.V0;
    const auto v = T.V2 - T.V0;
    const auto w = hit - T.V0;
    const auto uu = dot_product(u, u), uv = dot_product(u, v),
               vv = dot_product(v, v), wu = dot_product(w, u),
               wv = dot_product(w, v);
    const auto denom = uv * uv - uu * vv;

    const auto s = (uv * wv - vv * wu) / denom;
    if (s < 0.0 or s > 1.0) return false;
    const auto t = (uv * wu - uu * wv) / denom;
    if (t < 0.0 or (s + t) > 1.0) return false;
    return true;
}

// True if the point is OUTSIDE the closed mesh: a ray in a fixed generic
// direction crosses the surface an even number of times.
template<typename Pt>
bool Mesh::test_exclusion(const Pt point)
{
    const auto p_0 = float3{point.x, point.y, point.z};
    const auto p_1 = p_0 + float3{0.22788, 0.38849, 0.81499};
    const Ray R(p_0, p_1);
    int n_intersections = 0;
    for (const auto& facet : facets) n_intersections += intersect(R, facet);
    return (n_intersections % 2 == 0);
}

inline void Mesh::write_vtk(std::string output_tag)
{
    std::ofstream mesh_file("output/" + output_tag + ".mesh.vtk");
    assert(mesh_file.is_open());

    mesh_file << "# vtk DataFile Version 3.0\n"
              << output_tag + ".mesh"
              << "\n"
              << "ASCII\n"
              << "DATASET POLYDATA\n";
    mesh_file << "\nPOINTS " << 3 * facets.size() << " float\n";
    for (const auto& facet : facets)
        for (const float3& corner : {facet.V0, facet.V1, facet.V2})
            mesh_file << corner.x << " " << corner.y << " " << corner.z << "\n";
    mesh_file << "\nPOLYGONS " << facets.size() << " " << 4 * facets.size()
              << "\n";
    for (size_t i = 0; i < 3 * facets.size(); i += 3)
        mesh_file << "3 " << i << " " << i + 1 << " " << i + 2 << "\n";
}

inline void Mesh::copy_to_device()
{
    YB_CUDA(cudaMemcpy(d_vertices, vertices.data(),
        vertices.size() * sizeof(float3), cudaMemcpyHostToDevice));
}

template<typename Pt, template<typename> class Solver>
float Mesh::shape_comparison_mesh_to_points(Solution<Pt, Solver>& points)
{
    return shape_comparison(
        static_cast<int>(vertices.size()), points.get_d_n(), d_vertices, points.d_X);
}
