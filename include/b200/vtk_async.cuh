// Extension: asynchronous VTK output.
//
// The reference's output path blocks the simulation twice per frame: a
// cudaMemcpy of all n_max points (solvers.cuh:80-91) and a formatted ASCII
// write on the calling thread (vtk.cuh:94-214); examples/springs.cu:34 notes
// that this takes most of the run time. Vtk_async_output takes both off the
// critical path:
//
//   write(points)  enqueues, on the solver's stream, a device-side snapshot of
//                  the n live cells (n is read on the device) into one of a
//                  few staging slots and returns; the next take_step may start
//                  right away;
//   writer thread  waits for the snapshot, downloads exactly n cells into its
//                  own pinned buffer on its own stream, and writes the frame:
//                  the same legacy-VTK sections as Vtk_output -- byte for byte
//                  in ASCII mode, big-endian BINARY otherwise (8x smaller and
//                  no number formatting).
//
// write() blocks only when every slot is still waiting to be written.
#pragma once

#include <assert.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <sys/stat.h>

#include <condition_variable>
#include <deque>
#include <fstream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../cudebug.cuh"
#include "../links.cuh"
#include "../polarity.cuh"
#include "../property.cuh"
#include "grid_build.cuh"

template<typename Pt, template<typename> class Solver>
class Solution;

namespace yb {

// Copy the live cells, as 32-bit words, and their count.
__global__ void __launch_bounds__(256) snapshot_cells(const int* __restrict__ d_n,
    int n_max, int words_per_cell, const uint32_t* __restrict__ src,
    uint32_t* __restrict__ dst, int* __restrict__ count)
{
    const int n = live_cells(d_n, n_max);
    const long long words = static_cast<long long>(n) * words_per_cell;
    for (long long w = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
         w < words; w += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[w] = src[w];
    if (blockIdx.x == 0 && threadIdx.x == 0) *count = n;
}

inline uint32_t big_endian(uint32_t v) { return __builtin_bswap32(v); }

}  // namespace yb


template<typename Pt>
class Vtk_async_output {
public:
    Vtk_async_output(int n_max, std::string base_name,
        std::string output_path = "output/", bool binary = true, int n_slots = 2)
        : n_max{n_max}, base_name{base_name}, output_dir{output_path},
          binary{binary}, slots(n_slots)
    {
        if (output_dir.empty() || output_dir.back() != '/') output_dir.append("/");
        mkdir(output_dir.c_str(), 0755);
        const size_t bytes = static_cast<size_t>(n_max) * sizeof(Pt);
        for (auto& slot : slots) {
            YB_CUDA(cudaMalloc(&slot.d_cells, bytes));
            YB_CUDA(cudaMalloc(&slot.d_count, sizeof(int)));
            YB_CUDA(cudaEventCreateWithFlags(&slot.ready, cudaEventDisableTiming));
            free_slots.push_back(&slot);
        }
        YB_CUDA(cudaMallocHost(&h_cells, bytes));
        YB_CUDA(cudaMallocHost(&h_count, sizeof(int)));
        YB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        YB_CUDA(cudaGetDevice(&device));
        writer = std::thread([this] { run(); });
    }
    Vtk_async_output(const Vtk_async_output&) = delete;
    Vtk_async_output& operator=(const Vtk_async_output&) = delete;
    ~Vtk_async_output()
    {
        {
            std::lock_guard<std::mutex> lock(mutex);
            closing = true;
        }
        wake_writer.notify_all();
        writer.join();
        for (auto& slot : slots) {
            cudaFree(slot.d_cells);
            cudaFree(slot.d_count);
            cudaFree(slot.d_links);
            cudaFree(slot.d_link_count);
            for (auto* values : slot.d_properties) cudaFree(values);
            cudaEventDestroy(slot.ready);
        }
        cudaFreeHost(h_cells);
        cudaFreeHost(h_count);
        cudaFreeHost(h_links);
        for (auto& property : properties) cudaFreeHost(property.h_values);
        cudaStreamDestroy(copy_stream);
    }

    // Sections written after the positions of every frame, in this order:
    // scalar fields, then polarities (as unit normals, like write_polarity).
    void add_field(const char* data_name, float Pt::*field)
    {
        fields.push_back({data_name, lane_of(field)});
    }
    void add_polarity(float Pt::*theta, float Pt::*phi,
        const char* data_name = "polarity")
    {
        polarities.push_back({data_name, lane_of(theta), lane_of(phi)});
    }

    // Links as a LINES section behind the vertices, like Vtk_output::write_links;
    // per-cell properties (int, enum or float) as SCALARS behind the fields and
    // polarities, like Vtk_output::write_property. Both are snapshotted on the
    // device together with the cells (the links with their own device-side
    // count). Register before the first write().
    void add_links(Links& links_to_write)
    {
        assert(next_frame == 0 && links == nullptr);
        links = &links_to_write;
        const size_t n_links = static_cast<size_t>(links->n_max > 0 ? links->n_max : 1);
        for (auto& slot : slots) {
            YB_CUDA(cudaMalloc(&slot.d_links, n_links * sizeof(Link)));
            YB_CUDA(cudaMalloc(&slot.d_link_count, sizeof(int)));
        }
        YB_CUDA(cudaMallocHost(&h_links, n_links * sizeof(Link) + sizeof(int)));
    }
    template<typename Prop>
    void add_property(Property<Prop>& property)
    {
        static_assert(sizeof(Prop) == 4, "int, enum or float properties");
        assert(next_frame == 0 && property.n_max >= n_max);
        Property_field field;
        field.name = property.name;
        field.is_float = std::is_same<Prop, float>::value;
        field.d_source = reinterpret_cast<const uint32_t*>(property.d_prop);
        YB_CUDA(cudaMallocHost(&field.h_values, static_cast<size_t>(n_max) * 4));
        properties.push_back(field);
        for (auto& slot : slots) {
            uint32_t* values;
            YB_CUDA(cudaMalloc(&values, static_cast<size_t>(n_max) * 4));
            slot.d_properties.push_back(values);
        }
    }

    // Snapshot the state as it is in stream order and queue the frame.
    template<template<typename> class Solver>
    void write(Solution<Pt, Solver>& points)
    {
        assert(points.n_max <= n_max);
        Slot* slot;
        {
            std::unique_lock<std::mutex> lock(mutex);
            slot_freed.wait(lock, [this] { return !free_slots.empty(); });
            slot = free_slots.front();
            free_slots.pop_front();
        }
        const int words = sizeof(Pt) / sizeof(uint32_t);
        yb::snapshot_cells<<<yb::stride_grid(points.n_max * words, 256,
                                 yb::sm_count()),
            256, 0, points.stream>>>(points.d_n, points.n_max, words,
            reinterpret_cast<const uint32_t*>(points.d_X),
            reinterpret_cast<uint32_t*>(slot->d_cells), slot->d_count);
        if (links != nullptr)
            yb::snapshot_cells<<<yb::stride_grid(links->n_max * 2, 256,
                                     yb::sm_count()),
                256, 0, points.stream>>>(links->d_n, links->n_max, 2,
                reinterpret_cast<const uint32_t*>(links->d_link),
                reinterpret_cast<uint32_t*>(slot->d_links), slot->d_link_count);
        for (size_t k = 0; k < properties.size(); k++)
            yb::snapshot_cells<<<yb::stride_grid(points.n_max, 256, yb::sm_count()),
                256, 0, points.stream>>>(points.d_n, points.n_max, 1,
                properties[k].d_source, slot->d_properties[k], slot->d_count);
        YB_CUDA(cudaEventRecord(slot->ready, points.stream));
        slot->frame = next_frame++;
        {
            std::lock_guard<std::mutex> lock(mutex);
            queued.push_back(slot);
        }
        wake_writer.notify_one();
    }

    // Block until every queued frame is on disk.
    void wait()
    {
        std::unique_lock<std::mutex> lock(mutex);
        slot_freed.wait(lock, [this] {
            return free_slots.size() == slots.size() && in_flight == 0;
        });
    }
    int frames_written()
    {
        std::lock_guard<std::mutex> lock(mutex);
        return n_written;
    }
    std::string frame_path(int frame) const
    {
        return output_dir + base_name + "_" + std::to_string(frame) + ".vtk";
    }

private:
    struct Slot {
        Pt* d_cells = nullptr;
        int* d_count = nullptr;
        Link* d_links = nullptr;
        int* d_link_count = nullptr;
        std::vector<uint32_t*> d_properties;
        cudaEvent_t ready = nullptr;
        int frame = 0;
    };
    struct Property_field {
        std::string name;
        bool is_float = false;
        const uint32_t* d_source = nullptr;
        uint32_t* h_values = nullptr;  // pinned
    };
    struct Field {
        std::string name;
        int lane;
    };
    struct Polarity_field {
        std::string name;
        int theta, phi;
    };

    static int lane_of(float Pt::*member)
    {
        Pt probe{};
        return static_cast<int>(&(probe.*member) - reinterpret_cast<float*>(&probe));
    }

    void run()
    {
        cudaSetDevice(device);
        while (true) {
            Slot* slot;
            {
                std::unique_lock<std::mutex> lock(mutex);
                wake_writer.wait(
                    lock, [this] { return closing || !queued.empty(); });
                if (queued.empty()) return;  // closing and drained
                slot = queued.front();
                queued.pop_front();
            }
            YB_CUDA(cudaStreamWaitEvent(copy_stream, slot->ready, 0));
            YB_CUDA(cudaMemcpyAsync(h_count, slot->d_count, sizeof(int),
                cudaMemcpyDeviceToHost, copy_stream));
            YB_CUDA(cudaStreamSynchronize(copy_stream));
            const int n = *h_count;
            YB_CUDA(cudaMemcpyAsync(h_cells, slot->d_cells,
                static_cast<size_t>(n) * sizeof(Pt), cudaMemcpyDeviceToHost,
                copy_stream));
            int n_links = 0;
            if (links != nullptr) {
                int* h_link_count = reinterpret_cast<int*>(h_links + links->n_max);
                YB_CUDA(cudaMemcpyAsync(h_link_count, slot->d_link_count,
                    sizeof(int), cudaMemcpyDeviceToHost, copy_stream));
                YB_CUDA(cudaStreamSynchronize(copy_stream));
                n_links = *h_link_count;
                YB_CUDA(cudaMemcpyAsync(h_links, slot->d_links,
                    static_cast<size_t>(n_links) * sizeof(Link),
                    cudaMemcpyDeviceToHost, copy_stream));
            }
            for (size_t k = 0; k < properties.size(); k++)
                YB_CUDA(cudaMemcpyAsync(properties[k].h_values,
                    slot->d_properties[k], static_cast<size_t>(n) * 4,
                    cudaMemcpyDeviceToHost, copy_stream));
            YB_CUDA(cudaStreamSynchronize(copy_stream));
            const int frame = slot->frame;
            {
                // the device slot is free as soon as it is downloaded
                std::lock_guard<std::mutex> lock(mutex);
                in_flight++;
                free_slots.push_back(slot);
            }
            // (wait() must not see all slots free before the file is written)
            write_frame(frame, n, n_links);
            {
                std::lock_guard<std::mutex> lock(mutex);
                in_flight--;
                n_written++;
            }
            slot_freed.notify_all();
        }
    }

    const float* cell(int i) const
    {
        return reinterpret_cast<const float*>(h_cells + i);
    }
    float3 normal_of(int i, const Polarity_field& p) const
    {
        const Polarity angles{cell(i)[p.theta], cell(i)[p.phi]};
        float3 n = pol_to_float3(angles);  // as Vtk_output::write_polarity
        if (angles.theta == 0 && angles.phi == 0) n.z = 0;  // "no polarity"
        return n;
    }
    static void put(std::vector<uint32_t>& out, float v)
    {
        uint32_t bits;
        memcpy(&bits, &v, sizeof(bits));
        out.push_back(yb::big_endian(bits));
    }
    static void flush(std::ofstream& file, std::vector<uint32_t>& out)
    {
        file.write(reinterpret_cast<const char*>(out.data()),
            static_cast<std::streamsize>(out.size() * sizeof(uint32_t)));
        out.clear();
    }

    void write_frame(int frame, int n, int n_links)
    {
        std::ofstream file(frame_path(frame), std::ios::binary);
        assert(file.is_open());
        file << "# vtk DataFile Version 3.0\n"
             << base_name << "\n"
             << (binary ? "BINARY\n" : "ASCII\n") << "DATASET POLYDATA\n"
             << "\nPOINTS " << n << " float\n";
        std::vector<uint32_t> out;
        if (binary) {
            out.reserve(static_cast<size_t>(n) * 3);
            for (int i = 0; i < n; i++) {
                put(out, cell(i)[0]);
                put(out, cell(i)[1]);
                put(out, cell(i)[2]);
            }
            flush(file, out);
            file << "\n";
        } else {
            for (int i = 0; i < n; i++)
                file << cell(i)[0] << " " << cell(i)[1] << " " << cell(i)[2]
                     << "\n";
        }

        file << "\nVERTICES " << n << " " << 2 * n << "\n";
        if (binary) {
            for (int i = 0; i < n; i++) {
                out.push_back(yb::big_endian(1u));
                out.push_back(yb::big_endian(static_cast<uint32_t>(i)));
            }
            flush(file, out);
            file << "\n";
        } else {
            for (int i = 0; i < n; i++) file << "1 " << i << "\n";
        }

        if (links != nullptr) {
            file << "\nLINES " << n_links << " " << 3 * n_links << "\n";
            if (binary) {
                for (int i = 0; i < n_links; i++) {
                    out.push_back(yb::big_endian(2u));
                    out.push_back(yb::big_endian(static_cast<uint32_t>(h_links[i].a)));
                    out.push_back(yb::big_endian(static_cast<uint32_t>(h_links[i].b)));
                }
                flush(file, out);
                file << "\n";
            } else {
                for (int i = 0; i < n_links; i++)
                    file << "2 " << h_links[i].a << " " << h_links[i].b << "\n";
            }
        }

        if (!fields.empty() || !polarities.empty() || !properties.empty())
            file << "\nPOINT_DATA " << n << "\n";
        for (const auto& field : fields) {
            file << "SCALARS " << field.name << " float\n"
                 << "LOOKUP_TABLE default\n";
            if (binary) {
                for (int i = 0; i < n; i++) put(out, cell(i)[field.lane]);
                flush(file, out);
                file << "\n";
            } else {
                for (int i = 0; i < n; i++) file << cell(i)[field.lane] << "\n";
            }
        }
        for (const auto& polarity : polarities) {
            file << "NORMALS " << polarity.name << " float\n";
            for (int i = 0; i < n; i++) {
                const float3 v = normal_of(i, polarity);
                if (binary) {
                    put(out, v.x), put(out, v.y), put(out, v.z);
                } else {
                    file << v.x << " " << v.y << " " << v.z << "\n";
                }
            }
            if (binary) {
                flush(file, out);
                file << "\n";
            }
        }
        for (const auto& property : properties) {
            file << "SCALARS " << property.name << " "
                 << (property.is_float ? "float" : "int") << "\n"
                 << "LOOKUP_TABLE default\n";
            if (binary) {
                for (int i = 0; i < n; i++)
                    out.push_back(yb::big_endian(property.h_values[i]));
                flush(file, out);
                file << "\n";
            } else if (property.is_float) {
                for (int i = 0; i < n; i++) {
                    float v;
                    memcpy(&v, &property.h_values[i], sizeof(v));
                    file << v << "\n";
                }
            } else {
                for (int i = 0; i < n; i++)
                    file << static_cast<int>(property.h_values[i]) << "\n";
            }
        }
    }

    const int n_max;
    std::string base_name, output_dir;
    const bool binary;
    std::vector<Slot> slots;
    std::vector<Field> fields;
    std::vector<Polarity_field> polarities;
    std::vector<Property_field> properties;
    Links* links = nullptr;
    Link* h_links = nullptr;  // pinned, followed by the count
    Pt* h_cells = nullptr;
    int* h_count = nullptr;
    cudaStream_t copy_stream = nullptr;
    int device = 0;
    int next_frame = 0;

    std::mutex mutex;
    std::condition_variable wake_writer, slot_freed;
    std::deque<Slot*> free_slots, queued;
    int in_flight = 0, n_written = 0;
    bool closing = false;
    std::thread writer;
};
