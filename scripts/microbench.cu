// Chip peaks the sweep kernel is judged against (north_star: "microbenchmarked
// chip peaks"): FP32 FMA issue rate, shared-memory LDS.128 bandwidth, L2->SM
// bandwidth on a 64 MB working set, HBM read bandwidth on a 4 GB working set.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench scripts/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fma_peak(float* out, int iters)
{
    float a[8];
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 1e-3f + k;
    const float b = 1.000001f, c = 1e-7f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], b, c);
    }
    float s = 0;
    for (int k = 0; k < 8; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void lds_peak(float* out, int iters)
{
    __shared__ float4 tile[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x)
        tile[k] = make_float4(k, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int at = threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const float4 v = tile[(at + u * 32) & 1023];
            acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
        }
        at += 7;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void read_stream(const float4* __restrict__ in, size_t n, float* out)
{
    float4 acc = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n;
         i += size_t(gridDim.x) * blockDim.x) {
        const float4 v = in[i];
        acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.f) out[0] = 1;
}

template<typename F>
float time_ms(F launch, int reps)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < reps; r++) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 32 * 1024);

    const int iters = 4096;
    float ms = time_ms([&] { fma_peak<<<sms * 8, 256>>>(out, iters); }, 5);
    const double flops = 2.0 * sms * 8 * 256 * double(iters) * 16 * 8;
    const double fma_tflops = flops / (ms * 1e-3) / 1e12;

    ms = time_ms([&] { lds_peak<<<sms * 8, 256>>>(out, iters); }, 5);
    const double lds_bytes = 16.0 * sms * 8 * 256 * double(iters) * 8;
    const double lds_tbs = lds_bytes / (ms * 1e-3) / 1e12;

    float4* buf;
    const size_t small = size_t(64) << 20, big = size_t(4) << 30;
    cudaMalloc(&buf, big);
    cudaMemset(buf, 0, big);
    ms = time_ms([&] { read_stream<<<sms * 16, 256>>>(buf, small / 16, out); }, 20);
    const double l2_tbs = small / (ms * 1e-3) / 1e12;
    ms = time_ms([&] { read_stream<<<sms * 16, 256>>>(buf, big / 16, out); }, 3);
    const double hbm_tbs = big / (ms * 1e-3) / 1e12;

    printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp32_fma_tflops\": %.2f, "
           "\"fp32_lane_instr_per_s_T\": %.2f, \"smem_lds128_TBps\": %.2f, "
           "\"l2_read_64MB_TBps\": %.2f, \"hbm_read_4GB_TBps\": %.2f}\n",
        prop.name, sms, fma_tflops, fma_tflops / 2, lds_tbs, l2_tbs, hbm_tbs);
    return 0;
}
