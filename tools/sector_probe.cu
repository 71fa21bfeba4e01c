// tools/sector_probe.cu -- micro-benchmark behind a number in DESIGN.md / profiles/README.md: what ONE random 32-byte
// read (the fp32 density octet: two LDG.128 to one aligned 32-byte sector) costs in DRAM traffic on a B200 when the
// buffer is far larger than L2.  Run under ncu and compare dram__bytes_read.sum with the bytes asked for:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_probe tools/sector_probe.cu
//   ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./sector_probe [GiB=16] [reads per thread=64] [bytes=32|16|64|128] [L2 fetch granularity hint]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
template <int BYTES>
__global__ void k_probe(const uint4* __restrict__ buf, size_t n_units, int reads, float* __restrict__ out)
{
    uint32_t h = mix(blockIdx.x * blockDim.x + threadIdx.x + 1u);
    float    s = 0.0f;
    for (int i = 0; i < reads; i++)
    {
        h = mix(h + 0x9e3779b9u);
        const size_t u = (((size_t)h << 16) ^ mix(h)) % n_units;   // unit = BYTES-aligned block
        const uint4* p = buf + u * (BYTES / 16);
#pragma unroll
        for (int q = 0; q < BYTES / 16; q++)
        {
            uint4 v = __ldg(p + q);
            s += __uint_as_float(v.x) + __uint_as_float(v.w);
        }
    }
    if (s == 123.456f) out[0] = s;
}
int main(int argc, char** argv)
{
    const size_t gib   = argc > 1 ? (size_t)atoll(argv[1]) : 16;
    const int    reads = argc > 2 ? atoi(argv[2]) : 64;
    const int    bytes = argc > 3 ? atoi(argv[3]) : 32;
    const int    gran  = argc > 4 ? atoi(argv[4]) : 0;  // cudaLimitMaxL2FetchGranularity hint (32 / 64 / 128), 0 = leave the default
    if (gran) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
    size_t gran_now = 0;
    cudaDeviceGetLimit(&gran_now, cudaLimitMaxL2FetchGranularity);
    const size_t total = gib << 30;
    uint4* buf = nullptr;
    float* out = nullptr;
    if (cudaMalloc(&buf, total) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, total);
    const int    threads = 256, blocks = 148 * 64;
    const size_t n_units = total / bytes;
    cudaEvent_t  e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++)
    {
        cudaEventRecord(e0);
        if (bytes == 16) k_probe<16><<<blocks, threads>>>(buf, n_units, reads, out);
        else if (bytes == 32) k_probe<32><<<blocks, threads>>>(buf, n_units, reads, out);
        else if (bytes == 64) k_probe<64><<<blocks, threads>>>(buf, n_units, reads, out);
        else k_probe<128><<<blocks, threads>>>(buf, n_units, reads, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double asked = (double)blocks * threads * reads * bytes;
    printf("{\"l2_fetch_granularity\": %zu, \"buffer_GiB\": %zu, \"read_bytes\": %d, \"reads\": %.0f, \"bytes_asked\": %.0f, \"ms\": %.3f, \"asked_GBps\": %.1f}\n", gran_now, gib, bytes,
           (double)blocks * threads * reads, asked, ms, asked / ms * 1e-6);
    return cudaGetLastError() != cudaSuccess;
}
