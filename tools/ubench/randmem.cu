// Microbenchmark: random 32 B-sector access (dependent chains and independent gathers, plain loads and
// 64-bit atomicMax) as a function of the footprint -> TLB reach / page-walk cost on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o randmem randmem.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }

// MODE 0: dependent chain (next index depends on the loaded value): latency
// MODE 1: independent gathers, U in flight per thread: throughput
// MODE 2: independent fire-and-forget atomicMax
template <int MODE>
__global__ void __launch_bounds__(256) k(u64* buf, u64 n_mask, int iters, u64* out)
{
    u64 x = mix(blockIdx.x * 256ull + threadIdx.x + 12345);
    u64 acc = 0;
    if (MODE == 0) {
        for (int i = 0; i < iters; ++i) { const u64 v = buf[(x & n_mask) * 4]; x = mix(x + v + i); }
        acc = x;
    } else if (MODE == 1) {
        for (int i = 0; i < iters; i += 4) {
            u64 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) a[u] = buf[(mix(x + i + u) & n_mask) * 4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += a[u];
        }
    } else {
        for (int i = 0; i < iters; ++i) atomicMax(&buf[(mix(x + i) & n_mask) * 4], x + i);
    }
    if (acc == 0x1234567) out[0] = acc;
}

int main(int argc, char** argv)
{
    const size_t max_gb = argc > 1 ? atoi(argv[1]) : 64;
    u64* buf; u64* out;
    if (cudaMalloc(&buf, max_gb << 30) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(buf, 0, max_gb << 30);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%10s %14s %16s %16s %16s\n", "footprint", "chain ns/load", "chain Mld/s(full)", "gather Gld/s", "atomicMax Gop/s");
    for (size_t mb = 64; mb <= (max_gb << 10); mb *= 4) {
        const u64 n_mask = ((mb << 20) / 32) - 1;
        float ms;
        // latency: 1 warp
        k<0><<<1, 32>>>(buf, n_mask, 2000, out);
        cudaEventRecord(e0); k<0><<<1, 32>>>(buf, n_mask, 2000, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double lat = ms * 1e6 / 2000;
        // loaded chain: full occupancy
        const int G = 148 * 8;
        cudaEventRecord(e0); k<0><<<G, 256>>>(buf, n_mask, 200, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double chain = (double)G * 256 * 200 / ms / 1e3;
        cudaEventRecord(e0); k<1><<<G, 256>>>(buf, n_mask, 400, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double gat = (double)G * 256 * 400 / ms / 1e6;
        cudaEventRecord(e0); k<2><<<G, 256>>>(buf, n_mask, 200, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double atm = (double)G * 256 * 200 / ms / 1e6;
        printf("%8zu MB %14.0f %16.0f %16.2f %16.2f\n", mb, lat, chain, gat, atm);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
