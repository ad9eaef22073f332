// Microbenchmark: does the packed FP32 pipe (FMUL2/FADD2, sm_100a) double non-FMA FP32 throughput?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b){ u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float mul1(float a, float b){ float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float add1(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float m)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0f + threadIdx.x * 1e-6f + i;
    if (MODE == 0) {          // scalar: 16 independent FMUL+FADD chains
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = mul1(a[i], m); a[i] = mul1(a[i], 1.0009765625f); }
        }
    } else {                  // packed: 8 FMUL2 + 8 FADD2 on the same 16 values
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p[i] = pk(a[2 * i], a[2 * i + 1]);
        const u64 mm = pk(m, m), cc = pk(1.0009765625f, 1.0009765625f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { p[i] = mul2(p[i], mm); p[i] = mul2(p[i], cc); }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0,%1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(out, iters, 0.999f); else k<1><<<148 * 8, 256>>>(out, iters, 0.999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = 148.0 * 8 * 256 * (double)iters * 32;   // scalar-equivalent muls per iteration
            printf("%s rep %d: %.3f ms  %.2f T scalar-op/s\n", mode ? "packed f32x2" : "scalar f32  ", rep, ms, ops / ms / 1e9);
        }
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
