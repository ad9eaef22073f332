// Microbenchmark: issue rates of non-FMA FP32 instructions on sm_100a by operand form, and of the scorer's inner loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rates fp32_rates.cu && ./fp32_rates
// Prints thread-level results per clock per SM (128 = one 32-lane instruction per scheduler per clock, scalar).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define F1(op, r, a, b) asm volatile(op ".rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b))
#define F2(op, r, a, b) asm volatile(op ".rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b))
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 p) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); return a; }
__device__ __forceinline__ float hi(u64 p) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); return b; }

template <int MODE>
__global__ void __launch_bounds__(256, 2) k(float* out, const float* in, int iters, long long* cyc)
{
    __shared__ float4 xs[64 * 10];
    for (int i = threadIdx.x; i < 640; i += 256) xs[i] = make_float4(in[i & 63], in[(i + 1) & 63], in[(i + 2) & 63], in[(i + 3) & 63]);
    float a[16], b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = 1.0f + threadIdx.x * 1e-6f + i; b[i] = in[(threadIdx.x + i) & 63]; }
    u64 pa[8], pb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pa[i] = pk(a[2 * i], a[2 * i + 1]); pb[i] = pk(b[2 * i], b[2 * i + 1]); }
    __syncthreads();
    const long long t0 = clock64();
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // 32 FMUL, one register source + immediate
#pragma unroll
            for (int i = 0; i < 16; ++i) { F1("mul", a[i], a[i], 1.0009765625f); F1("mul", a[i], a[i], 0.9990234375f); }
        } else if (MODE == 1) {     // 32 FMUL, two register sources
#pragma unroll
            for (int i = 0; i < 16; ++i) { F1("mul", a[i], a[i], b[i]); F1("mul", a[i], a[i], b[(i + 5) & 15]); }
        } else if (MODE == 2) {     // 32 FADD, two register sources
#pragma unroll
            for (int i = 0; i < 16; ++i) { F1("add", a[i], a[i], b[i]); F1("add", a[i], a[i], b[(i + 5) & 15]); }
        } else if (MODE == 3) {     // 16 FMUL2 (32 results), two register-pair sources
#pragma unroll
            for (int i = 0; i < 8; ++i) { F2("mul", pa[i], pa[i], pb[i]); F2("mul", pa[i], pa[i], pb[(i + 3) & 7]); }
        } else if (MODE == 4) {     // 16 FADD2
#pragma unroll
            for (int i = 0; i < 8; ++i) { F2("add", pa[i], pa[i], pb[i]); F2("add", pa[i], pa[i], pb[(i + 3) & 7]); }
        } else if (MODE == 5) {     // scorer element, scalar: 4 rows x 8 dims from registers: sub, mul, mul, add (128 results)
#pragma unroll
            for (int dd = 0; dd < 8; ++dd) {
                float d, t;
                F1("add", d, a[dd], b[dd]); F1("mul", t, d, d); F1("mul", t, t, b[8 + dd]); F1("add", s0, s0, t);
                F1("add", d, a[8 + dd], b[dd]); F1("mul", t, d, d); F1("mul", t, t, b[8 + dd]); F1("add", s1, s1, t);
                F1("add", d, a[(dd + 3) & 15], b[dd]); F1("mul", t, d, d); F1("mul", t, t, b[8 + dd]); F1("add", s2, s2, t);
                F1("add", d, a[(dd + 11) & 15], b[dd]); F1("mul", t, d, d); F1("mul", t, t, b[8 + dd]); F1("add", s3, s3, t);
            }
        } else if (MODE == 6) {     // scorer element, packed: same 128 results as 48 packed + 32 scalar adds
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                u64 d, t;
                F2("add", d, pa[q], pb[q]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q]); F1("add", s0, s0, lo(t)); F1("add", s0, s0, hi(t));
                F2("add", d, pa[4 + q], pb[q]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q]); F1("add", s1, s1, lo(t)); F1("add", s1, s1, hi(t));
                F2("add", d, pa[(q + 2) & 7], pb[q]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q]); F1("add", s2, s2, lo(t)); F1("add", s2, s2, hi(t));
                F2("add", d, pa[(q + 5) & 7], pb[q]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q]); F1("add", s3, s3, lo(t)); F1("add", s3, s3, hi(t));
            }
        } else if (MODE == 7 || MODE == 8) {   // as 5 / 6 with x from shared memory (broadcast LDS.128), 4 rows x 8 dims
            const float4* r0 = xs + (it & 15) * 40;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 xv = r0[r * 10 + q];
                    float& s = r == 0 ? s0 : r == 1 ? s1 : r == 2 ? s2 : s3;
                    if (MODE == 7) {
                        const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) { float d, t; F1("add", d, xa[e], b[q * 4 + e]); F1("mul", t, d, d); F1("mul", t, t, b[8 + q * 4 + e]); F1("add", s, s, t); }
                    } else {
                        u64 d, t;
                        F2("add", d, pk(xv.x, xv.y), pb[q * 2]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q * 2]); F1("add", s, s, lo(t)); F1("add", s, s, hi(t));
                        F2("add", d, pk(xv.z, xv.w), pb[q * 2 + 1]); F2("mul", t, d, d); F2("mul", t, t, pb[4 + q * 2 + 1]); F1("add", s, s, lo(t)); F1("add", s, s, hi(t));
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = s0 + s1 + s2 + s3;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += lo(pa[i]) + hi(pa[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double results_per_iter, int ctas_per_sm)
{
    static float *out = nullptr, *in = nullptr; static long long* cyc = nullptr;
    if (!out) { cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&in, 64 * 4); cudaMalloc(&cyc, 148 * 4 * 8); cudaMemset(in, 0, 256); }
    const int iters = 1 << 13, grid = 148 * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, in, iters, cyc);
    k<MODE><<<grid, 256>>>(out, in, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 4]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < grid; ++i) mean += h[i]; mean /= grid;
    printf("%-58s %d CTA/SM: %7.1f results/clk/SM\n", name, ctas_per_sm, results_per_iter * iters * 256.0 * ctas_per_sm / mean);
}

int main()
{
    for (int c = 1; c <= 2; ++c) {
        run<0>("FMUL  reg, imm", 32, c);
        run<1>("FMUL  reg, reg", 32, c);
        run<2>("FADD  reg, reg", 32, c);
        run<3>("FMUL2 reg, reg (2 results each)", 32, c);
        run<4>("FADD2 reg, reg (2 results each)", 32, c);
        run<5>("scorer element scalar (sub mul mul add), regs", 128, c);
        run<6>("scorer element packed (3 packed + 2 FADD per 2), regs", 128, c);
        run<7>("scorer element scalar, x by broadcast LDS.128", 128, c);
        run<8>("scorer element packed, x by broadcast LDS.128", 128, c);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
