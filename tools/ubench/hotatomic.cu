// Microbenchmark: returning atomicAdd (one per warp) on K hot addresses from the whole GPU — what does a
// per-lane list counter cost?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hotatomic hotatomic.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int RET>
__global__ void __launch_bounds__(256) k(int* ctr, int K, int stride, int iters, int* out)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int acc = 0;
    for (int i = 0; i < iters; ++i) {
        int* p = ctr + (size_t)((warp + i) % K) * stride;
        if ((threadIdx.x & 31) == 0) {
            if (RET) acc += atomicAdd(p, 32);          // dependent: the next iteration waits for the value
            else atomicAdd(p, 32);
        }
        acc = __shfl_sync(0xffffffffu, acc, 0);
    }
    if (acc == 0x7fffffff) out[0] = acc;
}
int main()
{
    int *ctr, *out;
    cudaMalloc(&ctr, 64 << 20); cudaMalloc(&out, 4);
    cudaMemset(ctr, 0, 64 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%6s %8s %22s %22s\n", "K", "stride B", "returning Mop/s (ns/op/addr)", "fire&forget Mop/s");
    const int G = 148 * 4, iters = 200;
    for (int K : {1, 16, 128, 1024, 16384}) for (int stride : {1, 128}) {
        float ms0, ms1;
        k<1><<<G, 256>>>(ctr, K, stride, iters, out);
        cudaEventRecord(e0); k<1><<<G, 256>>>(ctr, K, stride, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms0, e0, e1);
        cudaEventRecord(e0); k<0><<<G, 256>>>(ctr, K, stride, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms1, e0, e1);
        const double ops = (double)G * 8 * iters;
        printf("%6d %8d %12.1f (%6.1f) %22.1f\n", K, stride * 4, ops / ms0 / 1e3, ms0 * 1e6 / (ops / K), ops / ms1 / 1e3);
    }
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
