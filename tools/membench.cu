// membench.cu -- HBM bandwidth of streaming kernels with different read:write mixes (profiling aid).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/membench tools/membench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int R, int W>   // each thread reads R uint4 and writes W uint4 per iteration
__global__ void __launch_bounds__(256) k_mix(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint4 acc = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint4 v = __ldg(src + i + (size_t)r * n);
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint4 o = acc;
            o.x += w;
            __stcs(dst + i + (size_t)w * n, o);
        }
        if (W == 0 && acc.x == 0x12345678u) dst[0] = acc;
    }
}

template <int R, int W>
void run(const char *name, uint4 *a, uint4 *b, size_t n, int blocks)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) k_mix<R, W><<<blocks, 256>>>(a, b, n);
    cudaEventRecord(e0);
    const int iters = 20;
    for (int i = 0; i < iters; i++) k_mix<R, W><<<blocks, 256>>>(a, b, n);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)n * 16 * (R + W) * iters;
    printf("%-22s blocks=%5d  %8.1f GB/s\n", name, blocks, bytes / ms / 1e6);
}

int main()
{
    const size_t n = (size_t)256 << 20 >> 4;     // 256 MiB per stream
    uint4 *a, *b;
    cudaMalloc(&a, n * 16 * 3);
    cudaMalloc(&b, n * 16 * 3);
    cudaMemset(a, 1, n * 16 * 3);
    for (int blocks : {148 * 8, 148 * 16, 148 * 64}) {
        run<1, 1>("copy 1:1", a, b, n, blocks);
        run<1, 2>("expand 1:2", a, b, n, blocks);
        run<1, 3>("expand 1:3", a, b, n, blocks);
        run<2, 1>("shrink 2:1", a, b, n, blocks);
        run<3, 1>("shrink 3:1", a, b, n, blocks);
        run<1, 0>("read only", a, b, n, blocks);
        run<0, 1>("write only", a, b, n, blocks);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
