// Probe: semantics of cvt.pack.sat.u8.s32.b32 and mad.wide.s32 on sm_100a (profiling aid, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/packsat_probe tools/packsat_probe.cu && /tmp/packsat_probe
#include <cstdio>
#include <cstdint>
__global__ void k(const int *a, const int *b, const unsigned *c, unsigned *d, long long *w, int n)
{
    int i = threadIdx.x;
    if (i >= n) return;
    unsigned r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a[i]), "r"(b[i]), "r"(c[i]));
    d[i] = r;
    long long p;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(p) : "r"(a[i]), "r"(312561664), "l"(0x80000000ll));
    w[i] = p;
}
int main()
{
    const int n = 6;
    int ha[n] = {1, 300, -5, 255, 17, 128}, hb[n] = {2, 7, 260, -1, 200, 64};
    unsigned hc[n] = {0, 0xAABBCCDD, 0x1234, 0xFFFFFFFF, 0x00010002, 0x80};
    int *a, *b; unsigned *c, *d; long long *w;
    cudaMalloc(&a, sizeof(ha)); cudaMalloc(&b, sizeof(hb)); cudaMalloc(&c, sizeof(hc)); cudaMalloc(&d, sizeof(hc)); cudaMalloc(&w, n * 8);
    cudaMemcpy(a, ha, sizeof(ha), cudaMemcpyHostToDevice); cudaMemcpy(b, hb, sizeof(hb), cudaMemcpyHostToDevice); cudaMemcpy(c, hc, sizeof(hc), cudaMemcpyHostToDevice);
    k<<<1, 32>>>(a, b, c, d, w, n);
    unsigned hd[n]; long long hw[n];
    cudaMemcpy(hd, d, sizeof(hd), cudaMemcpyDeviceToHost); cudaMemcpy(hw, w, sizeof(hw), cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) printf("a=%d b=%d c=%08x -> d=%08x   wide=%lld hi=%lld\n", ha[i], hb[i], hc[i], hd[i], hw[i], hw[i] >> 32);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
