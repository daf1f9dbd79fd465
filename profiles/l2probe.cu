// scratch: does B200's L2 keep a buffer of B MB resident between two read passes when (a) all SMs read it,
// (b) only SMs with smid in [lo, hi) read it?  Run under ncu and compare dram__bytes_read of pass 2.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void rd(const uint4 *p, size_t n, int lo, int hi, unsigned long long *sink) {
    unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid < lo || (int)smid >= hi) return;
    // CTAs that run: grid-stride over the buffer using a global ticket so the active CTAs cover everything
    __shared__ unsigned long long base;
    unsigned long long acc = 0;
    for (;;) {
        if (threadIdx.x == 0) base = atomicAdd(sink + 1, (unsigned long long)blockDim.x * 16);
        __syncthreads();
        unsigned long long b = base;
        __syncthreads();
        if (b >= n) break;
        for (int j = 0; j < 16; ++j) { size_t i = b + (size_t)j * blockDim.x + threadIdx.x; if (i < n) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; } }
    }
    if (acc == 0x123456789ull) sink[0] = acc;
}
int main(int argc, char **argv) {
    size_t mb = argc > 1 ? atoi(argv[1]) : 48; int lo = argc > 2 ? atoi(argv[2]) : 0, hi = argc > 3 ? atoi(argv[3]) : 1000;
    size_t n = mb * 1024 * 1024 / 16; uint4 *p; unsigned long long *sink;
    cudaMalloc(&p, n * 16); cudaMemset(p, 1, n * 16); cudaMalloc(&sink, 16);
    for (int pass = 0; pass < 3; ++pass) { cudaMemset(sink, 0, 16); rd<<<148 * 8, 256>>>(p, n, lo, hi, sink); cudaDeviceSynchronize(); }
    printf("done %zu MB smid [%d,%d) %s\n", mb, lo, hi, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
