#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// Can the FP64 pipe run beside the FMA-heavy integer pipe?  V=0: wide MADs only; V=1: DFMA only; V=2: both interleaved.
template <int V> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8], lo[8], hi[8];
    double x[8], acc[8];
    uint32_t b = seed * 7 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed * (k + 3) + threadIdx.x; lo[k] = k; hi[k] = seed + k; x[k] = 1.0 + 1e-9 * (threadIdx.x + k); acc[k] = 0.5 * k; }
    double y = 1.0 + 1e-10 * seed;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (V == 0 || V == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) { uint64_t c = ((uint64_t)hi[k] << 32) | lo[k]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a[k]), "r"(b)); lo[k] = (uint32_t)c; hi[k] = (uint32_t)(c >> 32); }
        }
        if (V == 1 || V == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[k]) : "d"(x[k]), "d"(y));
        }
    }
    uint32_t r = 0; double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) { r ^= lo[k] ^ hi[k]; s += acc[k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r ^ (uint32_t)__double2loint(s);
}
template <int V> void run(const char *name, uint32_t *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, iters = 20000;
    k<V><<<blocks, 256>>>(d, 3, 100);
    float best = 1e9;
    for (int r = 0; r < 3; r++) { float ms; cudaEventRecord(e0); k<V><<<blocks, 256>>>(d, 3, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double ops = 8.0 * iters * blocks * 256;
    printf("%-36s %.3f ms  %.2f T (8 ops of each kind per iteration)/s\n", name, best, ops / best / 1e9);
}
int main() {
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("mad.wide.u32 only", d);
    run<1>("fma.rn.f64 only", d);
    run<2>("both interleaved", d);
    return 0;
}
