#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// IMAD.WIDE issue-rate variants: what makes a wide MAD cost 2 vs 4 cycles on the FMA-heavy pipe?
template <int V> __global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a[8], b[8], lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed * (k + 3) + threadIdx.x; b[k] = seed * (7 * k + 1) + blockIdx.x; lo[k] = k; hi[k] = seed + k; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (V == 1) {
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %2, %3, %0; mad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(a[0]), "r"(b[0]));
        } else if (V == 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) { uint64_t c = ((uint64_t)hi[k] << 32) | lo[k]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a[0]), "r"(b[0])); lo[k] = (uint32_t)c; hi[k] = (uint32_t)(c >> 32); }
        } else if (V == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) { uint64_t c = ((uint64_t)hi[k] << 32) | lo[k]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a[k]), "r"(b[0])); lo[k] = (uint32_t)c; hi[k] = (uint32_t)(c >> 32); }
        } else if (V == 3) {
#pragma unroll
            for (int k = 0; k < 8; k++) { uint64_t c = ((uint64_t)hi[k] << 32) | lo[k]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a[k]), "r"(b[k])); lo[k] = (uint32_t)c; hi[k] = (uint32_t)(c >> 32); }
        } else if (V == 4) {   // carry chain, same a, b
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t madc.lo.cc.u32 %2, %8, %9, %2;\n\t madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                         "madc.lo.cc.u32 %4, %8, %9, %4;\n\t madc.hi.cc.u32 %5, %8, %9, %5;\n\t madc.lo.cc.u32 %6, %8, %9, %6;\n\t madc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]) : "r"(a[0]), "r"(b[0]));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0;\n\t madc.hi.cc.u32 %1, %8, %9, %1;\n\t madc.lo.cc.u32 %2, %8, %9, %2;\n\t madc.hi.cc.u32 %3, %8, %9, %3;\n\t"
                         "madc.lo.cc.u32 %4, %8, %9, %4;\n\t madc.hi.cc.u32 %5, %8, %9, %5;\n\t madc.lo.cc.u32 %6, %8, %9, %6;\n\t madc.hi.u32 %7, %8, %9, %7;"
                         : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7]) : "r"(a[0]), "r"(b[0]));
        } else if (V == 5) {   // carry chain, distinct a, same b (the fe_mul row pattern)
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                         "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;"
                         : "+r"(lo[0]), "+r"(hi[0]), "+r"(lo[1]), "+r"(hi[1]), "+r"(lo[2]), "+r"(hi[2]), "+r"(lo[3]), "+r"(hi[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]));
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t madc.hi.cc.u32 %1, %8, %12, %1;\n\t madc.lo.cc.u32 %2, %9, %12, %2;\n\t madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                         "madc.lo.cc.u32 %4, %10, %12, %4;\n\t madc.hi.cc.u32 %5, %10, %12, %5;\n\t madc.lo.cc.u32 %6, %11, %12, %6;\n\t madc.hi.u32 %7, %11, %12, %7;"
                         : "+r"(lo[4]), "+r"(hi[4]), "+r"(lo[5]), "+r"(hi[5]), "+r"(lo[6]), "+r"(hi[6]), "+r"(lo[7]), "+r"(hi[7]) : "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[1]));
        } else if (V == 6) {   // 32-bit IMAD lo only, distinct operands
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(a[k]), "r"(b[k]));
        } else if (V == 7) {   // IMAD.HI only
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(a[k]), "r"(b[0]));
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= lo[k] ^ hi[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int V> void run(const char *name, uint32_t *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, iters = 20000;
    k<V><<<blocks, 256>>>(d, 3, 100);
    float best = 1e9;
    for (int r = 0; r < 3; r++) { float ms; cudaEventRecord(e0); k<V><<<blocks, 256>>>(d, 3, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double ops = 8.0 * iters * blocks * 256;
    printf("%-44s %.3f ms  %.2f T wide-MAC/s\n", name, best, ops / best / 1e9);
}
int main() {
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("mad.wide same a,b", d);
    run<1>("mad.lo+mad.hi pair (no carry) same a,b", d);
    run<2>("mad.wide distinct a, same b", d);
    run<3>("mad.wide distinct a, distinct b", d);
    run<4>("carry chain (cc) same a,b", d);
    run<5>("carry chain (cc) distinct a, same b", d);
    run<6>("mad.lo 32-bit distinct", d);
    run<7>("mad.hi 32-bit distinct a", d);
    return 0;
}
