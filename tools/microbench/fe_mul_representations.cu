#define BPPP_FE_NOINLINE 1
#include "fe26_legacy.cuh"   // the 10 x 26-bit lazy-limb field of commit 59e3110 (git show 59e3110:bp_pp_b200/csrc/fe.cuh)
#include <cstdio>
#include <cuda_runtime.h>
using namespace bppp;
struct Fe8 { uint32_t v[8]; };

// acc[0..8) += {a0,a1,a2,a3} * b (pairs), carry out -> *top (fresh)
__device__ __forceinline__ void chain4_fresh_carry(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t &top) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
// acc[0..7) += ..., acc[7] = hi + carry (fresh top limb)
__device__ __forceinline__ void chain4_fresh_top(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b, uint32_t &top) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "=r"(top)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}
__device__ __forceinline__ void mul4(uint32_t *acc, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b) {
    asm("mul.lo.u32 %0, %8, %12; mul.hi.u32 %1, %8, %12; mul.lo.u32 %2, %9, %12; mul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12; mul.hi.u32 %5, %10, %12; mul.lo.u32 %6, %11, %12; mul.hi.u32 %7, %11, %12;"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
}

__device__ __forceinline__ Fe8 fe8_mul_inl(const Fe8 &a, const Fe8 &b) {
    uint32_t E[16], O[16];
    E[8] = 0;
    mul4(E, a.v[0], a.v[2], a.v[4], a.v[6], b.v[0]);
    mul4(O, a.v[1], a.v[3], a.v[5], a.v[7], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (i & 1) {
            chain4_fresh_carry(O + i - 1, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], O[i + 7]);
            chain4_fresh_top(E + i + 1, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i], E[i + 8]);
        } else {
            chain4_fresh_carry(E + i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i], E[i + 8]);
            chain4_fresh_top(O + i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i], O[i + 7]);
        }
    }
    // T = E + (O << 32)
    uint32_t T[16];
    T[0] = E[0];
    asm("add.cc.u32 %0, %1, %2;" : "=r"(T[1]) : "r"(E[1]), "r"(O[0]));
#pragma unroll
    for (int k = 2; k < 15; k++) asm("addc.cc.u32 %0, %1, %2;" : "=r"(T[k]) : "r"(E[k]), "r"(O[k - 1]));
    asm("addc.u32 %0, %1, %2;" : "=r"(T[15]) : "r"(E[15]), "r"(O[14]));
    // R = L + H*977 (+ H<<32)
    uint32_t R8, R9;
    chain4_fresh_carry(T, T[8], T[10], T[12], T[14], 977u, R8);
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(R8)
        : "r"(T[9]), "r"(T[11]), "r"(T[13]), "r"(T[15]), "r"(977u));
    asm("add.cc.u32 %0, %0, %9;\n\t addc.cc.u32 %1, %1, %10;\n\t addc.cc.u32 %2, %2, %11;\n\t addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t addc.cc.u32 %5, %5, %14;\n\t addc.cc.u32 %6, %6, %15;\n\t addc.cc.u32 %7, %7, %16;\n\t addc.u32 %8, 0, 0;"
        : "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "+r"(R8), "=r"(R9)
        : "r"(T[8]), "r"(T[9]), "r"(T[10]), "r"(T[11]), "r"(T[12]), "r"(T[13]), "r"(T[14]), "r"(T[15]));
    // fold V = R8 + R9 2^32 (< 2^34):  W = V * (2^32 + 977)
    uint32_t W0, W1, W2;
    asm("mul.lo.u32 %0, %3, 977;\n\t mul.hi.u32 %1, %3, 977;\n\t"
        "mad.lo.u32 %1, %4, 977, %1;\n\t"     // R9 * 977 < 2^12, hi(R8*977) < 2^10: no overflow
        "add.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %4, 0;"
        : "=&r"(W0), "=&r"(W1), "=&r"(W2) : "r"(R8), "r"(R9));
    uint32_t c2;
    asm("add.cc.u32 %0, %0, %9;\n\t addc.cc.u32 %1, %1, %10;\n\t addc.cc.u32 %2, %2, %11;\n\t addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t addc.cc.u32 %5, %5, 0;\n\t addc.cc.u32 %6, %6, 0;\n\t addc.cc.u32 %7, %7, 0;\n\t addc.u32 %8, 0, 0;"
        : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]), "+r"(T[3]), "+r"(T[4]), "+r"(T[5]), "+r"(T[6]), "+r"(T[7]), "=r"(c2)
        : "r"(W0), "r"(W1), "r"(W2));
    // wrapped past 2^256: the residue is below 2^67, add 2^32 + 977 once more (cannot carry past limb 2)
    asm("mad.lo.cc.u32 %0, %3, 977, %0;\n\t addc.cc.u32 %1, %1, %3;\n\t addc.u32 %2, %2, 0;" : "+r"(T[0]), "+r"(T[1]), "+r"(T[2]) : "r"(c2));
    Fe8 r;
#pragma unroll
    for (int k = 0; k < 8; k++) r.v[k] = T[k];
    return r;
}


__device__ __noinline__ Fe8 fe8_mul(Fe8 a, Fe8 b) { return fe8_mul_inl(a, b); }
struct Fe8x2 { Fe8 p, q; };
__device__ __noinline__ Fe8x2 fe8_mul2(Fe8 a, Fe8 b, Fe8 c, Fe8 d) { Fe8x2 r; r.p = fe8_mul_inl(a, b); r.q = fe8_mul_inl(c, d); return r; }
template <int BPS> __global__ void __launch_bounds__(64, BPS) k_occ(uint32_t *p, int n) {
    int t = blockIdx.x * 64 + threadIdx.x;
    Fe8 a, b, c;
    for (int k = 0; k < 8; k++) { a.v[k] = p[t * 8 + k]; b.v[k] = p[(t ^ 1) * 8 + k]; c.v[k] = p[(t ^ 2) * 8 + k]; }
#pragma unroll 1
    for (int i = 0; i < n; i++) { a = fe8_mul(a, b); c = fe8_mul(c, a); }
    for (int k = 0; k < 8; k++) p[t * 8 + k] = a.v[k] ^ c.v[k];
}
template <int BPS> __global__ void __launch_bounds__(64, BPS) k_dual(uint32_t *p, int n) {
    int t = blockIdx.x * 64 + threadIdx.x;
    Fe8 a, b, c;
    for (int k = 0; k < 8; k++) { a.v[k] = p[t * 8 + k]; b.v[k] = p[(t ^ 1) * 8 + k]; c.v[k] = p[(t ^ 2) * 8 + k]; }
#pragma unroll 1
    for (int i = 0; i < n; i++) { Fe8x2 r = fe8_mul2(a, b, c, a); a = r.p; c = r.q; }
    for (int k = 0; k < 8; k++) p[t * 8 + k] = a.v[k] ^ c.v[k];
}
__global__ void __launch_bounds__(64, 7) k_new(uint32_t *p, int n) {
    int t = blockIdx.x * 64 + threadIdx.x;
    Fe8 a, b, c;
    for (int k = 0; k < 8; k++) { a.v[k] = p[t * 8 + k]; b.v[k] = p[(t ^ 1) * 8 + k]; c.v[k] = p[(t ^ 2) * 8 + k]; }
#pragma unroll 1
    for (int i = 0; i < n; i++) { a = fe8_mul(a, b); c = fe8_mul(c, a); }
    for (int k = 0; k < 8; k++) p[t * 8 + k] = a.v[k] ^ c.v[k];
}
__global__ void __launch_bounds__(64, 7) k_old(uint32_t *p, int n) {
    int t = blockIdx.x * 64 + threadIdx.x;
    Fe a = fe_from_words(p + t * 8), b = fe_from_words(p + (t ^ 1) * 8), c = fe_from_words(p + (t ^ 2) * 8);
#pragma unroll 1
    for (int i = 0; i < n; i++) { a = fe_mul(a, b); c = fe_mul(c, a); }
    uint32_t w[8], w2[8]; fe_to_words(w, fe_normalize(a)); fe_to_words(w2, fe_normalize(c));
    for (int k = 0; k < 8; k++) p[t * 8 + k] = w[k] ^ w2[k];
}
// correctness: both paths, canonical output
__global__ void k_check(const uint32_t *in, uint32_t *out_new, uint32_t *out_old, int n) {
    int t = blockIdx.x * 64 + threadIdx.x;
    Fe8 a, b;
    for (int k = 0; k < 8; k++) { a.v[k] = in[t * 8 + k]; b.v[k] = in[(t ^ 1) * 8 + k]; }
    Fe a10 = fe_from_words(a.v), b10 = fe_from_words(b.v);
    for (int i = 0; i < n; i++) { a = fe8_mul(a, b); a10 = fe_mul(a10, b10); }
    Fe c = fe_normalize(fe_from_words(a.v));
    fe_to_words(out_new + t * 8, c);
    fe_to_words(out_old + t * 8, fe_normalize(a10));
}
int main() {
    const int blocks = 148 * 7, T = blocks * 64, n = 4000;
    uint32_t *h = (uint32_t *)malloc(T * 32), *d, *o1, *o2;
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < T * 8; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s >> 16); }
    // edge rows: all-ones (>= p), p itself, p-1, 0
    for (int k = 0; k < 8; k++) { h[k] = 0xFFFFFFFFu; h[8 + k] = 0xFFFFFFFFu; h[16 + k] = (k == 0 ? 0xFFFFFC2Fu : k == 1 ? 0xFFFFFFFEu : 0xFFFFFFFFu); h[24 + k] = 0xFFFFFFFFu; h[32 + k] = 0; h[40 + k] = 0xFFFFFFFFu; }
    cudaMalloc(&d, T * 32); cudaMalloc(&o1, T * 32); cudaMalloc(&o2, T * 32);
    cudaMemcpy(d, h, T * 32, cudaMemcpyHostToDevice);
    k_check<<<blocks, 64>>>(d, o1, o2, 37);
    uint32_t *r1 = (uint32_t *)malloc(T * 32), *r2 = (uint32_t *)malloc(T * 32);
    cudaMemcpy(r1, o1, T * 32, cudaMemcpyDeviceToHost); cudaMemcpy(r2, o2, T * 32, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < T * 8; i++) bad += r1[i] != r2[i];
    printf("check: %d mismatching words of %d (err %s)\n", bad, T * 8, cudaGetErrorString(cudaGetLastError()));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        float ms;
        cudaEventRecord(e0); k_new<<<blocks, 64>>>(d, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("new 8x32: %.3f ms  %.2f G fe_mul/s\n", ms, 2.0 * n * T / ms / 1e6);
        cudaEventRecord(e0); k_old<<<blocks, 64>>>(d, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("old 10x26: %.3f ms  %.2f G fe_mul/s\n", ms, 2.0 * n * T / ms / 1e6);
    }

#define OCC(BPS) { float ms; int bl = 148 * BPS; cudaEventRecord(e0); k_occ<BPS><<<bl, 64>>>(d, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); \
    printf("occ %2d blocks/SM: %.3f ms %.2f G fe_mul/s\n", BPS, ms, 2.0 * n * bl * 64 / ms / 1e6); \
    cudaEventRecord(e0); k_dual<BPS><<<bl, 64>>>(d, n); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); \
    printf("dual %2d blocks/SM: %.3f ms %.2f G fe_mul/s (%s)\n", BPS, ms, 2.0 * n * bl * 64 / ms / 1e6, cudaGetErrorString(cudaGetLastError())); }
    free(h); h = (uint32_t *)malloc(148 * 32 * 64 * 32); cudaFree(d); cudaMalloc(&d, 148 * 32 * 64 * 32); cudaMemset(d, 0x5a, 148 * 32 * 64 * 32);
    for (int rep = 0; rep < 2; rep++) { OCC(4) OCC(7) OCC(10) OCC(14) OCC(20) OCC(28) }
    return 0;
}
