// Device field arithmetic (PTX carry-chain paths of fe.cuh) against the portable host branches of the same header,
// on random and edge operands (non-canonical representatives, values that trigger the rare second folds).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I bp_pp_b200/csrc tests/cuda/fe_selftest.cu -o ...
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define BPPP_FE_NOINLINE 1
#include "ec.cuh"
#include "sc.cuh"
using namespace bppp;

enum { OP_MUL, OP_SQR, OP_ADD, OP_SUB, OP_NEG, OP_MULINT3, OP_MULINT8, OP_MULINT21, OP_NORM, OP_ISZERO, OP_EXPR, OP_INV, OP_PT, OP_PTJ, OP_PTX, OP_SCMUL, OP_SCSQR, OP_SCWIDE, OP_SCINV, OP_FULL, OP_PTX_INL, OP_PTJ_INL, OP_COUNT };
static const char *NAMES[] = {"mul", "sqr", "add", "sub", "neg", "mul_int3", "mul_int8", "mul_int21", "normalize", "is_zero", "expr(calls)", "inv/sqrt", "pt rcb", "pt jacobian", "pt xyzz", "sc_mul", "sc_sqr", "sc_reduce512", "sc_inv", "straight-line folds", "xyzz inlined", "jacobian inlined"};

__host__ __device__ inline Fe apply(int op, const Fe &a, const Fe &b) {
    switch (op) {
    case OP_MUL: return fe_mul_inl(a, b);
    case OP_SQR: return fe_sqr_inl(a);
    case OP_ADD: return fe_add(a, b);
    case OP_SUB: return fe_sub(a, b, 1);
    case OP_NEG: return fe_negate(a, 1);
    case OP_MULINT3: return fe_mul_int(a, 3);
    case OP_MULINT8: return fe_mul_int(a, 8);
    case OP_MULINT21: return fe_mul_int(a, 21);
    case OP_NORM: return fe_normalize(a);
    case OP_ISZERO: return fe_from_u32(fe_normalizes_to_zero(a) ? 1u : 0u);
    case OP_EXPR: {   // through the noinline fe_mul / fe_sqr device functions, mixed with adds and subs
        Fe t = fe_add(fe_mul(a, b), fe_sqr(a));
        Fe u = fe_sub(t, b, 1);
        Fe w = fe_mul(fe_add(a, b), fe_mul_int(u, 3));
        return fe_sub(fe_sqr(w), fe_mul(fe_negate(t, 1), w), 1);
    }
    case OP_INV: return fe_add(fe_inv(a), fe_sqrt_candidate(b));
    case OP_PT: {
        Pt p; p.x = a; p.y = b; p.z = fe_one();
        PtA q; q.x = b; q.y = a;
        Pt r = pt_add(pt_double(p), pt_add_mixed(p, q));
        return fe_add(fe_add(r.x, r.y), r.z);
    }
    case OP_PTJ: {
        PtJ p; p.x = a; p.y = b; p.z = fe_from_u32(5); p.inf = false;
        PtA q; q.x = b; q.y = a;
        PtJ r = ptj_add_mixed(ptj_double(ptj_double(p)), q);
        return fe_add(fe_add(r.x, r.y), r.z);
    }
    case OP_SCMUL: { Sc x, y; for (int k = 0; k < 8; k++) { x.v[k] = a.v[k]; y.v[k] = b.v[k]; } Sc r = sc_mul(x, y); return fe_from_words(r.v); }
    case OP_SCSQR: { Sc x; for (int k = 0; k < 8; k++) x.v[k] = a.v[k]; Sc r = sc_sqr(x); return fe_from_words(r.v); }
    case OP_SCWIDE: { uint32_t t[16]; for (int k = 0; k < 8; k++) { t[k] = a.v[k]; t[8 + k] = b.v[k]; } Sc r = sc_reduce512(t); return fe_from_words(r.v); }
    case OP_SCINV: { Sc x; for (int k = 0; k < 8; k++) x.v[k] = a.v[k]; x.v[7] &= 0x7FFFFFFFu; Sc r = sc_inv(x); return fe_from_words(r.v); }
    case OP_FULL: {   // the branch-free variants used inside inlined point formulas
        Fe t = fe_add_t<false>(fe_mul_inl_t<false>(a, b), fe_sqr_inl_t<false>(a));
        Fe u = fe_sub_t<false>(fe_mul_int_t<false>(t, 8), b);
        Fe w = fe_sub_t<false>(fe_add_t<false>(a, b), fe_mul_int_t<false>(u, 3));
        return fe_add_t<false>(fe_sub_t<false>(fe_zero(), w), fe_mul_int_t<false>(fe_add_t<false>(a, a), 2));
    }
    case OP_PTX_INL: {
        PtX p = ptx_identity();
        PtA q; q.x = b; q.y = a;
        PtA q2; q2.x = a; q2.y = b;
        p = ptx_add_mixed_t<true>(p, q); p = ptx_add_mixed_t<true>(p, q2); p = ptx_add_mixed_t<true>(p, q2); p = ptx_add_mixed_t<true>(p, q);
        return fe_add(fe_add(p.x, p.y), fe_add(p.zz, p.zzz));
    }
    case OP_PTJ_INL: {
        PtJ p; p.x = a; p.y = b; p.z = fe_from_u32(5); p.inf = false;
        PtA q; q.x = b; q.y = a;
        PtJ r = ptj_add_mixed_t<true>(ptj_double_t<true>(ptj_double_t<true>(p)), q);
        return fe_add(fe_add(r.x, r.y), r.z);
    }
    case OP_PTX: {
        PtX p = ptx_identity();
        PtA q; q.x = b; q.y = a;
        PtA q2; q2.x = a; q2.y = b;
        p = ptx_add_mixed(p, q); p = ptx_add_mixed(p, q2); p = ptx_add_mixed(p, q);
        Pt r = ptx_to_pt(p);
        return fe_add(fe_add(r.x, r.y), r.z);
    }
    }
    return fe_zero();
}
__global__ void k_apply(int op, const Fe *a, const Fe *b, Fe *r, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = apply(op, a[i], b[i]);
}

int main() {
    std::vector<Fe> edge;
    auto mk = [](uint32_t w0, uint32_t w1, uint32_t rest) { Fe f; f.v[0] = w0; f.v[1] = w1; for (int k = 2; k < 8; k++) f.v[k] = rest; return f; };
    const uint32_t F = 0xFFFFFFFFu;
    edge.push_back(mk(0, 0, 0)); edge.push_back(mk(1, 0, 0)); edge.push_back(mk(976, 0, 0)); edge.push_back(mk(977, 0, 0)); edge.push_back(mk(978, 0, 0));
    edge.push_back(mk(976, 1, 0)); edge.push_back(mk(977, 1, 0)); edge.push_back(mk(978, 1, 0));
    edge.push_back(mk(FE_P0 - 1, FE_P1, F)); edge.push_back(mk(FE_P0, FE_P1, F)); edge.push_back(mk(FE_P0 + 1, FE_P1, F)); edge.push_back(mk(F, FE_P1, F));
    edge.push_back(mk(0, F, F)); edge.push_back(mk(F - 977, F, F)); edge.push_back(mk(F - 976, F, F)); edge.push_back(mk(F - 1, F, F)); edge.push_back(mk(F, F, F));
    { Fe f = mk(0, 0, 0); f.v[7] = 0x80000000u; edge.push_back(f); }
    { Fe f = mk(F, F, F); f.v[7] = 0; edge.push_back(f); }
    { Fe f = mk(0x55555555u, 0x55555555u, 0x55555555u); edge.push_back(f); }
    std::vector<Fe> A, B;
    for (auto &x : edge) for (auto &y : edge) { A.push_back(x); B.push_back(y); }
    uint64_t s = 88172645463325252ull;
    for (int i = 0; i < 20000; i++) {
        Fe x, y;
        for (int k = 0; k < 8; k++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x.v[k] = (uint32_t)(s >> 16); s ^= s << 13; s ^= s >> 7; s ^= s << 17; y.v[k] = (uint32_t)(s >> 16); }
        if (i % 7 == 0) for (int k = 2; k < 8; k++) x.v[k] = F;      // near the top of the range
        if (i % 11 == 0) for (int k = 1; k < 8; k++) y.v[k] = (i & 1) ? F : 0;
        A.push_back(x); B.push_back(y);
    }
    int n = (int)A.size();
    Fe *da, *db, *dr;
    if (cudaMalloc(&da, n * sizeof(Fe)) != cudaSuccess) { printf("no CUDA device\n"); return 2; }
    cudaMalloc(&db, n * sizeof(Fe)); cudaMalloc(&dr, n * sizeof(Fe));
    cudaMemcpy(da, A.data(), n * sizeof(Fe), cudaMemcpyHostToDevice); cudaMemcpy(db, B.data(), n * sizeof(Fe), cudaMemcpyHostToDevice);
    std::vector<Fe> R(n);
    int total_bad = 0;
    for (int op = 0; op < OP_COUNT; op++) {
        k_apply<<<(n + 127) / 128, 128>>>(op, da, db, dr, n);
        cudaError_t e = cudaMemcpy(R.data(), dr, n * sizeof(Fe), cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int i = 0; i < n; i++) {
            // device and host may return different representatives of the same class only if an algorithm differs: require identical words
            Fe h = apply(op, A[i], B[i]);
            bool same = true;
            for (int k = 0; k < 8; k++) same &= h.v[k] == R[i].v[k];
            if (!same) { bad++; if (first < 0) first = i; }
        }
        printf("%-10s %d / %d mismatches%s", NAMES[op], bad, n, e == cudaSuccess ? "" : "  CUDA ERROR");
        if (first >= 0) {
            printf("  first at %d: a=", first); for (int k = 7; k >= 0; k--) printf("%08x", A[first].v[k]);
            printf(" b="); for (int k = 7; k >= 0; k--) printf("%08x", B[first].v[k]);
            printf(" dev="); for (int k = 7; k >= 0; k--) printf("%08x", R[first].v[k]);
            Fe h = apply(op, A[first], B[first]);
            printf(" host="); for (int k = 7; k >= 0; k--) printf("%08x", h.v[k]);
        }
        printf("\n");
        total_bad += bad;
    }
    printf("fe_selftest: %s\n", total_bad ? "FAIL" : "ok");
    return total_bad ? 1 : 0;
}
