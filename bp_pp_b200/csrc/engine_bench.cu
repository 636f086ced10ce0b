// libbppp.so, microbenchmark translation unit: integer-pipe and field/point op throughput (bppp_microbench).
#include "engine_common.cuh"

using namespace bppp;

static int fail(int code, const std::string &msg) { return engine_fail(code, msg); }

// ---- microbenchmarks ----
// MODE 0: mad.wide.u32 (IMAD.WIDE.U32, 32x32+64 -> 64)   1: mad.lo.u32 (IMAD)   2: add.u32 (IADD3)
// inline PTX so that exactly these instructions issue; 8 independent chains per thread, 64 warps per SM
template <int MODE>
__global__ void __launch_bounds__(256) k_mb_imad(uint64_t *out, uint32_t seed, int iters) {
    uint64_t acc[8];
    uint32_t lo[8];
    uint32_t x[8], y = seed | 1u;
#pragma unroll
    for (int k = 0; k < 8; k++) { acc[k] = (uint64_t)(threadIdx.x + k) * 0x9E3779B97F4A7C15ULL; x[k] = seed * (2 * k + 3) + threadIdx.x; lo[k] = x[k] ^ y; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(x[k]), "r"(y));
                if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[k]) : "r"(x[k]), "r"(y));
                if (MODE == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[k]) : "r"(x[k]));
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k] ^ lo[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void __launch_bounds__(64) k_mb_op(uint32_t *out, uint32_t seed, int iters) {
    Fe a = fe_from_u32(seed + threadIdx.x), b = fe_from_u32(seed * 3 + 1 + blockIdx.x);
    a.v[3] = threadIdx.x + 5; b.v[7] = blockIdx.x + 9; b.v[5] = 77;
    Pt p; p.x = a; p.y = b; p.z = fe_from_u32(1);
    PtA q; q.x = b; q.y = a;
    Sc s, u;
#pragma unroll
    for (int k = 0; k < 8; k++) { s.v[k] = seed * (k + 1) + threadIdx.x; u.v[k] = seed + 7 * k + blockIdx.x; }
    s.v[7] &= 0x7FFFFFFFu; u.v[7] &= 0x7FFFFFFFu;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (OP == 1) a = fe_mul(a, b);
        if (OP == 2) a = fe_sqr(a);
        if (OP == 3) s = sc_mul(s, u);
        if (OP == 4) p = pt_add_mixed(p, q);
        if (OP == 5) p = pt_double(p);
        if (OP == 6) p = pt_add(p, p);
        if (OP == 7) a = fe_add(fe_inv(a), b);             // safegcd (modinv.cuh)
        if (OP == 8) a = fe_add(fe_inv_fermat(a), b);      // a^(p-2)
        if (OP == 9) s = sc_add(sc_inv(s), u);
        if (OP == 10) s = sc_add(sc_inv_fermat(s), u);
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < FE_W; k++) r ^= a.v[k] ^ p.x.v[k] ^ p.y.v[k] ^ p.z.v[k];
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= s.v[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// ---- microbench ----
extern "C" int bppp_microbench(int device, double *out, int n_out) {
    if (!out || n_out < 8) return fail(BPPP_ERR_ARG, "need 8 outputs");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(BPPP_ERR_NO_DEVICE, "no CUDA device");
    CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, device));
    int sms = prop.multiProcessorCount;
    uint64_t *d64 = nullptr; uint32_t *d32 = nullptr;
    const int blocks = sms * 8;
    CUDA_OK(cudaMalloc(&d64, sizeof(uint64_t) * blocks * 256));
    CUDA_OK(cudaMalloc(&d32, sizeof(uint32_t) * blocks * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time_ms = [&](auto launch) -> float {
        launch(); cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        return best;
    };
    {
        const int iters = 2000;
        float ms = time_ms([&] { k_mb_imad<0><<<blocks, 256>>>(d64, 12345u, iters); });
        out[0] = (double)blocks * 256 * iters * 64 / (ms * 1e-3);
        if (n_out >= 10) {
            ms = time_ms([&] { k_mb_imad<1><<<blocks, 256>>>(d64, 12345u, iters); });
            out[8] = (double)blocks * 256 * iters * 64 / (ms * 1e-3);
            ms = time_ms([&] { k_mb_imad<2><<<blocks, 256>>>(d64, 12345u, iters); });
            out[9] = (double)blocks * 256 * iters * 64 / (ms * 1e-3);
        }
    }
    const int opblocks = sms * 16;
    auto run_op = [&](int op, int iters) -> double {
        float ms = 0;
        switch (op) {
            case 1: ms = time_ms([&] { k_mb_op<1><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 2: ms = time_ms([&] { k_mb_op<2><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 3: ms = time_ms([&] { k_mb_op<3><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 4: ms = time_ms([&] { k_mb_op<4><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 5: ms = time_ms([&] { k_mb_op<5><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 6: ms = time_ms([&] { k_mb_op<6><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 7: ms = time_ms([&] { k_mb_op<7><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 8: ms = time_ms([&] { k_mb_op<8><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            case 9: ms = time_ms([&] { k_mb_op<9><<<opblocks, 64>>>(d32, 777u, iters); }); break;
            default: ms = time_ms([&] { k_mb_op<10><<<opblocks, 64>>>(d32, 777u, iters); }); break;
        }
        return (double)opblocks * 64 * iters / (ms * 1e-3);
    };
    out[1] = run_op(1, 4000); out[2] = run_op(2, 4000); out[3] = run_op(3, 2000);
    out[4] = run_op(4, 400); out[5] = run_op(5, 400); out[6] = run_op(6, 400);
    if (n_out >= 14) { out[10] = run_op(7, 20); out[11] = run_op(8, 20); out[12] = run_op(9, 20); out[13] = run_op(10, 20); }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, device);
    out[7] = clk / 1000.0;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d64); cudaFree(d32);
    CUDA_OK(cudaGetLastError());
    return BPPP_OK;
}
