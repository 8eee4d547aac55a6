// FP64-pipe microbenchmark for sm_100a (B200): does DFMA issue beside IMAD.WIDE, and at what rate?
// Decides whether a floating-point limb product (52-bit limbs, fma.rz.f64 high/low halves) can share the
// field multiplier's work with the integer pipe.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_fp64 tools/microbench_fp64.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ITERS = 1 << 15;

#define WIDE_CHAIN(L, X, Y, H0, H1, H2) \
    asm volatile( \
        "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t" \
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t" \
        "madc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\t" \
        "madc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;" \
        : "+r"(L[0]), "+r"(L[1]), "+r"(L[2]), "+r"(L[3]), "+r"(L[4]), "+r"(L[5]), "+r"(L[6]), "+r"(L[7]) \
        : "r"(X), "r"(Y), "r"(H0), "r"(H1), "r"(H2))

// MODE 0: 16 DFMA / iter           MODE 1: 16 IMAD.WIDE (2 carry chains) / iter
// MODE 2: 16 DFMA + 16 WIDE        MODE 3: 16 DADD
// MODE 4: 16 DFMA + 16 64-bit integer adds (add.cc/addc)       MODE 5: 16 x 64-bit integer adds
// MODE 6: 8 DFMA + 16 WIDE         MODE 7: 32 DFMA + 16 WIDE
template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t lo[8], hi[8];
    double d[16];
    unsigned long long q[8];
    uint32_t x = seed + threadIdx.x * 2654435761u + blockIdx.x, y = seed * 40503u + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) { lo[i] = x + i * 77u; hi[i] = y ^ (i * 1234567u); q[i] = x * (i + 3ull); }
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = 1.0 + (double)(x & 1023) * 1e-9 * (i + 1);
    double m1 = 1.0 + 1e-12 * (y & 7), m2 = 1e-13 * (x & 15);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
        constexpr int ND = (MODE == 0 || MODE == 2 || MODE == 4) ? 16 : (MODE == 6 ? 8 : (MODE == 7 ? 32 : 0));
        constexpr bool W = (MODE == 1 || MODE == 2 || MODE == 6 || MODE == 7);
        if (W) WIDE_CHAIN(lo, x, y, hi[0], hi[1], hi[2]);
#pragma unroll
        for (int i = 0; i < ND / 2; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i % 16]) : "d"(m1), "d"(m2));
        if (MODE == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(lo[7]));   // (two bare chains in a loop send ptxas 12.9 into a spin)
        if (W) WIDE_CHAIN(hi, y, x, lo[4], lo[5], lo[6]);
#pragma unroll
        for (int i = ND / 2; i < ND; i++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i % 16]) : "d"(m1), "d"(m2));
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(m2));
        }
        if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.u64 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 3) % 8]));
        }
    }
    long long t1 = clock64();
    uint32_t acc = x ^ y;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= lo[i] ^ hi[i] ^ (uint32_t)q[i] ^ (uint32_t)(q[i] >> 32);
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= (uint32_t)__double_as_longlong(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, double dfma, double wide, double other, int sms, int bps, uint32_t* d_out, long long* d_cyc, bool last) {
    int blocks = sms * bps;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) bench<MODE><<<blocks, 256>>>(d_out, 12345u + w, d_cyc);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        bench<MODE><<<blocks, 256>>>(d_out, 999u + rep, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    long long* h = (long long*)malloc(sizeof(long long) * blocks);
    CK(cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < blocks; i++) cyc += (double)h[i]; cyc /= blocks;
    free(h);
    // cycles per iteration per SM sub-partition (each SMSP holds bps*8/4 warps)
    double warps_per_smsp = bps * 8 / 4.0;
    double cyc_per_iter_per_warp_slot = cyc / ITERS / warps_per_smsp;   // issue cycles one warp-iteration costs the SMSP
    double tot = (double)blocks * 256.0 * ITERS;
    printf("  {\"kernel\": \"%s\", \"dfma_per_iter\": %.0f, \"wide_per_iter\": %.0f, \"other_per_iter\": %.0f, \"smsp_cycles_per_warp_iter\": %.2f, \"dfma_T_per_s\": %.3f, \"wide_T_per_s\": %.3f, \"ms\": %.4f}%s\n",
           name, dfma, wide, other, cyc_per_iter_per_warp_slot, tot * dfma / (best * 1e-3) * 1e-12, tot * wide / (best * 1e-3) * 1e-12, best, last ? "" : ",");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount, bps = 4;
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * bps * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * bps));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"warps_per_sm\": %d, \"results\": [\n", p.name, sms, p.clockRate, bps * 8);
    run<0>("dfma only", 16, 0, 0, sms, bps, d_out, d_cyc, false);
    run<2>("16 dfma + 16 wide", 16, 16, 0, sms, bps, d_out, d_cyc, false);
    run<6>("8 dfma + 16 wide", 8, 16, 0, sms, bps, d_out, d_cyc, false);
    run<7>("32 dfma + 16 wide", 32, 16, 0, sms, bps, d_out, d_cyc, false);
    run<3>("dadd only", 0, 0, 16, sms, bps, d_out, d_cyc, false);
    run<4>("16 dfma + 16 add.u64", 16, 0, 16, sms, bps, d_out, d_cyc, false);
    run<5>("16 add.u64 only", 0, 0, 16, sms, bps, d_out, d_cyc, true);
    printf("]}\n");
    return 0;
}
