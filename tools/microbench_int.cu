// Integer-pipe issue-rate microbenchmark for sm_100a (B200): measures the INT32 multiply roofline that
// SURVEY.md section 8(d) asks for (not present in MEASURED_PEAKS.json).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_int tools/microbench_int.cu
//   ./tools/microbench_int  > gpurun_out/microbench_int.json
//
// Each kernel runs ILP independent dependency chains per thread so latency is hidden; rates are
// reported as lane-operations per clock per SM (from in-kernel clock64) and per second (CUDA events).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 1 << 17;

template <int MODE>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t lo[ILP], hi[ILP], z[ILP];
    uint32_t x = seed + threadIdx.x * 2654435761u + blockIdx.x, y = seed * 40503u + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) { lo[i] = x + i * 77u; hi[i] = y ^ (i * 1234567u); z[i] = x * (i + 3u); }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) {  // IMAD (32-bit multiply-add)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(x), "r"(y));
            } else if (MODE == 1) {  // IMAD.WIDE.U32, 64-bit accumulate, no carry
                asm volatile("{.reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t;}"
                             : "+r"(lo[i]), "+r"(hi[i]) : "r"(y));
            } else if (MODE == 8) {  // IMAD.WIDE.U32, multiplicand taken from the neighbouring chain
                asm volatile("{.reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0, %1}, t;}"
                             : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[(i + 1) % ILP]), "r"(y));
            } else if (MODE == 2) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(x), "r"(y));
            } else if (MODE == 4) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(y));
            } else if (MODE == 5) {  // IMAD.WIDE + one independent IADD3 per multiply
                asm volatile("{.reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t;}"
                             : "+r"(lo[i]), "+r"(hi[i]) : "r"(y));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(lo[(i + 4) % ILP]));
            } else if (MODE == 6) {  // IMAD.WIDE + two IADD3 per multiply
                asm volatile("{.reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t;}"
                             : "+r"(lo[i]), "+r"(hi[i]) : "r"(y));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(lo[(i + 4) % ILP]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(y) : "r"(hi[(i + 3) % ILP]));
            } else if (MODE == 7) {  // IMAD + one IADD3 per multiply
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(x), "r"(y));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(hi[i]) : "r"(lo[(i + 4) % ILP]));
            }
        }
        if (MODE == 9) {  // mode 3 plus two LOP3 per wide multiply on unrelated registers (does the ALU pipe co-issue?)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < ILP; i++) asm volatile("xor.b32 %0, %0, %1;" : "+r"(z[i]) : "r"(z[(i + 3) % ILP]));
        }
        if (MODE == 3 || MODE == 9) {  // carry chains as the field multiplier uses them: IMAD.WIDE.U32(.X) with predicate carry
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                "madc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\t"
                "madc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
                : "+r"(lo[0]), "+r"(lo[1]), "+r"(lo[2]), "+r"(lo[3]), "+r"(lo[4]), "+r"(lo[5]), "+r"(lo[6]), "+r"(lo[7])
                : "r"(x), "r"(y), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]));
            asm volatile(
                "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                "madc.lo.cc.u32 %4, %8, %11, %4;\n\tmadc.hi.cc.u32 %5, %8, %11, %5;\n\t"
                "madc.lo.cc.u32 %6, %8, %12, %6;\n\tmadc.hi.u32 %7, %8, %12, %7;"
                : "+r"(hi[0]), "+r"(hi[1]), "+r"(hi[2]), "+r"(hi[3]), "+r"(hi[4]), "+r"(hi[5]), "+r"(hi[6]), "+r"(hi[7])
                : "r"(y), "r"(x), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]));
        }
    }
    long long t1 = clock64();
    uint32_t acc = x ^ y;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= lo[i] ^ hi[i] ^ z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_clock_khz = 1;
template <int MODE>
static void run(const char* name, double mults_per_iter, int sms, int blocks_per_sm, uint32_t* d_out, long long* d_cyc, bool last) {
    int blocks = sms * blocks_per_sm;
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
    double ops_per_block = 256.0 * ITERS * mults_per_iter;
    double per_sec = ops_per_block * blocks / (best * 1e-3);
    // lanes per clock per SM from the SAME CUDA-event time as the T/s column, at the clock the device reports as its maximum
    // (the clocks log of the run shows whether it held it).  The first version of this tool divided by the blocks' own
    // clock64() spans instead, which over-counts when the resident blocks of an SM do not start together (102 vs 64 for IMAD).
    double per_clk_sm = per_sec / ((double)sms * (double)g_clock_khz * 1e3);
    printf("  {\"kernel\": \"%s\", \"mult_lanes_per_clk_per_sm\": %.2f, \"mult_Tops_per_s\": %.3f, \"ms\": %.4f, \"avg_block_cycles\": %.0f}%s\n",
           name, per_clk_sm, per_sec * 1e-12, best, cyc, last ? "" : ",");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    g_clock_khz = p.clockRate;
    int sms = p.multiProcessorCount, bps = 4;   // 4 x 256 threads = 32 warps per SM
    uint32_t* d_out; long long* d_cyc;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * sms * bps * 256));
    CK(cudaMalloc(&d_cyc, sizeof(long long) * sms * bps));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"ilp\": %d, \"warps_per_sm\": %d, \"results\": [\n", p.name, sms, p.clockRate, ILP, bps * 8);
    run<0>("imad_lo (mad.lo.u32)", ILP, sms, bps, d_out, d_cyc, false);
    run<1>("imad_wide (mad.wide.u32)", ILP, sms, bps, d_out, d_cyc, false);
    run<8>("imad_wide pure (multiplicand from neighbour chain)", ILP, sms, bps, d_out, d_cyc, false);
    run<2>("imad_hi (mad.hi.u32)", ILP, sms, bps, d_out, d_cyc, false);
    run<3>("imad_wide_carry_chain (mad.lo.cc/madc.hi.cc pairs)", 8, sms, bps, d_out, d_cyc, false);
    run<9>("imad_wide_carry_chain + 2 LOP3 per multiply (mults counted)", 8, sms, bps, d_out, d_cyc, false);
    run<4>("iadd3 (add.u32; counted as ops)", ILP, sms, bps, d_out, d_cyc, false);
    run<5>("imad_wide + 1 alu op each (mults counted)", ILP, sms, bps, d_out, d_cyc, false);
    run<6>("imad_wide + 2 alu ops each (mults counted)", ILP, sms, bps, d_out, d_cyc, false);
    run<7>("imad_lo + 1 alu op each (mults counted)", ILP, sms, bps, d_out, d_cyc, true);
    printf("]}\n");
    return 0;
}
