// Throughput of the field multiplier in its calling conventions, at the occupancies the kernels run at:
//   reg   : operands in registers, non-inlined multiplier (what k_decode / k_challenge / k_equation call)
//   smem  : operands in shared-memory slots (csrc/fqs.cuh)
//   smem2 : two independent products per call (one call, two interleavable instruction streams)
// Prints G products/s per (mode, resident CTAs per SM).  Peak for reference: 8.89 T IMAD.WIDE/s / 112 per product = 79.4 G/s.
#include <cstdio>
#include <cuda_runtime.h>
#include "../jubjub_schnorr_b200/csrc/fqs.cuh"
using namespace jjs;

__device__ __noinline__ void fqs_mul2_fn(fqh d0, fqh a0, fqh b0, fqh d1, fqh a1, fqh b1) {
    fq x0, y0, x1, y1, r0, r1;
    fqs_ld(x0, a0); fqs_ld(y0, b0); fqs_ld(x1, a1); fqs_ld(y1, b1);
    uint32_t t0[16], t1[16];
    mul_wide(t0, x0.l, y0.l);
    mul_wide(t1, x1.l, y1.l);
    redc(r0.l, t0);
    redc(r1.l, t1);
    fqs_st(d0, r0); fqs_st(d1, r1);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_bench(int iters, const fq* in, fq* out) {
    const Eq2Slots s = eq2_slots(threadIdx.x);
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    fq a = in[t], b = in[t + gridDim.x * blockDim.x];
    if (MODE == 0) {
        fq c = b, d = a;
        for (int i = 0; i < iters; i++) { fq_mul(a, a, b); fq_mul(c, c, d); }
        fq_add(a, a, c);
        out[t] = a;
    } else {
        fqs_st(s.X, a); fqs_st(s.Y, b); fqs_st(s.Z, b); fqs_st(s.T, a);
        if (MODE == 1) for (int i = 0; i < iters; i++) { fqs_mul(s.X, s.X, s.Y); fqs_mul(s.Z, s.Z, s.T); }
        else for (int i = 0; i < iters; i++) fqs_mul2_fn(s.X, s.X, s.Y, s.Z, s.Z, s.T);
        fq x, z;
        fqs_ld(x, s.X); fqs_ld(z, s.Z);
        fq_add(x, x, z);
        out[t] = x;
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000;
    fq *in, *out;
    size_t maxthreads = (size_t)sms * 16 * 128;
    cudaMalloc(&in, sizeof(fq) * 2 * maxthreads);
    cudaMalloc(&out, sizeof(fq) * maxthreads);
    cudaMemset(in, 0x5a, sizeof(fq) * 2 * maxthreads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"reg", "smem", "smem2"};
    unsigned long long check[3] = {0, 0, 0};
    for (int mode = 0; mode < 3; mode++) {
        for (int ctas = 2; ctas <= 8; ctas++) {
            // dynamic shared memory sized so that exactly `ctas` CTAs fit on an SM
            size_t smem = (size_t)(227 * 1024) / ctas - 1024;
            smem = smem / 1024 * 1024;
            if (smem < sizeof(uint4) * 2 * EQ2_SLOTS * 128) continue;
            void (*k)(int, const fq*, fq*) = mode == 0 ? k_bench<0> : (mode == 1 ? k_bench<1> : k_bench<2>);
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 128, smem);
            int grid = sms * occ;
            k<<<grid, 128, smem>>>(100, in, out);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k<<<grid, 128, smem>>>(iters, in, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            double prods = 2.0 * iters * (double)grid * 128;
            fq h;
            cudaMemcpy(&h, out, sizeof(fq), cudaMemcpyDeviceToHost);
            check[mode] = ((unsigned long long)h.l[1] << 32) | h.l[0];
            printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"want\": %d, \"ms\": %.3f, \"Gprod_per_s\": %.2f, \"frac_of_peak\": %.3f, \"err\": \"%s\"}\n", names[mode], occ, ctas, ms,
                   prods / ms / 1e6, prods / ms / 1e6 / 79.4, cudaGetErrorString(cudaGetLastError()));
        }
    }
    printf("checks %llx %llx %llx (must agree)\n", check[0], check[1], check[2]);
    return 0;
}
