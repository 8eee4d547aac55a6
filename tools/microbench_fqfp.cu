// Field-multiplier microbenchmark for sm_100a: the integer multiplier (fq_mul, IMAD.WIDE) against the FP64-pipe
// multiplier (fq_mul_fp, DFMA) and mixes of the two, in the launch shape of the equation kernel (128 threads,
// 3 CTAs per SM, multiplier called as a real function).  Also checks fq_mul_fp == fq_mul on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_fqfp tools/microbench_fqfp.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "fq_fp.cuh"

using namespace jjs;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ fq seed_fq(uint32_t s) {
    fq x;
#pragma unroll
    for (int i = 0; i < 8; i++) { s = s * 1664525u + 1013904223u; x.l[i] = s; }
    x.l[7] &= 0x3fffffffu;  // < 2^254 < q
    return x;
}

__global__ void __launch_bounds__(128) k_check(uint32_t* mismatches, int rounds) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fq a = seed_fq(t * 2 + 1), b = seed_fq(t * 2 + 2);
    uint32_t bad = 0;
    for (int r = 0; r < rounds; r++) {
        fq x, y;
        fq_mul(x, a, b);
        fq_mul_fp(y, a, b);
        if (!fq_eq(x, y)) bad++;
        a = b;
        b = x;
    }
    if (bad) atomicAdd(mismatches, bad);
}

// NI integer multiplications then NF floating-point ones per iteration, on NI + NF independent running products
template <int NI, int NF>
__global__ void __launch_bounds__(128, 3) k_bench(uint32_t* out, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fq v[NI + NF + 1];
#pragma unroll
    for (int i = 0; i < NI + NF + 1; i++) v[i] = seed_fq(t * 16 + i);
    const fq y = seed_fq(t + 77);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NI; i++) fq_mul(v[i], v[i], y);
#pragma unroll
        for (int i = NI; i < NI + NF; i++) fq_mul_fp(v[i], v[i], y);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NI + NF; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc ^= v[i].l[k];
    out[t] = acc;
}

// NP fused (integer + floating-point) pairs and NI further integer multiplications per iteration
template <int NP, int NI>
__global__ void __launch_bounds__(128, 3) k_bench_pair(uint32_t* out, int iters) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fq v[2 * NP + NI + 1];
#pragma unroll
    for (int i = 0; i < 2 * NP + NI + 1; i++) v[i] = seed_fq(t * 16 + i);
    const fq y = seed_fq(t + 77);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NP; i++) fq_mul_pair(v[2 * i], v[2 * i], y, v[2 * i + 1], v[2 * i + 1], y);
#pragma unroll
        for (int i = 2 * NP; i < 2 * NP + NI; i++) fq_mul(v[i], v[i], y);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 2 * NP + NI; i++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc ^= v[i].l[k];
    out[t] = acc;
}

template <int NP, int NI>
void run_pair(const char* name, uint32_t* d_out, int blocks, int iters, bool last) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_bench_pair<NP, NI><<<blocks, 128>>>(d_out, iters / 4);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_bench_pair<NP, NI><<<blocks, 128>>>(d_out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double muls = (double)blocks * 128 * iters * (2 * NP + NI);
    printf("  {\"mix\": \"%s\", \"pairs_per_iter\": %d, \"int_per_iter\": %d, \"ms\": %.3f, \"G_mul_per_s\": %.2f}%s\n", name, NP, NI, best, muls / (best * 1e-3) * 1e-9,
           last ? "" : ",");
}

template <int NI, int NF>
void run(const char* name, uint32_t* d_out, int blocks, int iters, bool last) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_bench<NI, NF><<<blocks, 128>>>(d_out, iters / 4);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        k_bench<NI, NF><<<blocks, 128>>>(d_out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double muls = (double)blocks * 128 * iters * (NI + NF);
    printf("  {\"mix\": \"%s\", \"int_per_iter\": %d, \"fp_per_iter\": %d, \"ms\": %.3f, \"G_mul_per_s\": %.2f}%s\n", name, NI, NF, best, muls / (best * 1e-3) * 1e-9,
           last ? "" : ",");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int blocks = p.multiProcessorCount * 3 * 4;
    uint32_t *d_out, *d_bad, bad = 0;
    CK(cudaMalloc(&d_out, sizeof(uint32_t) * blocks * 128));
    CK(cudaMalloc(&d_bad, 4));
    CK(cudaMemset(d_bad, 0, 4));
    k_check<<<blocks, 128>>>(d_bad, 64);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    const int iters = 2000;
    printf("{\"device\": \"%s\", \"checked\": %d, \"mismatches\": %u, \"results\": [\n", p.name, blocks * 128 * 64, bad);
    run<4, 0>("int only", d_out, blocks, iters, false);
    run<0, 4>("fp only", d_out, blocks, iters, false);
    run<3, 1>("3 int + 1 fp", d_out, blocks, iters, false);
    run<2, 1>("2 int + 1 fp", d_out, blocks, iters, false);
    run<1, 1>("1 int + 1 fp", d_out, blocks, iters, false);
    run<5, 1>("5 int + 1 fp", d_out, blocks, iters, false);
    run_pair<2, 0>("fused pairs only", d_out, blocks, iters, false);
    run_pair<1, 1>("1 fused pair + 1 int", d_out, blocks, iters, false);
    run_pair<1, 2>("1 fused pair + 2 int", d_out, blocks, iters, true);
    printf("]}\n");
    return bad ? 1 : 0;
}
