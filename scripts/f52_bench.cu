// f52_bench.cu -- throughput of the FP64-pipe Montgomery product (field52.cuh) against the IMAD one (field.cuh).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/f52_bench scripts/f52_bench.cu
#include <cfenv>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../halo2-prover_b200/csrc/field.cuh"
#include "field52.cuh"
#include "karatsuba_experiment.cuh"
using namespace h2b;

#define CHAINS 4

__global__ void __launch_bounds__(128) k_mul_i(const Fe *in, Fe *out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x[CHAINS], y = load_fe(&in[t]);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = load_fe(&in[t + 1 + c]);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = Fq::mul(x[c], y);
    }
    Fe r = x[0];
#pragma unroll
    for (int c = 1; c < CHAINS; c++) r = Fq::add(r, x[c]);
    store_fe(&out[t], r);
}

__global__ void __launch_bounds__(128) k_mul_k(const Fe *in, Fe *out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fe x[CHAINS], y = load_fe(&in[t]);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = load_fe(&in[t + 1 + c]);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = mul_karatsuba<FqP>(x[c], y);
    }
    Fe r = x[0];
#pragma unroll
    for (int c = 1; c < CHAINS; c++) r = Fq::add(r, x[c]);
    store_fe(&out[t], r);
}

__host__ __device__ inline void f_chain(N52 (&x)[CHAINS], const N52 &y, int iters) {
    const D52 yd = f52_to_d(y);
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int it = 0; it < iters; it++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < CHAINS; c++) x[c] = (c & 1) ? f52_sqr(f52_to_d(x[c])) : f52_mul(f52_to_d(x[c]), yd);
    }
}
__global__ void __launch_bounds__(128) k_mul_f(const N52 *in, N52 *out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    N52 x[CHAINS], y = in[t];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = in[t + 1 + c];
    f_chain(x, y, iters);
#pragma unroll
    for (int c = 0; c < CHAINS; c++) out[(size_t)t * CHAINS + c] = x[c];
}
__global__ void __launch_bounds__(128) k_mulonly_f(const N52 *in, N52 *out, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    N52 x[CHAINS], y = in[t];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = in[t + 1 + c];
    const D52 yd = f52_to_d(y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = f52_mul(f52_to_d(x[c]), yd);
    }
#pragma unroll
    for (int c = 0; c < CHAINS; c++) out[(size_t)t * CHAINS + c] = x[c];
}

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

int main(int argc, char **argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int only = argc > 2 ? atoi(argv[2]) : 0;
    for (int bpsm = 1; bpsm <= 4; bpsm++) {
        if (only && bpsm != only) continue;
        int blocks = p.multiProcessorCount * bpsm, threads = blocks * 128;
        std::vector<Fe> hi(threads + 8);
        std::vector<N52> hf(threads + 8);
        for (auto &e : hi) {
            for (int i = 0; i < 8; i++) e.l[i] = (uint32_t)rnd();
            e.l[7] &= 0x0fffffff;
        }
        for (auto &e : hf) {
            for (int i = 0; i < 5; i++) e.l[i] = rnd() & kMask52;
            e.l[4] &= (1ull << 46) - 1;  // < 2^254
        }
        Fe *di, *doi;
        N52 *df, *dof;
        cudaMalloc(&di, hi.size() * sizeof(Fe));
        cudaMalloc(&doi, hi.size() * sizeof(Fe));
        cudaMalloc(&df, hf.size() * sizeof(N52));
        cudaMalloc(&dof, hf.size() * sizeof(N52) * CHAINS);
        cudaMemcpy(di, hi.data(), hi.size() * sizeof(Fe), cudaMemcpyHostToDevice);
        cudaMemcpy(df, hf.data(), hf.size() * sizeof(N52), cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float ms_i, ms_f, ms_m;
        k_mul_i<<<blocks, 128>>>(di, doi, 10);
        k_mul_f<<<blocks, 128>>>(df, dof, 10);
        k_mulonly_f<<<blocks, 128>>>(df, dof, 10);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_mul_i<<<blocks, 128>>>(di, doi, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms_i, e0, e1);
        {
            float ms_k;
            std::vector<Fe> ri(threads), rk(threads);
            cudaMemcpy(ri.data(), doi, threads * sizeof(Fe), cudaMemcpyDeviceToHost);
            k_mul_k<<<blocks, 128>>>(di, doi, 10);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_mul_k<<<blocks, 128>>>(di, doi, iters);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms_k, e0, e1);
            cudaMemcpy(rk.data(), doi, threads * sizeof(Fe), cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int t = 0; t < threads; t++)
                for (int i = 0; i < 8; i++) bad += ri[t].l[i] != rk[t].l[i];
            printf("%d blocks/SM: Karatsuba form %.3f ms (%.1f G modmul/s), ratio to IMAD form %.3f, results %s\n", bpsm, ms_k,
                   (double)threads * CHAINS * iters / ms_k * 1e-6, ms_i / ms_k, bad ? "DIFFER" : "identical");
        }
        cudaEventRecord(e0);
        k_mulonly_f<<<blocks, 128>>>(df, dof, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms_m, e0, e1);
        cudaEventRecord(e0);
        k_mul_f<<<blocks, 128>>>(df, dof, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms_f, e0, e1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
        double muls = (double)threads * CHAINS * iters;
        printf("%d blocks/SM x128 thr: IMAD form %.3f ms (%.1f G modmul/s) | FP64 form mul %.3f ms (%.1f G/s) | mul+sqr mix %.3f ms (%.1f G/s) | ratio %.2f / %.2f\n",
               bpsm, ms_i, muls / ms_i * 1e-6, ms_m, muls / ms_m * 1e-6, ms_f, muls / ms_f * 1e-6, ms_i / ms_m, ms_i / ms_f);
        // device == host (same source, std::fma under round-toward-zero) on the first few threads
        std::vector<N52> got(16 * CHAINS);
        cudaMemcpy(got.data(), dof, got.size() * sizeof(N52), cudaMemcpyDeviceToHost);
        fesetround(FE_TOWARDZERO);
        int bad = 0;
        for (int t = 0; t < 16; t++) {
            N52 x[CHAINS];
            for (int c = 0; c < CHAINS; c++) x[c] = hf[t + 1 + c];
            f_chain(x, hf[t], iters);
            for (int c = 0; c < CHAINS; c++)
                for (int i = 0; i < 5; i++) bad += x[c].l[i] != got[t * CHAINS + c].l[i];
        }
        fesetround(FE_TONEAREST);
        printf("   device vs host chain of %d products: %s\n", iters, bad ? "MISMATCH" : "identical");
        cudaFree(di); cudaFree(doi); cudaFree(df); cudaFree(dof);
    }
    return 0;
}
