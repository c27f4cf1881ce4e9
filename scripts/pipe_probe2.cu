// pipe_probe2.cu -- do warps doing FP64-pipe work and warps doing IMAD.WIDE work co-run at full rate on one SM?
// Half the warps of every block run an IMAD.WIDE stream (the 32-bit-limb Montgomery product), the other half a
// DFMA/DADD/IADD stream in the ratio of a double-precision (52-bit-limb) Montgomery product.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

// mode bit0: run I-warps, bit1: run F-warps.  F_ALU = integer adds per 3 FP64 ops.
template <int F_ALU>
__global__ void __launch_bounds__(256) corun(uint32_t *out, unsigned long long *cyc, uint32_t seed, int iters_i, int iters_f, int mode,
                                             int split) {
    const int warp = threadIdx.x >> 5;
    // split: how many of the 8 warps are I-warps
    const bool is_i = warp < split;
    uint32_t b = seed | 1, c = seed * 3 + 1;
    uint64_t w[8];
    uint32_t lo[8], hi[8];
    double d[8], e = 1.0 + seed * 1e-9, f = 0.5 + seed * 1e-9;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        w[i] = threadIdx.x + i + seed;
        lo[i] = (uint32_t)w[i];
        hi[i] = seed;
        d[i] = (double)w[i];
    }
    __syncthreads();
    unsigned long long t0 = clock64();
    if (is_i) {
        if (mode & 1) {
#pragma unroll 1
            for (int it = 0; it < iters_i; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"(c));
            }
        }
    } else {
        if (mode & 2) {
#pragma unroll 1
            for (int it = 0; it < iters_f; it++) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    double t;
                    asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(t) : "d"(e), "d"(f), "d"(d[i]));
                    asm volatile("sub.rz.f64 %0, %1, %0;" : "+d"(d[i]) : "d"(t));
                    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e), "d"(f));
                    if (F_ALU >= 2)
                        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(b), "r"(c));
                    if (F_ALU >= 4)
                        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;"
                                     : "+r"(lo[(i + 4) & 7]), "+r"(hi[(i + 4) & 7])
                                     : "r"(c), "r"(b));
                }
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= lo[i] ^ hi[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double_as_longlong(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 8 + warp] = t1 - t0;
}

template <int F_ALU>
void run(uint32_t *out, unsigned long long *cyc, int blocks, int iters_i, int iters_f, int mode, int split) {
    corun<F_ALU><<<blocks, 256>>>(out, cyc, 7, iters_i, iters_f, mode, split);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    corun<F_ALU><<<blocks, 256>>>(out, cyc, 9, iters_i, iters_f, mode, split);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    static unsigned long long h[8 * 4096];
    cudaMemcpy(h, cyc, blocks * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long mi = 0, mf = 0;
    for (int i = 0; i < blocks * 8; i++) {
        if ((i & 7) < split) mi = h[i] > mi ? h[i] : mi;
        else mf = h[i] > mf ? h[i] : mf;
    }
    // per SMSP: 4 blocks x 8 warps / 4 SMSP = 8 warps, `split` of them I-warps
    double wide_per_smsp = (mode & 1) ? (double)split * iters_i * 8 : 0;
    double prod_per_smsp = (mode & 2) ? (double)(8 - split) * iters_f * 8 : 0;  // one "52x52 product" = 3 FP64 + F_ALU int adds
    printf("F_ALU=%d split=%d mode=%d: %.3f ms | I-warps %llu cyc (%.2f cyc per wide/SMSP) | F-warps %llu cyc (%.2f cyc per 52x52 product/SMSP)\n",
           F_ALU, split, mode, ms, mi, mi ? mi / wide_per_smsp : 0.0, mf, mf ? mf / prod_per_smsp : 0.0);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 4;
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaMalloc(&cyc, blocks * 8 * 8);
    // alone
    run<4>(out, cyc, blocks, 4096, 0, 1, 4);
    run<4>(out, cyc, blocks, 0, 2048, 2, 4);
    run<2>(out, cyc, blocks, 0, 2048, 2, 4);
    run<0>(out, cyc, blocks, 0, 2048, 2, 4);
    // together, tuned so that both finish at about the same time if they overlap perfectly
    run<4>(out, cyc, blocks, 4096, 2048, 3, 4);
    run<2>(out, cyc, blocks, 4096, 2048, 3, 4);
    run<0>(out, cyc, blocks, 4096, 2048, 3, 4);
    run<4>(out, cyc, blocks, 4096, 1024, 3, 4);
    run<2>(out, cyc, blocks, 4096, 1536, 3, 4);
    run<4>(out, cyc, blocks, 4096, 4096, 3, 2);
    run<4>(out, cyc, blocks, 4096, 1024, 3, 6);
    return 0;
}
