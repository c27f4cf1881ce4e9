// pipe_probe.cu -- issue-rate microbenchmarks for the sm_100a pipes the field arithmetic can use.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pipe_probe scripts/pipe_probe.cu
// Prints, per instruction mix, lane-ops per clock per SM (clock64 deltas, max over blocks).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

enum Mix { IMAD, IMADWIDE, IMADHI, DFMA, DADD, IADD3, IADDC, WIDE_DFMA_1_1, WIDE_DFMA_1_2, WIDE_DFMA_1_3, WIDE_DFMA_IADD,
           WIDE_DFMA2_IADD2, LOP3, SHF, DFMA_IADD_1_1, DFMA_IADD_1_2, IMAD_DFMA_1_1, I2D, NMIX };
static const char *names[] = {"imad.lo", "imad.wide", "imad.hi", "dfma.rz", "dadd.rz", "iadd3", "iadd.cc/addc", "wide+dfma 1:1",
                              "wide+dfma 1:2", "wide+dfma 1:3", "wide+dfma+iadd 1:1:1", "wide+2dfma+2iadd", "lop3", "shf", "dfma+iadd 1:1",
                              "dfma+iadd 1:2", "imad.lo+dfma 1:1", "cvt.f64.u32"};

template <int MIX>
__global__ void __launch_bounds__(256) probe(uint32_t *out, unsigned long long *cyc, uint32_t seed) {
    uint32_t a[8], b = seed | 1, c = seed * 3 + 1;
    uint64_t w[8];
    double d[8], e = 1.0 + seed * 1e-9, f = 0.5 + seed * 1e-9;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = threadIdx.x + i + seed;
        w[i] = a[i];
        d[i] = (double)a[i];
    }
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MIX == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MIX == IMADWIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"(c));
            if (MIX == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MIX == DFMA) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e), "d"(f));
            if (MIX == DADD) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e));
            if (MIX == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MIX == IADDC) {
                uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(lo), "+r"(hi) : "r"(b), "r"(c));
                w[i] = ((uint64_t)hi << 32) | lo;
            }
            if (MIX == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MIX == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MIX == I2D) {
                asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[i]) : "r"(a[i]));
                a[i] += 1;
            }
            if (MIX == WIDE_DFMA_1_1 || MIX == WIDE_DFMA_1_2 || MIX == WIDE_DFMA_1_3 || MIX == WIDE_DFMA_IADD ||
                MIX == WIDE_DFMA2_IADD2) {
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"(c));
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e), "d"(f));
                if (MIX == WIDE_DFMA_1_2 || MIX == WIDE_DFMA_1_3 || MIX == WIDE_DFMA2_IADD2)
                    asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[(i + 4) & 7]) : "d"(f), "d"(e));
                if (MIX == WIDE_DFMA_1_3) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[(i + 2) & 7]) : "d"(f), "d"(f));
                if (MIX == WIDE_DFMA_IADD || MIX == WIDE_DFMA2_IADD2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (MIX == WIDE_DFMA2_IADD2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 4) & 7]) : "r"(c));
            }
            if (MIX == DFMA_IADD_1_1 || MIX == DFMA_IADD_1_2) {
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e), "d"(f));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
                if (MIX == DFMA_IADD_1_2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 4) & 7]) : "r"(c));
            }
            if (MIX == IMAD_DFMA_1_1) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
                asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(e), "d"(f));
            }
        }
    }
    unsigned long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double_as_longlong(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int ops_per_inner(int mix) {
    switch (mix) {
        case IADDC: return 2;
        case WIDE_DFMA_1_1: return 2;
        case WIDE_DFMA_1_2: return 3;
        case WIDE_DFMA_1_3: return 4;
        case WIDE_DFMA_IADD: return 3;
        case WIDE_DFMA2_IADD2: return 5;
        case DFMA_IADD_1_1: return 2;
        case DFMA_IADD_1_2: return 3;
        case IMAD_DFMA_1_1: return 2;
        case I2D: return 2;
        default: return 1;
    }
}

template <int MIX>
void run(uint32_t *out, unsigned long long *cyc, int blocks) {
    probe<MIX><<<blocks, 256>>>(out, cyc, 7);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MIX><<<blocks, 256>>>(out, cyc, 9);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[4096], mx = 0;
    cudaMemcpy(h, cyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    for (int i = 0; i < blocks; i++) mx = h[i] > mx ? h[i] : mx;
    // 4 blocks of 256 threads per SM -> 1024 lanes per SM issue ITERS*8*ops each
    double lane_ops_per_sm = 1024.0 * ITERS * 8 * ops_per_inner(MIX);
    double total = lane_ops_per_sm * (blocks / 4);
    printf("%-24s %8.2f lane-ops/clk/SM (clock64)   %8.1f Gops/s (events, %.3f ms)   cycles/warp-instr/SMSP %.2f\n", names[MIX],
           lane_ops_per_sm / (double)mx, total / (ms * 1e6), ms, (double)mx / (ITERS * 8.0 * ops_per_inner(MIX) * 8.0));
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 4;
    printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaMalloc(&cyc, blocks * 8);
    run<IMAD>(out, cyc, blocks);
    run<IMADWIDE>(out, cyc, blocks);
    run<IMADHI>(out, cyc, blocks);
    run<DFMA>(out, cyc, blocks);
    run<DADD>(out, cyc, blocks);
    run<IADD3>(out, cyc, blocks);
    run<IADDC>(out, cyc, blocks);
    run<LOP3>(out, cyc, blocks);
    run<SHF>(out, cyc, blocks);
    run<I2D>(out, cyc, blocks);
    run<WIDE_DFMA_1_1>(out, cyc, blocks);
    run<WIDE_DFMA_1_2>(out, cyc, blocks);
    run<WIDE_DFMA_1_3>(out, cyc, blocks);
    run<WIDE_DFMA_IADD>(out, cyc, blocks);
    run<WIDE_DFMA2_IADD2>(out, cyc, blocks);
    run<DFMA_IADD_1_1>(out, cyc, blocks);
    run<DFMA_IADD_1_2>(out, cyc, blocks);
    run<IMAD_DFMA_1_1>(out, cyc, blocks);
    return 0;
}
