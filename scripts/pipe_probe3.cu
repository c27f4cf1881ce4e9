// pipe_probe3.cu -- is DFMA throughput register-file bound when its operands are distinct registers?
// and does the exact 5-instruction limb-product pattern of field52.cuh overlap its FP64 and ALU halves?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(128) probe(double *out, const double *in, int iters) {
    double a[8], b[8], d[8];
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = in[threadIdx.x + i];
        b[i] = in[threadIdx.x + 8 + i];
        d[i] = in[threadIdx.x + 16 + i];
        acc[i] = i;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) d[i] = __fma_rz(a[i], b[(i + 3) & 7], d[i]);                  // 3 distinct register operands
            if (MODE == 1) d[i] = __fma_rz(a[0], b[0], d[i]);                            // two operands reused
            if (MODE == 2) d[i] = __fma_rz(a[i], b[(i + 3) & 7], 0x1p104) + d[i] * 0.0;  // placeholder (not used)
            if (MODE == 3) {  // the limb product: hi, sub, lo, two 3-input 64-bit adds
                const double h = __fma_rz(a[i], b[(i + 3) & 7], 0x1p104);
                const double s = (0x1p104 + 0x1p52) - h;
                const double l = __fma_rz(a[i], b[(i + 3) & 7], s);
                acc[i] += (uint64_t)__double_as_longlong(h) + (uint64_t)__double_as_longlong(l);
            }
            if (MODE == 4) {  // FP64 half only
                const double h = __fma_rz(a[i], b[(i + 3) & 7], 0x1p104);
                const double s = (0x1p104 + 0x1p52) - h;
                d[i] = __fma_rz(a[i], b[(i + 3) & 7], s) + d[i];
            }
            if (MODE == 5) {  // ALU half only: 3-input 64-bit add
                acc[i] += acc[(i + 1) & 7] + (uint64_t)__double_as_longlong(a[i]);
            }
            if (MODE == 6) {  // hi only + one 2-input 64-bit add
                const double h = __fma_rz(a[i], b[(i + 3) & 7], 0x1p104);
                acc[i] += (uint64_t)__double_as_longlong(h);
            }
        }
        if (MODE == 3 || MODE == 6) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __longlong_as_double((long long)((acc[i] & 0xfffffffffffffull) | 0x4330000000000000ull)) - 0x1p52;
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r += d[i] + (double)acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char *name, double *out, double *in, int blocks, double fp64_per, double alu_per) {
    probe<MODE><<<blocks, 128>>>(out, in, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 128>>>(out, in, ITERS);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double warps_per_smsp = blocks / (double)p.multiProcessorCount;  // 4 warps per block, 4 SMSPs
    double steps = warps_per_smsp * ITERS * 8.0;                      // inner statements per SMSP
    double cyc = ms * 1e-3 * p.clockRate * 1e3;
    printf("%-44s warps/SMSP %.0f: %.3f ms, %.2f cycles per statement per SMSP (FP64 instrs %.0f, ALU instrs %.0f)\n", name, warps_per_smsp, ms,
           cyc / steps, fp64_per, alu_per);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *in, *out;
    cudaMalloc(&in, 4096 * 8);
    cudaMalloc(&out, p.multiProcessorCount * 8 * 128 * 8);
    cudaMemset(in, 0x3f, 4096 * 8);
    for (int w = 2; w <= 8; w *= 2) {
        int blocks = p.multiProcessorCount * w;
        run<0>("dfma, 3 distinct register operands", out, in, blocks, 1, 0);
        run<1>("dfma, a and b reused", out, in, blocks, 1, 0);
        run<4>("limb product, FP64 half (dfma,dadd,dfma,dadd)", out, in, blocks, 4, 0);
        run<5>("limb product, ALU half (iadd3 + iadd3.x)", out, in, blocks, 0, 2);
        run<6>("hi only (dfma) + 64-bit add", out, in, blocks, 1, 2);
        run<3>("limb product (dfma,dadd,dfma + iadd3,iadd3.x)", out, in, blocks, 3, 2);
    }
    return 0;
}
