// lat_probe.cu -- latency of Montgomery products / EC additions for a LONE warp (the regime of the small-MSM tail kernels)
#include <cstdio>
#include "../halo2-prover_b200/csrc/field.cuh"
#include "../halo2-prover_b200/csrc/curve.cuh"
#include "../halo2-prover_b200/csrc/msm_comb.cuh"
using namespace h2b;

template <int MODE>
__global__ void __launch_bounds__(32) k(const Fe *in, Fe *out, int iters, long long *cyc) {
    Fe a = load_fe(&in[threadIdx.x]), b = load_fe(&in[threadIdx.x + 32]);
    Fe x0 = a, x1 = b, x2 = Fq::add(a, b), x3 = Fq::dbl(b);
    XYZZ p, q;
    p.x = a; p.y = b; p.zz = x2; p.zzz = x3;
    q.x = b; q.y = x2; q.zz = x3; q.zzz = a;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {  // 4 dependent products
            x0 = Fq::mul(x0, b); x0 = Fq::mul(x0, b); x0 = Fq::mul(x0, b); x0 = Fq::mul(x0, b);
        } else if (MODE == 1) {  // 4 independent products
            x0 = Fq::mul(x0, b); x1 = Fq::mul(x1, b); x2 = Fq::mul(x2, b); x3 = Fq::mul(x3, b);
        } else if (MODE == 4) {  // 4 dependent products, 64-bit C arithmetic (no PTX carry flag)
            x0 = Fq::mul_portable(x0, b); x0 = Fq::mul_portable(x0, b); x0 = Fq::mul_portable(x0, b); x0 = Fq::mul_portable(x0, b);
        } else if (MODE == 5) {  // 4 independent products, 64-bit C arithmetic
            x0 = Fq::mul_portable(x0, b); x1 = Fq::mul_portable(x1, b); x2 = Fq::mul_portable(x2, b); x3 = Fq::mul_portable(x3, b);
        } else if (MODE == 6) {  // addition shared by teams of 4 lanes
            p = xyzz_add_team4(p, q, threadIdx.x & 3, 0xffffffffu);
        } else if (MODE == 2) {  // xyzz_add (noinline, 12M + 2S)
            xyzz_add(p, q);
        } else if (MODE == 3) {  // xyzz_dbl_ni
            p = xyzz_dbl_ni(p);
        }
    }
    long long t1 = clock64();
    Fe r = Fq::add(Fq::add(x0, x1), Fq::add(x2, x3));
    r = Fq::add(r, Fq::add(p.x, Fq::add(p.y, Fq::add(p.zz, p.zzz))));
    store_fe(&out[blockIdx.x * 32 + threadIdx.x], r);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    Fe *in, *out;
    long long *cyc, h;
    cudaMalloc(&in, 64 * sizeof(Fe));
    cudaMalloc(&out, 148 * 32 * sizeof(Fe));
    cudaMalloc(&cyc, 8);
    cudaMemset(in, 0x11, 64 * sizeof(Fe));
    const int iters = 200;
    const char *names[] = {"4 dependent modmul", "4 independent modmul", "xyzz_add (14 modmul)", "xyzz_dbl (9 modmul)", "4 dependent portable", "4 independent portable", "xyzz_add_team4"};
    for (int m = 0; m < 7; m++) {
        for (int rep = 0; rep < 2; rep++) {
            if (m == 0) k<0><<<148, 32>>>(in, out, iters, cyc);
            if (m == 1) k<1><<<148, 32>>>(in, out, iters, cyc);
            if (m == 2) k<2><<<148, 32>>>(in, out, iters, cyc);
            if (m == 3) k<3><<<148, 32>>>(in, out, iters, cyc);
            if (m == 4) k<4><<<148, 32>>>(in, out, iters, cyc);
            if (m == 5) k<5><<<148, 32>>>(in, out, iters, cyc);
            if (m == 6) k<6><<<148, 32>>>(in, out, iters, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-24s %8.0f cycles per iteration (lone warp per SM)\n", names[m], (double)h / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
