// batch_affine_probe.cu -- what would bucket accumulation cost with affine additions and a shared inversion?
//
// Not product code: an upper-bound probe for DESIGN.md section 7.  Each thread owns K independent pairs (P_i, Q_i)
// gathered from a table of affine points by random index (like the accumulate kernel's gathers) and adds them:
//   pass 1: d_i = x2 - x1, prefix products p_i = p_(i-1) * d_i written to a per-thread array in global memory;
//   one field inversion of p_K (Fermat);
//   pass 2 (backwards): 1/d_i = inv * p_(i-1), inv *= d_i, lambda = (y2 - y1)/d_i, x3 = lambda^2 - x1 - x2,
//           y3 = lambda (x1 - x3) - y1, result stored (64 B).
// 6 products per addition + 1 for the prefix + the inversion's ~320 / K.  The baseline is the XYZZ mixed addition
// of the accumulate kernel (10 products) over the same gathers, one running sum per thread.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/bap scripts/batch_affine_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../halo2-prover_b200/csrc/curve.cuh"
using namespace h2b;

__global__ void __launch_bounds__(128, 4)
batch_affine_kernel(const Affine *__restrict__ table, const uint32_t *__restrict__ idx, uint32_t K, Fe *__restrict__ prefix,
                    Affine *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t *my = idx + (size_t)t * 2 * K;
    const size_t T = (size_t)gridDim.x * blockDim.x;  // prefix[i][t], out[i][t]: coalesced across the warp
    Fe *pp = prefix + t;
    Fe acc = Fq::one();
    for (uint32_t i = 0; i < K; i++) {
        const Fe x1 = load_fe_ro(&table[my[2 * i]].x), x2 = load_fe_ro(&table[my[2 * i + 1]].x);
        store_fe(&pp[(size_t)i * T], acc);        // p_(i-1)
        acc = Fq::mul(acc, Fq::sub(x2, x1));
    }
    Fe inv = Fq::inv(acc);
    for (int i = (int)K - 1; i >= 0; i--) {
        const Affine p = load_affine(&table[my[2 * i]]), q = load_affine(&table[my[2 * i + 1]]);
        const Fe d = Fq::sub(q.x, p.x);
        const Fe di = Fq::mul(inv, load_fe(&pp[(size_t)i * T]));
        inv = Fq::mul(inv, d);
        const Fe lam = Fq::mul(Fq::sub(q.y, p.y), di);
        Affine r;
        r.x = Fq::sub(Fq::sub(Fq::sqr(lam), p.x), q.x);
        r.y = Fq::sub(Fq::mul(lam, Fq::sub(p.x, r.x)), p.y);
        store_fe(&out[(size_t)i * T + t].x, r.x);
        store_fe(&out[(size_t)i * T + t].y, r.y);
    }
}

__global__ void __launch_bounds__(128, 4)
xyzz_madd_kernel(const Affine *__restrict__ table, const uint32_t *__restrict__ idx, uint32_t K, XYZZ *__restrict__ out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t *my = idx + (size_t)t * K;
    XYZZ acc = xyzz_identity();
    Affine p = load_affine(&table[my[0]]);
    for (uint32_t i = 0; i < K; i++) {
        Affine pn = p;
        if (i + 1 < K) pn = load_affine(&table[my[i + 1]]);
        xyzz_madd(acc, p);
        p = pn;
    }
    store_xyzz(&out[t], acc);
}

int main(int argc, char **argv) {
    const uint32_t lg_table = 24;
    const size_t ntab = (size_t)1 << lg_table;
    Affine *table;
    cudaMalloc(&table, ntab * sizeof(Affine));
    {
        std::vector<uint32_t> h(ntab * 16);
        uint64_t s = 88172645463325252ull;
        for (auto &v : h) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            v = (uint32_t)s & 0x0fffffffu;  // < q in every limb: valid field elements (not curve points; cost is identical)
        }
        cudaMemcpy(table, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    }
    const uint32_t threads = 148 * 4 * 128 * 4;  // 4 full waves
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (uint32_t K : {32u, 64u, 128u, 256u, 512u}) {
        const size_t pairs = (size_t)threads * K;
        uint32_t *idx;
        Fe *prefix;
        Affine *out;
        XYZZ *xo;
        cudaMalloc(&idx, pairs * 2 * 4);
        cudaMalloc(&prefix, pairs * sizeof(Fe));
        cudaMalloc(&out, pairs * sizeof(Affine));
        cudaMalloc(&xo, (size_t)threads * 2 * sizeof(XYZZ));
        {
            std::vector<uint32_t> h(pairs * 2);
            uint64_t s = 1234567 + K;
            for (auto &v : h) {
                s ^= s << 13; s ^= s >> 7; s ^= s << 17;
                v = (uint32_t)(s >> 20) & (uint32_t)(ntab - 1);
            }
            cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        }
        float ms_b = 0, ms_x = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            batch_affine_kernel<<<threads / 128, 128>>>(table, idx, K, prefix, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_b, e0, e1);
            cudaEventRecord(e0);
            xyzz_madd_kernel<<<threads / 128, 128>>>(table, idx, K, xo);  // same number of additions: threads x K
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms_x, e0, e1);
        }
        printf("K = %3u: batch-affine %8.3f ms (%.2f G add/s)   XYZZ mixed %8.3f ms for %zu additions (%.2f G add/s)   ratio %.2f   %s\n",
               K, ms_b, pairs / ms_b * 1e-6, ms_x, pairs, pairs / ms_x * 1e-6, ms_x / ms_b, cudaGetErrorString(cudaGetLastError()));
        cudaFree(idx);
        cudaFree(prefix);
        cudaFree(out);
        cudaFree(xo);
    }
    return 0;
}
