// ntt.cuh -- radix-2^S shared-memory NTT passes over BN254 Fr for sm_100a.
//
// Device replacement for halo2_proofs @6b43b6b src/arithmetic.rs:185-290 (best_fft,
// recursive_butterfly_arithmetic) and the scaling loops around it in
// src/poly/domain.rs (ifft divisor, distribute_powers_zeta, zero-extension,
// truncation), which the reference reaches from create_proof
// (/root/reference/circuits/src/utils.rs:83-91, :105-120).
//
// Contract kept from the reference: natural order in, natural order out,
// X[K] = sum_n x[n] * omega^(n*K), Montgomery in / Montgomery out, fully reduced.
// The algorithm is NOT the reference's (bit-reverse + radix-2 DIT sweeps over the
// whole array).  It is an autosort (Stockham) decimation-in-frequency transform
// split into at most four passes; each pass stages a [2^S rows] x [C columns] tile
// in shared memory with 128-bit coalesced loads of C adjacent elements per row,
// runs S butterfly levels on chip, multiplies by the inter-pass twiddle from a
// cached table of powers of omega, and stores C adjacent elements per output row.
// The scaling steps of the domain transforms are fused into the first load
// (coset powers, zero padding) and the last store (1/n, inverse coset powers,
// truncation), so every transform costs exactly its passes and nothing else.
//
// Pass t (Ns = product of earlier radices, R = 2^S, M = N / R), for q in [0, M):
//   in : y[q + r*M]                                   r in [0, R)
//   out: y'[(q / Ns) * Ns * R + (q mod Ns) + Ns * K]  K in [0, R)
//        = omega^(Ns * (q / Ns) * K) * sum_r y[q + r*M] * (omega^M)^(r*K)
// The last pass has q / Ns == 0 (no twiddles) and writes exactly the cells it
// read, so it may run in place.
#pragma once
#include "field.cuh"

namespace h2b {

struct NttIo {
    uint32_t n_in;   // input elements present (the rest of the 2^log_n domain reads as 0)
    uint32_t n_out;  // output elements kept (truncation)
    uint32_t pro;    // 1: multiply input i by pro_c[i % 3] (i % 3 == 0 untouched)
    uint32_t epi;    // 1: multiply output i by epi_c[i % 3]
    Fe pro_c[3];
    Fe epi_c[3];
};

// W[e] = omega^e for e in [0, n): two small tables then one product per entry.
__global__ void ntt_pow_small_kernel(Fe omega, uint32_t lo_bits, uint32_t n_lo, uint32_t n_hi,
                                     Fe *tbl_lo, Fe *tbl_hi) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lo) store_fe(&tbl_lo[i], Fr::pow_u64(omega, i));
    if (i < n_hi) store_fe(&tbl_hi[i], Fr::pow_u64(omega, (uint64_t)i << lo_bits));
}
__global__ void ntt_pow_table_kernel(const Fe *__restrict__ tbl_lo, const Fe *__restrict__ tbl_hi,
                                     uint32_t lo_bits, uint32_t n, Fe *__restrict__ W) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    Fe lo = load_fe_ro(&tbl_lo[e & ((1u << lo_bits) - 1)]);
    Fe hi = load_fe_ro(&tbl_hi[e >> lo_bits]);
    store_fe(&W[e], Fr::mul(lo, hi));
}

template <int S>
H2B_DI uint32_t bitrev_s(uint32_t u) {
    return __brev(u) >> (32 - S);
}

// One pass.  S = log2 radix, C = columns per tile, NT = threads per block.
template <int S, int C, int NT>
__global__ void __launch_bounds__(NT)
ntt_pass_kernel(const Fe *in, Fe *out, const Fe *__restrict__ W, uint32_t log_n, uint32_t log_ns,
                uint32_t last, NttIo io) {
    constexpr int R = 1 << S;
    constexpr int TILE = R * C;
    extern __shared__ uint4 smem_u4[];
    uint4 *s_lo = smem_u4;              // low 16 bytes of tile elements  [R][C]
    uint4 *s_hi = smem_u4 + TILE;       // high 16 bytes
    uint4 *t_lo = smem_u4 + 2 * TILE;   // inner twiddles (omega^M)^t, t < R/2
    uint4 *t_hi = t_lo + (R / 2 > 0 ? R / 2 : 1);

    const uint32_t M = 1u << (log_n - S);        // columns in the whole pass
    const uint32_t q0 = blockIdx.x * C;
    const uint32_t tid = threadIdx.x;

    // inner twiddles: W[t * M]
    for (uint32_t t = tid; t < R / 2; t += NT) {
        const uint4 *p = reinterpret_cast<const uint4 *>(&W[(size_t)t * M]);
        t_lo[t] = __ldg(p);
        t_hi[t] = __ldg(p + 1);
    }
    // load tile: element (r, col) <- in[q0 + col + r*M]
    for (uint32_t e = tid; e < TILE; e += NT) {
        uint32_t col = e % C, r = e / C;
        uint32_t idx = q0 + col + r * M;
        Fe v;
        if (idx < io.n_in) {
            v = load_fe(&in[idx]);
            if (io.pro) {
                uint32_t m3 = idx % 3;
                if (m3) v = Fr::mul(v, io.pro_c[m3]);
            }
        } else {
            v = Fr::zero();
        }
        s_lo[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        s_hi[e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
    __syncthreads();

    // S decimation-in-frequency levels, natural order in -> bit-reversed rows out
#pragma unroll 1
    for (int lvl = 0; lvl < S; lvl++) {
        const uint32_t h = (uint32_t)R >> (lvl + 1);  // half span in rows
        for (uint32_t b = tid; b < (uint32_t)(TILE / 2); b += NT) {
            uint32_t col = b % C, i = b / C;           // i in [0, R/2)
            uint32_t lo_i = i & (h - 1);
            uint32_t row0 = ((i - lo_i) << 1) + lo_i;
            uint32_t e0 = row0 * C + col, e1 = e0 + h * C;
            uint4 a0 = s_lo[e0], a1 = s_hi[e0], b0 = s_lo[e1], b1 = s_hi[e1];
            Fe x, y;
            x.l[0] = a0.x; x.l[1] = a0.y; x.l[2] = a0.z; x.l[3] = a0.w;
            x.l[4] = a1.x; x.l[5] = a1.y; x.l[6] = a1.z; x.l[7] = a1.w;
            y.l[0] = b0.x; y.l[1] = b0.y; y.l[2] = b0.z; y.l[3] = b0.w;
            y.l[4] = b1.x; y.l[5] = b1.y; y.l[6] = b1.z; y.l[7] = b1.w;
            Fe sum = Fr::add(x, y);
            Fe dif = Fr::sub(x, y);
            uint32_t tw = lo_i << lvl;  // exponent of omega^M, < R/2
            if (tw != 0) {
                uint4 w0 = t_lo[tw], w1 = t_hi[tw];
                Fe w;
                w.l[0] = w0.x; w.l[1] = w0.y; w.l[2] = w0.z; w.l[3] = w0.w;
                w.l[4] = w1.x; w.l[5] = w1.y; w.l[6] = w1.z; w.l[7] = w1.w;
                dif = Fr::mul(dif, w);
            }
            s_lo[e0] = make_uint4(sum.l[0], sum.l[1], sum.l[2], sum.l[3]);
            s_hi[e0] = make_uint4(sum.l[4], sum.l[5], sum.l[6], sum.l[7]);
            s_lo[e1] = make_uint4(dif.l[0], dif.l[1], dif.l[2], dif.l[3]);
            s_hi[e1] = make_uint4(dif.l[4], dif.l[5], dif.l[6], dif.l[7]);
        }
        __syncthreads();
    }

    // store: row u holds output K = bitrev_S(u)
    const uint32_t ns_mask = (1u << log_ns) - 1;
    for (uint32_t e = tid; e < TILE; e += NT) {
        uint32_t col = e % C, k = e / C;  // iterate K in natural order so stores walk forward
        uint32_t u = S ? bitrev_s<(S ? S : 1)>(k) : 0;
        uint32_t q = q0 + col;
        uint32_t jp = q >> log_ns, p = q & ns_mask;
        uint32_t oidx = (jp << (log_ns + S)) + p + (k << log_ns);
        if (oidx >= io.n_out) continue;
        uint32_t se = u * C + col;
        uint4 a0 = s_lo[se], a1 = s_hi[se];
        Fe v;
        v.l[0] = a0.x; v.l[1] = a0.y; v.l[2] = a0.z; v.l[3] = a0.w;
        v.l[4] = a1.x; v.l[5] = a1.y; v.l[6] = a1.z; v.l[7] = a1.w;
        if (!last) {
            uint32_t ex = (jp * k) << log_ns;  // < N
            if (ex) v = Fr::mul(v, load_fe_ro(&W[ex]));
        }
        if (io.epi) v = Fr::mul(v, io.epi_c[oidx % 3]);
        store_fe(&out[oidx], v);
    }
}

// a[i] *= c[i % m]  (parallelize-style elementwise maps: divide_by_vanishing_poly, log_n = 0 scaling)
__global__ void fr_scale_cyclic_kernel(Fe *a, uint32_t n, const Fe *__restrict__ c, uint32_t m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fe(&a[i], Fr::mul(load_fe(&a[i]), load_fe_ro(&c[i % m])));
}

}  // namespace h2b
