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
// split into at most four passes; each pass works on a [2^S rows] x [C columns] tile:
// 128-bit coalesced loads of C adjacent elements per row straight into registers,
// S butterfly levels in rounds of three levels held in registers (8 rows per thread,
// shared memory only between rounds), the inter-pass twiddle from a cached table of
// powers of omega, and stores of C adjacent elements per output row.
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
    // batch: gridDim.y independent transforms; transform b reads in + b * bin, writes out + b * bout (elements)
    uint32_t bin, bout;
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

H2B_DI Fe fe_from_u4(const uint4 &a, const uint4 &b) {
    Fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// RB decimation-in-frequency levels on 2^RB rows held in registers.  The rows are
// u_t = row_base + t * st of the tile (row_base mod st = lo), the levels are the tile levels
// lvl .. lvl + RB - 1; the butterfly (u, u + h) of tile level L uses omega_R^((u mod h) << L).
template <int RB>
H2B_DI void dif_levels(Fe (&a)[1 << RB], uint32_t lo, uint32_t st, uint32_t lvl,
                       const uint4 *__restrict__ t_lo, const uint4 *__restrict__ t_hi) {
#pragma unroll
    for (int q = 0; q < RB; q++) {
        const int half = 1 << (RB - 1 - q);
#pragma unroll
        for (int t = 0; t < (1 << RB); t++) {
            if (t & half) continue;
            const Fe x = a[t], y = a[t + half];
            a[t] = Fr::add(x, y);
            Fe d = Fr::sub(x, y);
            const uint32_t e = (lo + (uint32_t)(t & (half - 1)) * st) << (lvl + q);
            if (e != 0) d = Fr::mul(d, fe_from_u4(t_lo[e], t_hi[e]));
            a[t + half] = d;
        }
    }
}

struct PassArgs {
    const Fe *in;
    Fe *out;
    const Fe *W;
    uint32_t log_n, log_ns, last, q0, M;
};

// Rounds of a pass: 3 levels per round while possible (2 + 2 when four remain), the first round
// reads global memory, the last one writes it, intermediate results live in shared memory.
template <int S, int C, int NT, int LVL>
H2B_DI void pass_rounds(const PassArgs &pa, const NttIo &io, uint4 *s_lo, uint4 *s_hi,
                        const uint4 *t_lo, const uint4 *t_hi, uint32_t tid) {
    constexpr int REM = S - LVL;
    constexpr int RB = (REM >= 3 && REM != 4) ? 3 : (REM >= 2 ? 2 : 1);
    constexpr bool FIRST = LVL == 0, LAST = LVL + RB == S;
    constexpr uint32_t ST = (1u << S) >> (LVL + RB);
    constexpr uint32_t GROUPS = ((1u << S) * C) >> RB;
    for (uint32_t g = tid; g < GROUPS; g += NT) {
        const uint32_t col = g % C, rest = g / C;
        const uint32_t lo = rest % ST, hi = rest / ST;
        const uint32_t row_base = hi * (ST << RB) + lo;
        Fe a[1 << RB];
        if (FIRST) {
#pragma unroll
            for (int t = 0; t < (1 << RB); t++) {
                const uint32_t idx = pa.q0 + col + (row_base + t * ST) * pa.M;
                Fe v;
                if (idx < io.n_in) {
                    v = load_fe(&pa.in[idx]);
                    if (io.pro) {
                        const uint32_t m3 = idx % 3;
                        if (m3) v = Fr::mul(v, io.pro_c[m3]);
                    }
                } else {
                    v = Fr::zero();
                }
                a[t] = v;
            }
        } else {
#pragma unroll
            for (int t = 0; t < (1 << RB); t++) {
                const uint32_t e = (row_base + t * ST) * C + col;
                a[t] = fe_from_u4(s_lo[e], s_hi[e]);
            }
        }
        dif_levels<RB>(a, lo, ST, LVL, t_lo, t_hi);
        if (LAST) {
            // ST == 1: the rows are row_base .. row_base + 2^RB - 1; row u holds output K = bitrev_S(u)
            const uint32_t q = pa.q0 + col;
            const uint32_t jp = q >> pa.log_ns, p = q & ((1u << pa.log_ns) - 1);
#pragma unroll
            for (int t = 0; t < (1 << RB); t++) {
                const uint32_t u = row_base + t * ST;
                const uint32_t k = bitrev_s<S>(u);
                const uint32_t oidx = (jp << (pa.log_ns + S)) + p + (k << pa.log_ns);
                if (oidx >= io.n_out) continue;
                Fe v = a[t];
                if (!pa.last) {
                    const uint32_t ex = (jp * k) << pa.log_ns;  // < N
                    if (ex) v = Fr::mul(v, load_fe_ro(&pa.W[ex]));
                }
                if (io.epi) v = Fr::mul(v, io.epi_c[oidx % 3]);
                store_fe(&pa.out[oidx], v);
            }
        } else {
#pragma unroll
            for (int t = 0; t < (1 << RB); t++) {
                const uint32_t e = (row_base + t * ST) * C + col;
                s_lo[e] = make_uint4(a[t].l[0], a[t].l[1], a[t].l[2], a[t].l[3]);
                s_hi[e] = make_uint4(a[t].l[4], a[t].l[5], a[t].l[6], a[t].l[7]);
            }
        }
    }
    if constexpr (!LAST) {
        __syncthreads();
        pass_rounds<S, C, NT, LVL + RB>(pa, io, s_lo, s_hi, t_lo, t_hi, tid);
    }
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

    PassArgs pa;
    pa.in = in + (size_t)blockIdx.y * io.bin;
    pa.out = out + (size_t)blockIdx.y * io.bout;
    pa.W = W;
    pa.log_n = log_n; pa.log_ns = log_ns; pa.last = last;
    pa.M = 1u << (log_n - S);           // columns in the whole pass
    pa.q0 = blockIdx.x * C;
    const uint32_t tid = threadIdx.x;

    // inner twiddles: W[t * M]
    for (uint32_t t = tid; t < R / 2; t += NT) {
        const uint4 *p = reinterpret_cast<const uint4 *>(&W[(size_t)t * pa.M]);
        t_lo[t] = __ldg(p);
        t_hi[t] = __ldg(p + 1);
    }
    __syncthreads();
    pass_rounds<S, C, NT, 0>(pa, io, s_lo, s_hi, t_lo, t_hi, tid);
}

// a[i] *= c[i % m]  (parallelize-style elementwise maps: divide_by_vanishing_poly, log_n = 0 scaling)
__global__ void fr_scale_cyclic_kernel(Fe *a, uint32_t n, const Fe *__restrict__ c, uint32_t m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fe(&a[i], Fr::mul(load_fe(&a[i]), load_fe_ro(&c[i % m])));
}

}  // namespace h2b
