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
// S butterfly levels in rounds of up to three levels held in registers (2, 4 or 8 rows
// per thread, a swizzled shared-memory exchange between rounds, block barrier only after
// the first), inner twiddles from per-level tables, the inter-pass twiddle from a cached
// table of powers of omega, and stores of C adjacent elements per output row.
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

// Inner twiddles of a radix-2^S pass, one table per tile level L, entries contiguous in the butterfly index so that the
// lanes of a warp read neighbouring entries:  TW[off(L) + j] = omega_R^(j << L),  j < R >> (L + 1),  off(L) = R - (R >> L)
// (R - 1 entries; omega_R = omega^(N / R)).  Built once per (omega, log_n, S) from the table of all powers.
__global__ void ntt_pass_twiddles_kernel(const Fe *__restrict__ W, uint32_t log_n, uint32_t S, Fe *__restrict__ TW) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, R = 1u << S;
    if (i >= R - 1) return;
    uint32_t L = 0;
    while (i >= R - (R >> (L + 1))) L++;
    const uint32_t j = i - (R - (R >> L));
    store_fe(&TW[i], load_fe_ro(&W[((size_t)j << L) << (log_n - S)]));
}

// Shared-memory slot of tile element i = row * C + col.  The two 16-byte halves of an element live in two planes;
// within a plane the low three index bits are XORed with the next three, which makes every access pattern of the
// rounds below conflict-free for 8 elements per thread (a quarter-warp touches eight distinct 16-byte bank groups
// whether its lanes walk consecutive rows or rows 2, 4 or 8 apart; checked by enumeration for every tile shape).
H2B_DI uint32_t tile_slot(uint32_t i) { return i ^ ((i >> 3) & 7u); }

// Out of line: the fused input / output scalings of the domain transforms sit on paths a plain best_fft never takes,
// and eight inlined copies of each would double the kernel's instruction footprint.
static __device__ __noinline__ Fe fr_mul_ni(Fe a, Fe b) { return Fr::mul(a, b); }

struct PassArgs {
    const Fe *in;
    Fe *out;
    const Fe *W;    // all powers of omega (inter-pass twiddles)
    const Fe *TW;   // inner twiddles of this pass, per level (ntt_pass_twiddles_kernel)
    uint32_t log_n, log_ns, last, q0, M;
    uint32_t warp_sync;  // bit r set: the exchange after round r stays inside each warp (__syncwarp suffices)
};

// 16-byte asynchronous global -> shared copies (LDGSTS): twiddles are fetched ahead of their use without holding
// registers; .ca keeps the lines in L1 for the other blocks of the SM, which read the same tables.
H2B_DI void cp_async16(void *smem_dst, const void *gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
H2B_DI void cp_async_fe(uint4 *s_lo, uint4 *s_hi, uint32_t slot, const Fe *src) {
    cp_async16(&s_lo[slot], src);
    cp_async16(&s_hi[slot], reinterpret_cast<const uint4 *>(src) + 1);
}
H2B_DI void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// ---- the two mechanisms BASELINE.json's north_star names, as measured alternatives (VAR below)
// TMA (cp.async.bulk, SASS UBLKCP) staging of the first round's tile: one bulk copy per tile row into a linear
// staging area, completion counted by an mbarrier.
H2B_DI void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
H2B_DI void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
H2B_DI void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
H2B_DI void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
// Warp-shuffle exchange between two rounds of 4 elements per thread: the 2-bit register index is transposed with the
// 2-bit lane field at bit b (element t of the lane with field m goes to register m of the lane with field t).
H2B_DI void shfl_exchange4(Fe (&a)[4], uint32_t b) {
    const uint32_t m = (threadIdx.x >> b) & 3u;
    Fe o[4];
#pragma unroll
    for (int t = 0; t < 4; t++) o[t] = a[t];
#pragma unroll
    for (int d = 1; d < 4; d++) {
        const uint32_t pick = m ^ (uint32_t)d;  // the register this lane sends, and the one it receives into
        Fe send;
#pragma unroll
        for (int i = 0; i < 8; i++)
            send.l[i] = pick == 0 ? a[0].l[i] : (pick == 1 ? a[1].l[i] : (pick == 2 ? a[2].l[i] : a[3].l[i]));
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t r = __shfl_xor_sync(0xffffffffu, send.l[i], d << b);
#pragma unroll
            for (int t = 0; t < 4; t++)
                if (pick == (uint32_t)t) o[t].l[i] = r;
        }
    }
#pragma unroll
    for (int t = 0; t < 4; t++) a[t] = o[t];
}

// One pass.  S = log2 radix, C = columns per tile, EL = log2 elements per thread, NT = threads per block
// (NT * 2^EL >= 2^S * C; surplus threads only take part in the barriers), MINB = blocks per SM the register
// allocation is held to.
//
// A thread keeps E = 2^EL tile elements in registers and the S levels run in rounds of EL levels on them; between
// rounds the tile is exchanged through shared memory.  Round r works on the rows whose bits [s_r, s_r + EL) are the
// register index (s_0 = S - EL: the largest strides first, decimation in frequency); after the first exchange the
// sub-transforms of different warps are independent, so only that exchange needs a block barrier -- the later ones
// are __syncwarp (the host works out which, `warp_sync`).  The loop over rounds is NOT unrolled: there is one copy of
// the butterfly code whatever S is.  When EL does not divide S the last round is partial: it runs the last S mod EL
// levels of the same code.
//
// Values stay in [0, 2N) between butterflies (x + y is brought back with one conditional subtraction of 2N,
// x - y + 2N < 4N goes into the twiddle product as it is, and a Montgomery product of a value below 4N with a
// canonical twiddle is below 2N without its final subtraction: N < 2^254); intermediate passes store such values
// and only the last pass reduces to the canonical representative, so the results are bit-identical to the
// reference's.
//
// No multiplication waits for a global load.  The twiddles of the first round (E - 1 per thread, each used once per
// tile) are copied asynchronously into the thread's OWN tile slots, which are idle until the first exchange; those of
// the later rounds (2^(S - EL) - 1 values shared by the whole block) into a small table behind the tile; the
// inter-pass twiddles of the last round again into the thread's own slots, as soon as it has read them for the last
// time.  All of them are then read with shared-memory latency.
// VAR: 0 = the product path; 1 = first-round tile staged by TMA bulk copies (plain transforms only); 2 = warp-local
// exchanges by shuffles instead of shared memory (4 elements per thread, single-column tiles only).
template <int S, int C, int EL, int NT, int MINB, int VAR = 0>
__global__ void __launch_bounds__(NT, MINB)
ntt_pass_kernel(const Fe *in, Fe *out, const Fe *__restrict__ W, const Fe *__restrict__ TW, uint32_t log_n,
                uint32_t log_ns, uint32_t last, uint32_t warp_sync, NttIo io) {
    constexpr int R = 1 << S, E = 1 << EL, TILE = R * C;
    constexpr uint32_t GROUPS = TILE / E;
    constexpr uint32_t LATE = (R >> EL) - 1;  // twiddles of the levels >= EL
    extern __shared__ uint4 smem_u4[];
    uint4 *s_lo = smem_u4;                   // low 16 bytes of the tile elements
    uint4 *s_hi = smem_u4 + TILE;            // high 16 bytes
    uint4 *tw_late = smem_u4 + 2 * TILE;     // TW[R - (R >> EL) ..): low halves, then (R >> EL entries on) high halves
    uint4 *stage = tw_late + 2 * (R >> EL);  // VAR 1: the tile as TMA delivers it, [row][col] x 32 bytes, then the mbarrier
    uint64_t *mbar = reinterpret_cast<uint64_t *>(stage + 2 * TILE);

    const Fe *src = in + (size_t)blockIdx.y * io.bin;
    Fe *dst = out + (size_t)blockIdx.y * io.bout;
    const uint32_t M = 1u << (log_n - S);  // columns in the whole pass
    const uint32_t q0 = blockIdx.x * C;
    const uint32_t g = threadIdx.x;
    const bool active = GROUPS >= NT || g < GROUPS;
    const uint32_t col = g % C, rest = g / C;

    // later rounds' twiddles -> shared table (visible to the block after the first exchange's barrier)
    if constexpr (LATE > 0)
        for (uint32_t i = g; i < 2 * LATE; i += NT)
            cp_async16(&tw_late[(i >> 1) + (i & 1u) * (R >> EL)], reinterpret_cast<const uint4 *>(TW + (R - (R >> EL))) + i);

    if constexpr (VAR == 1) {
        if (g == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (g == 0) mbar_expect_tx(mbar, (uint32_t)(TILE * sizeof(Fe)));
    }

    Fe a[E];
    uint32_t lvl = 0;
    bool in_regs = false;  // VAR 2: the next round's elements are already in the registers (shuffle exchange)
    constexpr int kRoundUnroll = EL <= 2 ? 16 : 1;
    // 8 elements per thread: ONE copy of the round body (12 inlined products per round; unrolled, the kernel outgrows
    // the instruction cache).  Fewer elements per thread: rounds unrolled (4 resp. 1 products per round), which makes
    // every stride, slot offset and trivial-twiddle test of a round a compile-time constant.
#pragma unroll kRoundUnroll
    for (uint32_t r = 0; lvl < S; r++) {
        const uint32_t er = min((uint32_t)EL, S - lvl);          // levels of this round
        const uint32_t s = er < EL ? 0u : S - lvl - EL;           // register field = row bits [s, s + EL)
        const uint32_t lo = rest & ((1u << s) - 1u), hi = rest >> s;
        const uint32_t row_base = (hi << (s + EL)) | lo;
        const bool final_round = lvl + er >= S;
        if (active) {
            if (r == 0) {
                // first-round twiddles -> own slots: level q needs T_q[j], j = (t' << s) | lo, t' < E >> (q + 1);
                // it goes to the slot of register number (E >> (q + 1)) - 1 + t'
#pragma unroll
                for (int q = 0; q < EL; q++) {
                    const int half = E >> (q + 1);
                    const Fe *T = TW + (R - (R >> q));
#pragma unroll
                    for (int tp = 0; tp < half; tp++) {
                        const uint32_t j = ((uint32_t)tp << s) | lo;
                        const uint32_t e = tile_slot((row_base + ((uint32_t)(half - 1 + tp) << s)) * C + col);
                        cp_async_fe(s_lo, s_hi, e, &T[j]);
                    }
                }
                if constexpr (VAR == 1) {
                    // one bulk copy per tile row (C x 32 bytes), issued by the thread that holds the row's first column
                    if (col == 0) {
#pragma unroll
                        for (int t = 0; t < E; t++) {
                            const uint32_t row = row_base + ((uint32_t)t << s);
                            bulk_g2s(stage + 2 * row * C, &src[q0 + (size_t)row * M], (uint32_t)(C * sizeof(Fe)), mbar);
                        }
                    }
                    mbar_wait(mbar, 0);
#pragma unroll
                    for (int t = 0; t < E; t++) {
                        const uint32_t e = (row_base + ((uint32_t)t << s)) * C + col;
                        a[t] = fe_from_u4(stage[2 * e], stage[2 * e + 1]);
                    }
                }
                // first round: straight from global memory (128-bit loads of C adjacent elements per row), with the
                // fused input scaling (coset powers) and zero padding
#pragma unroll
                for (int t = 0; t < E && VAR != 1; t++) {
                    const uint32_t idx = q0 + col + (row_base + ((uint32_t)t << s)) * M;
                    Fe v;
                    if (idx < io.n_in) {
                        v = load_fe(&src[idx]);
                        if (io.pro) {
                            const uint32_t m3 = idx % 3;
                            if (m3) v = fr_mul_ni(v, io.pro_c[m3]);
                        }
                    } else {
                        v = Fr::zero();
                    }
                    a[t] = v;
                }
                cp_async_wait_all();
            } else {
                if (!in_regs) {
#pragma unroll
                    for (int t = 0; t < E; t++) {
                        const uint32_t e = tile_slot((row_base + ((uint32_t)t << s)) * C + col);
                        a[t] = fe_from_u4(s_lo[e], s_hi[e]);
                    }
                }
                if (final_round && !last) {
                    // inter-pass twiddles omega^(jp * K << log_ns) of this thread's outputs -> its own slots, which it
                    // has just read for the last time; they arrive while the last butterflies run
                    const uint32_t q = q0 + col, jp = q >> log_ns;
#pragma unroll
                    for (int t = 0; t < E; t++) {
                        const uint32_t k = bitrev_s<S>(row_base + (uint32_t)t);
                        const uint32_t ex = (jp * k) << log_ns;  // < N
                        cp_async_fe(s_lo, s_hi, tile_slot((row_base + (uint32_t)t) * C + col), &W[ex]);
                    }
                }
            }
            // the last `er` of the EL static levels: static level q pairs registers t and t + (E >> (q + 1))
#pragma unroll
            for (int q = 0; q < EL; q++) {
                if ((uint32_t)q + er >= (uint32_t)EL) {  // uniform
                    const int half = E >> (q + 1);
                    const uint32_t L = lvl + (uint32_t)q - ((uint32_t)EL - er);  // tile level of this static level
                    // twiddle source: own slots in the first round, the shared table afterwards
                    const uint4 *T = tw_late + ((R >> EL) - (R >> L));
#pragma unroll
                    for (int t = 0; t < E; t++) {
                        if (t & half) continue;
                        const Fe x = a[t], y = a[t + half];
                        a[t] = Fr::add_2n(x, y);
                        Fe d = Fr::sub_2n(x, y);  // in (0, 4N)
                        const uint32_t tp = (uint32_t)(t & (half - 1));
                        const uint32_t j = (tp << s) | lo;  // butterfly index mod its half-span
                        if (j != 0) {
                            Fe w;
                            if (r == 0) {
                                const uint32_t e = tile_slot((row_base + (((uint32_t)(half - 1) + tp) << s)) * C + col);
                                w = fe_from_u4(s_lo[e], s_hi[e]);
                            } else {
                                w = fe_from_u4(T[j], T[j + (R >> EL)]);
                            }
                            d = Fr::mul_lazy(d, w);  // back in [0, 2N)
                        } else {
                            d = Fr::reduce_2n(d);
                        }
                        a[t + half] = d;
                    }
                }
            }
        }
        if (!active && r == 0) cp_async_wait_all();  // its share of the shared twiddle table
        lvl += er;
        if constexpr (VAR == 2 && EL == 2 && C == 1) {
            // exchange by shuffles when both rounds are full and the lane field lies inside the warp
            in_regs = false;
            if (lvl < S && er == EL && S - lvl >= (uint32_t)EL && ((warp_sync >> r) & 1u) && r != 0) {
                const uint32_t s_next = S - lvl - EL;  // the thread bits [s_next, s_next + EL) swap with the register index
                if (s_next + EL <= 5) {
                    shfl_exchange4(a, s_next);
                    in_regs = true;
                    continue;
                }
            }
        }
        if (lvl < S) {
            if (active) {
#pragma unroll
                for (int t = 0; t < E; t++) {
                    const uint32_t e = tile_slot((row_base + ((uint32_t)t << s)) * C + col);
                    s_lo[e] = make_uint4(a[t].l[0], a[t].l[1], a[t].l[2], a[t].l[3]);
                    s_hi[e] = make_uint4(a[t].l[4], a[t].l[5], a[t].l[6], a[t].l[7]);
                }
            }
            // (the shared twiddle table is published by the first barrier: multi-warp blocks never skip it)
            if (((warp_sync >> r) & 1u) && (r != 0 || NT == 32)) __syncwarp();
            else __syncthreads();
        } else if (active) {
            // last round (s == 0): the registers are rows row_base .. row_base + E - 1; row u holds output K = bitrev_S(u)
            const uint32_t q = q0 + col;
            const uint32_t jp = q >> log_ns, p = q & ((1u << log_ns) - 1);
            if (!last && r != 0) cp_async_wait_all();
#pragma unroll
            for (int t = 0; t < E; t++) {
                const uint32_t u = row_base + (uint32_t)t;
                const uint32_t k = bitrev_s<S>(u);
                const uint32_t oidx = (jp << (log_ns + S)) + p + (k << log_ns);
                if (oidx >= io.n_out) continue;
                Fe v = a[t];
                if (!last) {
                    const uint32_t ex = (jp * k) << log_ns;  // < N
                    if (ex) {
                        Fe w;
                        if (r != 0) {
                            const uint32_t e = tile_slot(u * C + col);
                            w = fe_from_u4(s_lo[e], s_hi[e]);
                        } else {
                            w = load_fe_ro(&W[ex]);  // single-round tile: nothing to hide the load behind
                        }
                        v = Fr::mul_lazy(v, w);
                    }
                }
                // intermediate passes hand their values on in [0, 2N); the last pass stores canonical ones
                if (io.epi) v = fr_mul_ni(v, io.epi_c[oidx % 3]);
                else if (last) v = Fr::reduce_once(v);
                store_fe(&dst[oidx], v);
            }
        }
    }
}

// a[i] *= c[i % m]  (parallelize-style elementwise maps: divide_by_vanishing_poly, log_n = 0 scaling)
__global__ void fr_scale_cyclic_kernel(Fe *a, uint32_t n, const Fe *__restrict__ c, uint32_t m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_fe(&a[i], Fr::mul(load_fe(&a[i]), load_fe_ro(&c[i % m])));
}

}  // namespace h2b
