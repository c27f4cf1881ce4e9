// msm_comb.cuh -- commits against a SMALL registered SRS without buckets.
//
// The reference's own circuits prove at k = 4 ... 14 (arithmetic k = 4, Poseidon k = 7, Collatz
// k = 10).  At those sizes a Pippenger MSM is a latency chain, not a throughput problem: the bucket reduction
// is 20-40 SERIAL group additions of ~6 us each (a lone warp needs ~840 cycles per Montgomery product), and
// sorting, fix-up and reduction cost 4-5x the accumulation itself.  With the bases static and HBM at 180 GB
// the buckets can be removed altogether: for every point, window and digit magnitude the multiple
//      comb[(w * M + (d - 1)) * n + i] = d * 2^(c*w) * P_i          d = 1 .. M = 2^(c-1)
// is precomputed once (c = 8: 32 windows x 128 multiples = 256 KiB per point, 4 GiB at n = 2^14), and a commit
// is the plain SUM of n * W table entries: digits -> entry indices (no sort, no histogram), slice sums with
// mixed additions, a binary tree over the slice sums.  Depth: L + log2(n * W / L) additions.
#pragma once
#include "msm.cuh"

namespace h2b {

// The table is built in two launches.  msm_precompute_kernel (msm.cuh, wstride = M * n) fills the d = 1 column,
// 2^(c*w) * P_i, with one doubling chain per point.  Then one thread per (point, window) adds the base to itself
// M - 1 times in XYZZ coordinates and normalises ALL its multiples with one inversion: in a chain of mixed
// additions of the same base  ZZ_(d+1) = ZZ_d * P_d^2,  ZZZ_(d+1) = ZZZ_d * P_d^3  with  P_d = x_B * ZZ_d - X_d,  so
// ZZ_d = T_d^2 and ZZZ_d = T_d^3 for the running product T_d of the P_j, and 1 / T_d = (1 / T_(d+1)) * P_d: the raw
// X, Y of a multiple are parked in its table slot, the P_d in a scratch array, and a backward sweep from the single
// inverse 1 / T_M rewrites every slot as x = X / T^2, y = Y / T^3.  16 products per entry instead of ~395
// (one Fermat inversion each).  Affine coordinates are unique, so the table is bit-identical to the former one.
//   comb   : the table, column d = 1 already filled
//   scratch: (M - 1) * n * W field elements, [d][t]
static __device__ __noinline__ Fe xyzz_madd_keep_p(XYZZ &acc, const Affine &p) {
    const Fe u2 = Fq::mul(p.x, acc.zz);
    const Fe s2 = Fq::mul(p.y, acc.zzz);
    const Fe pp_ = Fq::sub(u2, acc.x);
    const Fe rr = Fq::sub(s2, acc.y);
    const Fe pp = Fq::sqr(pp_);
    const Fe ppp = Fq::mul(pp_, pp);
    const Fe q = Fq::mul(acc.x, pp);
    const Fe x3 = Fq::sub(Fq::sub(Fq::sub(Fq::sqr(rr), ppp), q), q);
    const Fe y3 = Fq::mul2_sub(rr, Fq::sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = Fq::mul(acc.zz, pp);
    acc.zzz = Fq::mul(acc.zzz, ppp);
    return pp_;
}
__global__ void __launch_bounds__(128)
msm_comb_multiples_kernel(uint32_t n, uint32_t c, uint32_t W, Affine *comb, Fe *__restrict__ scratch) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, threads = n * W;
    if (t >= threads) return;
    const uint32_t i = t % n, w = t / n, M = 1u << (c - 1);
    Affine *dst = comb + (size_t)w * M * n + i;
    Affine base;
    base.x = load_fe(&dst[0].x);
    base.y = load_fe(&dst[0].y);
    if (affine_is_identity(base)) {
        for (uint32_t d = 1; d < M; d++) {
            store_fe(&dst[(size_t)d * n].x, Fq::zero());
            store_fe(&dst[(size_t)d * n].y, Fq::zero());
        }
        return;
    }
    if (M < 2) return;
    // 2 * base: T_2 = P_1 = 2 y (dbl-2008-s-1 from an affine point: ZZ = (2y)^2, ZZZ = (2y)^3)
    XYZZ acc = xyzz_dbl_ni(xyzz_from_affine(base));
    Fe run = Fq::dbl(base.y);  // T_(d+1) after step d
    store_fe(&scratch[t], run);
    store_fe(&dst[n].x, acc.x);
    store_fe(&dst[n].y, acc.y);
#pragma unroll 1
    for (uint32_t d = 2; d < M; d++) {  // slot d holds (d + 1) * base; d * base + base is never a doubling or a
        const Fe p = xyzz_madd_keep_p(acc, base);  // cancellation: the group order is a 254-bit prime
        store_fe(&scratch[(size_t)(d - 1) * threads + t], p);
        run = Fq::mul(run, p);
        store_fe(&dst[(size_t)d * n].x, acc.x);
        store_fe(&dst[(size_t)d * n].y, acc.y);
    }
    Fe inv = Fq::inv(run);  // 1 / T_M
#pragma unroll 1
    for (uint32_t d = M - 1; d >= 1; d--) {
        const Fe i2 = Fq::sqr(inv);
        Affine *slot = &dst[(size_t)d * n];
        const Fe x = Fq::mul(load_fe(&slot->x), i2);
        const Fe y = Fq::mul(Fq::mul(load_fe(&slot->y), i2), inv);
        store_fe(&slot->x, x);
        store_fe(&slot->y, y);
        inv = Fq::mul(inv, load_fe(&scratch[(size_t)(d - 1) * threads + t]));  // 1 / T_d
    }
}

// Scalars -> table indices.  entries[(col * W + w) * n + i] = index | sign << 31; a zero digit points at the
// identity entry that closes the table (index W * M * stride).
__global__ void __launch_bounds__(256)
msm_comb_digits_kernel(const Fe *__restrict__ scalars, MsmCfg cfg, uint32_t *__restrict__ entries) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cfg.n * cfg.cols) return;
    const uint32_t col = t / cfg.n, i = t - col * cfg.n;
    Fe s = Fr::from_mont(load_fe_ro(&scalars[t]));
    uint32_t l[9];
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]), "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7])
        : "r"(s.l[0]), "r"(s.l[1]), "r"(s.l[2]), "r"(s.l[3]), "r"(s.l[4]), "r"(s.l[5]), "r"(s.l[6]),
          "r"(s.l[7]), "r"(cfg.half[0]), "r"(cfg.half[1]), "r"(cfg.half[2]), "r"(cfg.half[3]),
          "r"(cfg.half[4]), "r"(cfg.half[5]), "r"(cfg.half[6]), "r"(cfg.half[7]));
    l[8] = 0;
    const uint32_t ident = cfg.windows * cfg.bpw * cfg.stride;
    uint32_t *dst = entries + (size_t)col * cfg.windows * cfg.n + i;
    for (uint32_t w = 0; w < cfg.windows; w++) {
        const int32_t d = digit_at(l, w, cfg);
        uint32_t e = ident;
        if (d != 0) {
            const uint32_t neg = d < 0;
            const uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
            e = ((w * cfg.bpw + mag - 1) * cfg.stride + i) | (neg << 31);
        }
        dst[(size_t)w * cfg.n] = e;
    }
}

// Slice sums: thread (col, s) adds entries [s * L, (s + 1) * L) of its column.
__global__ void __launch_bounds__(128)
msm_comb_sum_kernel(const Affine *__restrict__ comb, const uint32_t *__restrict__ entries, uint32_t per_col, uint32_t L,
                    uint32_t slices, XYZZ *__restrict__ partial) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x, col = blockIdx.y;
    if (s >= slices) return;
    const uint32_t *e = entries + (size_t)col * per_col;
    const uint32_t begin = s * L, end = min(begin + L, per_col);
    XYZZ acc = xyzz_identity();
    uint32_t v = e[begin];
    Affine p = load_affine(&comb[v & 0x7fffffffu]);
    for (uint32_t k = begin; k < end; k++) {
        uint32_t vn = 0;
        Affine pn;
        const bool more = k + 1 < end;
        if (more) {
            vn = e[k + 1];
            pn = load_affine(&comb[vn & 0x7fffffffu]);
        }
        if (!affine_is_identity(p)) {
            if (v >> 31) p.y = Fq::neg(p.y);
            xyzz_madd(acc, p);
        }
        v = vn;
        p = pn;
    }
    store_xyzz(&partial[(size_t)col * slices + s], acc);
}

// A group addition shared by a team of 4 adjacent lanes.  A lone warp needs ~840 cycles per Montgomery product
// (its carry chains serialise), so a 14-product addition is ~6 us for one thread; the tree below has few
// active nodes and idle lanes, so four lanes take one product each per round and exchange the results by
// shuffles: 4 rounds instead of 14 products (add-2008-s: {U1,U2,S1,S2}, {PP,RR,ZZ1*ZZ2,ZZZ1*ZZZ2},
// {PPP,Q,ZZ3}, {R(Q-X3),S1*PPP,ZZZ3}).  a and b are replicated in the 4 lanes, and so is the result.  The
// shuffles run unconditionally (mask = the lanes of all teams doing an addition in this step); identities
// and equal / opposite x-coordinates are patched afterwards without communication.
H2B_DI Fe fe_team(const Fe &v, int src, uint32_t mask) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(mask, v.l[i], src, 4);
    return r;
}
H2B_DI XYZZ xyzz_add_team4(const XYZZ &a, const XYZZ &b, uint32_t lane4, uint32_t mask) {
    const bool l0 = lane4 == 0, l1 = lane4 == 1, l2 = lane4 == 2;
    // round 1
    Fe r = Fq::mul(fe_sel(l0, a.x, fe_sel(l1, b.x, fe_sel(l2, a.y, b.y))),
                   fe_sel(l0, b.zz, fe_sel(l1, a.zz, fe_sel(l2, b.zzz, a.zzz))));
    const Fe u1 = fe_team(r, 0, mask), u2 = fe_team(r, 1, mask), s1 = fe_team(r, 2, mask), s2 = fe_team(r, 3, mask);
    const Fe p = Fq::sub(u2, u1), rr_ = Fq::sub(s2, s1);
    // round 2
    r = Fq::mul(fe_sel(l0, p, fe_sel(l1, rr_, fe_sel(l2, a.zz, a.zzz))), fe_sel(l0, p, fe_sel(l1, rr_, fe_sel(l2, b.zz, b.zzz))));
    const Fe pp = fe_team(r, 0, mask), r2 = fe_team(r, 1, mask), zm = fe_team(r, 2, mask), zn = fe_team(r, 3, mask);
    // round 3
    r = Fq::mul(fe_sel(l0, p, fe_sel(l1, u1, zm)), pp);
    const Fe ppp = fe_team(r, 0, mask), q = fe_team(r, 1, mask), zz3 = fe_team(r, 2, mask);
    const Fe x3 = Fq::sub(Fq::sub(Fq::sub(r2, ppp), q), q);
    // round 4
    r = Fq::mul(fe_sel(l0, rr_, fe_sel(l1, s1, zn)), fe_sel(l0, Fq::sub(q, x3), ppp));
    const Fe t0 = fe_team(r, 0, mask), t1 = fe_team(r, 1, mask), zzz3 = fe_team(r, 2, mask);
    XYZZ out;
    out.x = x3;
    out.y = Fq::sub(t0, t1);
    out.zz = zz3;
    out.zzz = zzz3;
    // patches (uniform inside a team, no communication)
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    if (Fq::is_zero(p)) {  // same x: doubling or P + (-P)
        XYZZ t = a;
        xyzz_add(t, b);
        return t;
    }
    return out;
}

// One level group of the tree: block b of column `col` folds in[col][b * span, (b + 1) * span) into
// out[col][b]  (span = blockDim.x * per_thread; per_thread sequential additions, then a shared-memory tree).
__global__ void __launch_bounds__(256)
msm_comb_tree_kernel(const XYZZ *__restrict__ in, uint32_t count, uint32_t per_thread, XYZZ *__restrict__ out) {
    extern __shared__ uint4 tree_smem[];
    XYZZ *sh = reinterpret_cast<XYZZ *>(tree_smem);
    const uint32_t col = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    const XYZZ *src = in + (size_t)col * count;
    const uint32_t base = (blockIdx.x * nt + tid) * per_thread;
    XYZZ acc = xyzz_identity();
    for (uint32_t k = 0; k < per_thread; k++) {
        if (base + k < count) {
            XYZZ q = load_xyzz(&src[base + k]);
            xyzz_add(acc, q);
        }
    }
    store_xyzz(&sh[tid], acc);
    __syncthreads();
    for (uint32_t stride = nt >> 1; stride > 0; stride >>= 1) {
        if (stride * 4 <= nt) {
            // enough idle lanes: node t of this level is added by the team of lanes 4t .. 4t+3
            const uint32_t team = tid >> 2;
            const bool active = team < stride;
            const uint32_t mask = __ballot_sync(0xffffffffu, active);
            if (active) {
                const XYZZ a = load_xyzz(&sh[team]);
                const XYZZ b = load_xyzz(&sh[team + stride]);
                const XYZZ c = xyzz_add_team4(a, b, tid & 3, mask);
                __syncwarp(mask);
                if ((tid & 3) == 0) store_xyzz(&sh[team], c);
            }
        } else if (tid < stride) {
            XYZZ a = load_xyzz(&sh[tid]);
            XYZZ b = load_xyzz(&sh[tid + stride]);
            xyzz_add(a, b);
            store_xyzz(&sh[tid], a);
        }
        __syncthreads();
    }
    if (tid == 0) store_xyzz(&out[(size_t)col * gridDim.x + blockIdx.x], load_xyzz(&sh[0]));
}

}  // namespace h2b
