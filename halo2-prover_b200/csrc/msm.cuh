// msm.cuh -- Pippenger multi-scalar multiplication over BN254 G1 for sm_100a.
//
// Device replacement for halo2_proofs @6b43b6b src/arithmetic.rs:28-140
// (multiexp_serial) and :147-180 (best_multiexp), reached in the reference from
// ParamsKZG::commit / commit_lagrange (src/poly/kzg/commitment.rs:319, :363) under
// create_proof (/root/reference/circuits/src/utils.rs:83-91, :105-120).
//
// Contract kept: sum_i coeffs[i] * bases[i] as a group element (callers normalise
// to affine before the transcript, so the Projective representative is free);
// Montgomery-form inputs; (0,0) bases and zero scalars contribute nothing.
// The algorithm is NOT the reference's per-thread unsigned-window loop:
//   1. digits   : scalar -> canonical -> signed c-bit digits (carry-free recoding), stored
//                 window-major; per-(window,|digit|) bucket histogram with global atomics
//   2. scan     : exclusive scan of bucket sizes; compacted list of non-empty buckets
//   3. scatter  : window-major counting-sort of (point index, sign) into bucket order
//   4. accumulate: one thread per SLICE of L consecutive sorted entries, whatever buckets they
//                 belong to -- every lane performs the same number of XYZZ mixed additions
//                 (8M + 2S) over gathered affine bases, so skewed scalars and the short top
//                 window cost nothing extra; buckets wholly inside a slice are stored
//                 directly, pieces of buckets that straddle slices go to per-slice slots
//   5. fixup    : straddling buckets are folded from their pieces (thread per bucket, one
//                 block per bucket when a bucket spans many slices)
//   6. reduce   : per window, sum_k k * B_k by running sums over bucket groups and
//                 a shared-memory tree across groups
//   7. final    : Horner over windows (c doublings per window) -> projective
#pragma once
#include "curve.cuh"

namespace h2b {

struct MsmCfg {
    uint32_t n;        // points
    uint32_t c;        // window bits
    uint32_t windows;  // W
    uint32_t bpw;      // buckets per window = 2^(c-1)
    uint32_t nb;       // W * bpw
    uint32_t slice;    // L: sorted entries per accumulation thread
    uint32_t lgrp;     // log2 of buckets per reduction group
    uint32_t half[8];  // sum over windows w < W-1 of 2^(c*w + c-1): turns unsigned windows into signed digits
    // Precomputed-window mode (registered SRS): table[v * stride + i] = 2^(c*t*v) * P_i holds every t-th window power, so
    // window w = t*v + r feeds bucket set r (bucket = |digit| - 1) through table block v: t bucket sets, and a Horner
    // over t sums (c doublings each) at the end.  t = 1 (the default): ONE shared bucket set and no Horner pass; t > 1
    // trades 1/t of the table's HBM for t - 1 extra reductions.
    uint32_t shared;   // t: 0 = no table (one bucket set per window), >= 1 = bucket sets over a precomputed table
    uint32_t stride;   // points per window block of the table (the registered SRS length)
    uint32_t ioff;     // index of this chunk's first point inside the SRS
    // Batched commit (shared mode only): `cols` polynomials of n scalars each, laid out one after the
    // other, against the same bases; column q owns buckets [q * bpw, (q + 1) * bpw).
    // (with t bucket sets per column: [q * t * bpw, (q + 1) * t * bpw))
    uint32_t cols;     // >= 1
};

// Signed digits without a sequential carry: with s' = s + half (one 256-bit addition),
//   digit_w = ((s' >> c*w) mod 2^c) - 2^(c-1)   for w < W-1      in [-2^(c-1), 2^(c-1))
//   digit_top = s' >> c*(W-1)                                    in [0, 2^(c-1)]
// and sum_w digit_w * 2^(c*w) = s.  Bucket of a non-zero digit: w * 2^(c-1) + |digit| - 1.
H2B_DI int32_t digit_at(const uint32_t (&sp)[9], uint32_t w, const MsmCfg &cfg) {
    const uint32_t off = w * cfg.c;
    const uint32_t idx = off >> 5, sh = off & 31;
    const uint64_t two = ((uint64_t)sp[idx + 1] << 32) | sp[idx];
    const uint32_t raw = (uint32_t)(two >> sh);
    if (w + 1 == cfg.windows) return (int32_t)raw;
    return (int32_t)(raw & ((1u << cfg.c) - 1)) - (int32_t)(1u << (cfg.c - 1));
}

// Pass 1 over the scalars: Montgomery -> canonical, add `half`, and for every window write the
// encoded digit (0 = none, else bucket-in-window + 1, sign in bit 31) to digits[w * n + i]
// (coalesced per window) and count it in the bucket histogram.
__global__ void __launch_bounds__(256)
msm_digits_kernel(const Fe *__restrict__ scalars, MsmCfg cfg, uint32_t *__restrict__ counts,
                  uint32_t *__restrict__ digits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // column-major: i = column * n + point
    const uint32_t total = cfg.n * cfg.cols;
    if (i >= total) return;
    const uint32_t colbase = cfg.cols > 1 ? (i / cfg.n) * cfg.shared * cfg.bpw : 0u;
    Fe s = Fr::from_mont(load_fe_ro(&scalars[i]));
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
    for (uint32_t w = 0; w < cfg.windows; w++) {
        const int32_t d = digit_at(l, w, cfg);
        uint32_t enc = 0;
        if (d != 0) {
            const uint32_t neg = d < 0;
            const uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
            enc = mag | (neg << 31);
            atomicAdd(&counts[(cfg.shared ? colbase + (w % cfg.shared) * cfg.bpw : w * cfg.bpw) + mag - 1], 1u);
        }
        digits[(size_t)w * total + i] = enc;
    }
}

// Pass 2, window-major (blockIdx.y = window, so the blocks of one window run together and its
// ~n * 4 B region of `sorted` stays L2-resident while it is being filled): counting-sort scatter.
__global__ void __launch_bounds__(256)
msm_scatter_kernel(const uint32_t *__restrict__ digits, MsmCfg cfg, uint32_t *__restrict__ cursor,
                   uint32_t *__restrict__ sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t w = blockIdx.y;
    const uint32_t total = cfg.n * cfg.cols;
    if (i >= total) return;
    const uint32_t enc = __ldg(&digits[(size_t)w * total + i]);
    if (enc == 0) return;
    const uint32_t mag = enc & 0x7fffffffu;
    const uint32_t col = cfg.cols > 1 ? i / cfg.n : 0u;
    const uint32_t pos =
        atomicAdd(&cursor[(cfg.shared ? (col * cfg.shared + w % cfg.shared) * cfg.bpw : w * cfg.bpw) + mag - 1], 1u);
    const uint32_t idx = cfg.shared ? (w / cfg.shared) * cfg.stride + cfg.ioff + (i - col * cfg.n) : i;
    sorted[pos] = idx | (enc & 0x80000000u);
}

// Exclusive scans over the nb buckets in three small launches (block sums, scan of block sums,
// rescan with offsets):
//   cursor[b]  = sum_{b' < b} counts[b']                       (write position of the scatter)
//   ne_off[j], ne_id[j] = offset and id of the j-th NON-EMPTY bucket; ne_off[J] = total entries
//   totals[0] = total entries, totals[1] = J
// Blocks of 1024 threads, `ipt` consecutive buckets per thread, at most 1024 blocks.
__global__ void __launch_bounds__(1024)
msm_scan_sums_kernel(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t ipt,
                     uint2 *__restrict__ block_sums) {
    __shared__ uint32_t wa[32], wb[32];
    const uint32_t tid = threadIdx.x;
    const uint32_t lo = (blockIdx.x * 1024 + tid) * ipt;
    uint32_t sa = 0, sb = 0;
    for (uint32_t k = 0; k < ipt; k++) {
        uint32_t b = lo + k;
        if (b < nb) {
            uint32_t cnt = counts[b];
            sa += cnt;
            sb += cnt != 0;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        sa += __shfl_down_sync(0xffffffffu, sa, d);
        sb += __shfl_down_sync(0xffffffffu, sb, d);
    }
    if ((tid & 31) == 0) { wa[tid >> 5] = sa; wb[tid >> 5] = sb; }
    __syncthreads();
    if (tid < 32) {
        sa = wa[tid];
        sb = wb[tid];
        for (int d = 16; d > 0; d >>= 1) {
            sa += __shfl_down_sync(0xffffffffu, sa, d);
            sb += __shfl_down_sync(0xffffffffu, sb, d);
        }
        if (tid == 0) block_sums[blockIdx.x] = make_uint2(sa, sb);
    }
}
// One block: exclusive scan of the (<= 1024) block sums in place; totals out.
__global__ void __launch_bounds__(1024)
msm_scan_blocks_kernel(uint2 *__restrict__ block_sums, uint32_t nblocks, uint32_t *__restrict__ totals) {
    __shared__ uint32_t sh_a[1024], sh_b[1024];
    const uint32_t tid = threadIdx.x;
    uint2 v = tid < nblocks ? block_sums[tid] : make_uint2(0, 0);
    sh_a[tid] = v.x;
    sh_b[tid] = v.y;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t va = 0, vb = 0;
        if (tid >= d) { va = sh_a[tid - d]; vb = sh_b[tid - d]; }
        __syncthreads();
        sh_a[tid] += va;
        sh_b[tid] += vb;
        __syncthreads();
    }
    if (tid < nblocks) block_sums[tid] = make_uint2(sh_a[tid] - v.x, sh_b[tid] - v.y);
    if (tid == 1023) { totals[0] = sh_a[1023]; totals[1] = sh_b[1023]; }
}
__global__ void __launch_bounds__(1024)
msm_scan_apply_kernel(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t ipt,
                      const uint2 *__restrict__ block_sums, uint32_t *__restrict__ cursor,
                      uint32_t *__restrict__ ne_off, uint32_t *__restrict__ ne_id) {
    __shared__ uint32_t sh_a[1024], sh_b[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t lo = (blockIdx.x * 1024 + tid) * ipt;
    uint32_t sa = 0, sb = 0;
    for (uint32_t k = 0; k < ipt; k++) {
        uint32_t b = lo + k;
        if (b < nb) {
            uint32_t cnt = counts[b];
            sa += cnt;
            sb += cnt != 0;
        }
    }
    sh_a[tid] = sa;
    sh_b[tid] = sb;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t va = 0, vb = 0;
        if (tid >= d) { va = sh_a[tid - d]; vb = sh_b[tid - d]; }
        __syncthreads();
        sh_a[tid] += va;
        sh_b[tid] += vb;
        __syncthreads();
    }
    const uint2 base = block_sums[blockIdx.x];
    uint32_t ra = base.x + sh_a[tid] - sa, rb = base.y + sh_b[tid] - sb;
    for (uint32_t k = 0; k < ipt; k++) {
        uint32_t b = lo + k;
        if (b < nb) {
            uint32_t cnt = counts[b];
            cursor[b] = ra;
            if (cnt) {
                ne_off[rb] = ra;
                ne_id[rb] = b;
                rb++;
            }
            ra += cnt;
        }
    }
    // the thread that owns the last bucket closes the list: ne_off[J] = total
    if (lo < nb && lo + ipt >= nb) ne_off[rb] = ra;
}

// One thread per slice of cfg.slice consecutive sorted entries.
//   piece kinds: DIRECT (bucket begins and ends inside the slice) -> bucket_sums[id]
//                HEAD   (bucket began in an earlier slice)         -> head[s]
//                TAIL   (bucket begins here, continues past the slice end) -> tail[s], tail_j[s] = j
__global__ void __launch_bounds__(128, 4)  // 4 blocks per SM: the register budget is 128 (ncu: 3 blocks at 130 registers)
msm_accumulate_kernel(const Affine *__restrict__ bases, const uint32_t *__restrict__ sorted,
                      const uint32_t *__restrict__ ne_off, const uint32_t *__restrict__ ne_id,
                      const uint32_t *__restrict__ totals, MsmCfg cfg, XYZZ *__restrict__ bucket_sums,
                      XYZZ *__restrict__ head, XYZZ *__restrict__ tail, int32_t *__restrict__ tail_j,
                      int32_t *__restrict__ head_j) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = totals[0], J = totals[1];
    const uint64_t begin64 = (uint64_t)s * cfg.slice;
    if (begin64 >= total) return;
    const uint32_t begin = (uint32_t)begin64;
    const uint32_t end = min(begin + cfg.slice, total);
    uint32_t lo = 0, hi = J;  // largest j in [0, J) with ne_off[j] <= begin  (ne_off[0] == 0)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (ne_off[mid] <= begin) lo = mid; else hi = mid;
    }
    uint32_t j = lo;
    uint32_t bend = ne_off[j + 1];
    bool started_before = ne_off[j] < begin;
    if (started_before) head_j[s] = (int32_t)j;  // this slice holds a middle / last piece of bucket j

    XYZZ acc = xyzz_identity();
    uint32_t e = begin;
    uint32_t v = sorted[e];
    Affine p = load_affine(&bases[v & 0x7fffffffu]);
    while (true) {
        if (e == bend) {  // bucket j is complete
            if (started_before) store_xyzz(&head[s], acc);
            else store_xyzz(&bucket_sums[ne_id[j]], acc);
            acc = xyzz_identity();
            started_before = false;
            j++;
            bend = ne_off[j + 1];
        }
        uint32_t vn = 0;
        Affine pn;
        const bool more = e + 1 < end;
        if (more) {  // prefetch the next point while this one is added
            vn = sorted[e + 1];
            pn = load_affine(&bases[vn & 0x7fffffffu]);
        }
        if (!affine_is_identity(p)) {
            if (v >> 31) p.y = Fq::neg(p.y);
            xyzz_madd(acc, p);
        }
        if (!more) break;
        v = vn;
        p = pn;
        e++;
    }
    if (started_before) {
        store_xyzz(&head[s], acc);           // middle or last piece of a long bucket
    } else if (bend > end) {
        store_xyzz(&tail[s], acc);           // first piece of a bucket that continues
        tail_j[s] = (int32_t)j;
    } else {
        store_xyzz(&bucket_sums[ne_id[j]], acc);
    }
}

// Buckets that straddle slices.  The slice holding the first piece (tail) owns the bucket:
// sum = tail[s0] + head[s0+1] + ... + head[s0+last].  The pieces of one bucket are adjacent slice
// slots, so they are folded as a pairwise tree, one launch per level r (piece o absorbs piece
// o + 2^r when o is a multiple of 2^(r+1)): every thread performs at most one addition per role and
// level, whatever the bucket sizes, instead of one thread walking a whole bucket.  Five levels cover
// spans of up to 32 slices; longer ones (skewed scalars, short top window) are queued in level 0
// and folded by one block each (msm_fixup_heavy_kernel).  The last level publishes the sums.
constexpr uint32_t kFixupLevels = 5;
__global__ void __launch_bounds__(128)
msm_fixup_level_kernel(const uint32_t *__restrict__ ne_off, const uint32_t *__restrict__ ne_id,
                       const uint32_t *__restrict__ totals, MsmCfg cfg, XYZZ *__restrict__ head,
                       XYZZ *__restrict__ tail, const int32_t *__restrict__ tail_j,
                       const int32_t *__restrict__ head_j, XYZZ *__restrict__ bucket_sums,
                       uint32_t *__restrict__ heavy, uint32_t r) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = totals[0];
    if ((uint64_t)s * cfg.slice >= total) return;
    const uint32_t step = 1u << r;
    const int32_t jh = head_j[s];
    if (jh >= 0) {  // a non-first piece of bucket jh
        const uint32_t s0 = ne_off[jh] / cfg.slice;
        const uint32_t last = (ne_off[jh + 1] - 1) / cfg.slice - s0, o = s - s0;
        if (last < (1u << kFixupLevels) && (o & (2 * step - 1)) == 0 && o + step <= last) {
            XYZZ a = load_xyzz(&head[s]);
            XYZZ b = load_xyzz(&head[s + step]);
            xyzz_add(a, b);
            store_xyzz(&head[s], a);
        }
    }
    const int32_t jt = tail_j[s];
    if (jt >= 0) {  // the first piece: this slice owns bucket jt
        const uint32_t last = (ne_off[jt + 1] - 1) / cfg.slice - s;
        if (last >= (1u << kFixupLevels)) {
            if (r == 0) heavy[1 + atomicAdd(&heavy[0], 1u)] = s;
            return;
        }
        const bool fold = step <= last, publish = r + 1 == kFixupLevels;
        if (!fold && !publish) return;
        XYZZ a = load_xyzz(&tail[s]);
        if (fold) {
            XYZZ b = load_xyzz(&head[s + step]);
            xyzz_add(a, b);
            if (!publish) store_xyzz(&tail[s], a);
        }
        if (publish) store_xyzz(&bucket_sums[ne_id[jt]], a);
    }
}
__global__ void __launch_bounds__(128)
msm_fixup_heavy_kernel(const uint32_t *__restrict__ ne_off, const uint32_t *__restrict__ ne_id, MsmCfg cfg,
                       const XYZZ *__restrict__ head, const XYZZ *__restrict__ tail,
                       const int32_t *__restrict__ tail_j, const uint32_t *__restrict__ heavy,
                       XYZZ *__restrict__ bucket_sums) {
    __shared__ uint4 comb_smem[128 * 8];
    XYZZ *sh = reinterpret_cast<XYZZ *>(comb_smem);
    const uint32_t nheavy = heavy[0], tid = threadIdx.x;
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const uint32_t s = heavy[1 + h];
        const int32_t j = tail_j[s];
        const uint32_t s_last = (ne_off[j + 1] - 1) / cfg.slice;
        XYZZ acc = xyzz_identity();
        if (tid == 0) acc = load_xyzz(&tail[s]);
        for (uint32_t t = s + 1 + tid; t <= s_last; t += 128) {
            XYZZ q = load_xyzz(&head[t]);
            xyzz_add(acc, q);
        }
        store_xyzz(&sh[tid], acc);
        __syncthreads();
        for (uint32_t stride = 64; stride > 0; stride >>= 1) {
            if (tid < stride) {
                XYZZ a = load_xyzz(&sh[tid]);
                XYZZ b2 = load_xyzz(&sh[tid + stride]);
                xyzz_add(a, b2);
                store_xyzz(&sh[tid], a);
            }
            __syncthreads();
        }
        if (tid == 0) store_xyzz(&bucket_sums[ne_id[j]], load_xyzz(&sh[0]));
        __syncthreads();
    }
}

// dst[b] += src[b] for every bucket: merges the bucket sums of a later chunk into the running ones.
__global__ void __launch_bounds__(128)
msm_merge_kernel(XYZZ *__restrict__ dst, const XYZZ *__restrict__ src, uint32_t nb) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    XYZZ q = load_xyzz(&src[b]);
    if (xyzz_is_identity(q)) return;
    XYZZ acc = load_xyzz(&dst[b]);
    xyzz_add(acc, q);
    store_xyzz(&dst[b], acc);
}

// Per window: sum_{k=1..bpw} k * B_k.  Grid = (blocks per window, windows); thread g of a window
// owns the L = 2^lgrp buckets k in (g*L, (g+1)*L]: running sums give sum (k - g*L) * B_k, the group
// offset (g*L) * sum B_k is added by a short double-and-add, a shared-memory tree folds the block,
// and each block writes one partial (window_partials[w * gridDim.x + blockIdx.x]).
__global__ void __launch_bounds__(256)
msm_reduce_kernel(const XYZZ *__restrict__ bucket_sums, MsmCfg cfg, XYZZ *__restrict__ window_partials) {
    extern __shared__ uint4 red_smem[];
    XYZZ *sh = reinterpret_cast<XYZZ *>(red_smem);
    const uint32_t w = blockIdx.y, tid = threadIdx.x, G = blockDim.x;
    const uint32_t g = blockIdx.x * G + tid;  // group index inside the window
    const uint32_t L = 1u << cfg.lgrp;
    const XYZZ *bk = bucket_sums + (size_t)w * cfg.bpw + (size_t)g * L;
    XYZZ running = xyzz_identity(), acc = xyzz_identity();
    for (int k = (int)L - 1; k >= 0; k--) {
        XYZZ s = load_xyzz(&bk[k]);
        xyzz_add(running, s);
        xyzz_add(acc, running);
    }
    // acc = sum (k - g*L) * B_k ; add (g*L) * running = 2^lgrp * (g * running)
    if (g != 0 && !xyzz_is_identity(running)) {
        XYZZ t = xyzz_identity();
        for (int bit = 31 - __clz(g); bit >= 0; bit--) {
            t = xyzz_dbl_ni(t);
            if ((g >> bit) & 1) xyzz_add(t, running);
        }
        for (uint32_t d = 0; d < cfg.lgrp; d++) t = xyzz_dbl_ni(t);
        xyzz_add(acc, t);
    }
    store_xyzz(&sh[tid], acc);
    __syncthreads();
    for (uint32_t stride = G >> 1; stride > 0; stride >>= 1) {
        if (tid < stride) {
            XYZZ a = load_xyzz(&sh[tid]);
            XYZZ b2 = load_xyzz(&sh[tid + stride]);
            xyzz_add(a, b2);
            store_xyzz(&sh[tid], a);
        }
        __syncthreads();
    }
    if (tid == 0) store_xyzz(&window_partials[(size_t)w * gridDim.x + blockIdx.x], load_xyzz(&sh[0]));
}

// Fold the per-block partials of each window: one warp-sized block per window.
__global__ void __launch_bounds__(32)
msm_window_fold_kernel(const XYZZ *__restrict__ window_partials, uint32_t per_window,
                       XYZZ *__restrict__ window_sums) {
    __shared__ uint4 fold_smem[32 * 8];
    XYZZ *sh = reinterpret_cast<XYZZ *>(fold_smem);
    const uint32_t w = blockIdx.x, lane = threadIdx.x;
    XYZZ acc = xyzz_identity();
    for (uint32_t i = lane; i < per_window; i += 32) {
        XYZZ p = load_xyzz(&window_partials[(size_t)w * per_window + i]);
        xyzz_add(acc, p);
    }
    store_xyzz(&sh[lane], acc);
    __syncwarp();
    for (uint32_t stride = 16; stride > 0; stride >>= 1) {
        if (lane < stride) {
            XYZZ a = load_xyzz(&sh[lane]);
            XYZZ b2 = load_xyzz(&sh[lane + stride]);
            xyzz_add(a, b2);
            store_xyzz(&sh[lane], a);
        }
        __syncwarp();
    }
    if (lane == 0) store_xyzz(&window_sums[w], load_xyzz(&sh[0]));
}

// The c * W doublings of the final Horner are a pure latency chain (one modmul is ~900 cycles for a
// single thread because its carry chains serialise), and it is the floor of every MSM however small.
// Four lanes share one doubling: the 9 products of dbl-2008-s-1 are issued as 3 rounds of
// independent products, one per lane, and exchanged with warp shuffles.
H2B_DI Fe fe_bcast4(const Fe &v, int src) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xfu, v.l[i], src);
    return r;
}
H2B_DI Fe fe_sel(bool c, const Fe &a, const Fe &b) {
    Fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
// p is replicated in lanes 0..3 and is not the identity.
H2B_DI void xyzz_dbl_coop4(XYZZ &p, uint32_t lane) {
    const Fe u = Fq::dbl(p.y);
    // round 1: lane 0: V = U^2, lane 1: XX = X^2
    const Fe a1 = fe_sel(lane == 0, u, p.x);
    const Fe r1 = Fq::mul(a1, a1);
    const Fe v = fe_bcast4(r1, 0), xx = fe_bcast4(r1, 1);
    const Fe m = Fq::add(Fq::dbl(xx), xx);
    // round 2: lane 0: W = U*V, lane 1: S = X*V, lane 2: M^2, lane 3: ZZ' = V*ZZ
    const Fe a2 = fe_sel(lane == 0, u, fe_sel(lane == 1, p.x, fe_sel(lane == 2, m, v)));
    const Fe b2 = fe_sel(lane == 2, m, fe_sel(lane == 3, p.zz, v));
    const Fe r2 = Fq::mul(a2, b2);
    const Fe w = fe_bcast4(r2, 0), sx = fe_bcast4(r2, 1), mm = fe_bcast4(r2, 2), zz3 = fe_bcast4(r2, 3);
    const Fe x3 = Fq::sub(Fq::sub(mm, sx), sx);
    // round 3: lane 0: W*Y, lane 1: ZZZ' = W*ZZZ, lane 2: M*(S - X')
    const Fe a3 = fe_sel(lane == 2, m, w);
    const Fe b3 = fe_sel(lane == 0, p.y, fe_sel(lane == 1, p.zzz, Fq::sub(sx, x3)));
    const Fe r3 = Fq::mul(a3, b3);
    const Fe wy = fe_bcast4(r3, 0), zzz3 = fe_bcast4(r3, 1), msx = fe_bcast4(r3, 2);
    p.x = x3;
    p.y = Fq::sub(msx, wy);
    p.zz = zz3;
    p.zzz = zzz3;
}

// Horner over windows, high to low; result as a homogeneous projective point (96 B).
// Launched with one warp; lanes 0..3 cooperate, the accumulator is replicated in them.
__global__ void __launch_bounds__(32, 1)
msm_final_kernel(const XYZZ *__restrict__ window_sums, MsmCfg cfg, Projective *out) {
    const uint32_t lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 4) return;
    XYZZ acc = xyzz_identity();
#pragma unroll 1
    for (int w = (int)cfg.windows - 1; w >= 0; w--) {
        if (!xyzz_is_identity(acc)) {  // uniform across the 4 lanes
#pragma unroll 1
            for (uint32_t d = 0; d < cfg.c; d++) xyzz_dbl_coop4(acc, lane);
        }
        XYZZ s = load_xyzz(&window_sums[w]);
        xyzz_add(acc, s);
    }
    if (lane == 0) {
        Projective j = xyzz_to_projective(acc);
        store_fe(&out->x, j.x);
        store_fe(&out->y, j.y);
        store_fe(&out->z, j.z);
    }
}

// Batched commit: column q's result is the Horner of its t bucket-set sums (t = 1: the single sum, no doubling).
__global__ void __launch_bounds__(32)
msm_batch_out_kernel(const XYZZ *__restrict__ window_sums, uint32_t cols, uint32_t t, uint32_t c, Projective *__restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cols) return;
    XYZZ acc = load_xyzz(&window_sums[(size_t)q * t + (t - 1)]);
#pragma unroll 1
    for (int r = (int)t - 2; r >= 0; r--) {
#pragma unroll 1
        for (uint32_t d = 0; d < c; d++) acc = xyzz_dbl_ni(acc);
        XYZZ s = load_xyzz(&window_sums[(size_t)q * t + r]);
        xyzz_add(acc, s);
    }
    const Projective j = xyzz_to_projective(acc);
    store_fe(&out[q].x, j.x);
    store_fe(&out[q].y, j.y);
    store_fe(&out[q].z, j.z);
}

// table[w * wstride + i] = 2^(c*w) * bases[i] in affine form, w in [0, W): the one-time precomputation behind the
// shared-bucket mode of a registered SRS (wstride = n; the bases of ParamsKZG are static) and the first column of the
// bucket-free table of msm_comb.cuh (wstride = 2^(c-1) * n).
// One thread per point walks the whole doubling chain in Jacobian coordinates (2M + 5S per doubling) and normalises
// kPreGroup windows with ONE inversion (Montgomery's trick over the Z of the window boundaries): the raw X, Y
// of a window are parked in its table slot, Z and the prefix products live in local memory, and the backward
// sweep rewrites the slots in affine form.  (The per-window Fermat inversion this replaces was two thirds of
// the kernel: 380 products against 20 doublings.)  Affine coordinates are unique, so the table is bit-identical.
constexpr uint32_t kPreGroup = 16;
__global__ void __launch_bounds__(128)
msm_precompute_kernel(const Affine *__restrict__ bases, uint32_t n, uint32_t c, uint32_t W, size_t wstride,
                      Affine *table) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine p = load_affine(&bases[i]);
    store_fe(&table[i].x, p.x);
    store_fe(&table[i].y, p.y);
    if (affine_is_identity(p)) {
        for (uint32_t w = 1; w < W; w++) {
            store_fe(&table[(size_t)w * wstride + i].x, Fq::zero());
            store_fe(&table[(size_t)w * wstride + i].y, Fq::zero());
        }
        return;
    }
    Jac cur;
    cur.x = p.x;
    cur.y = p.y;
    cur.z = Fq::one();
#pragma unroll 1
    for (uint32_t w0 = 1; w0 < W; w0 += kPreGroup) {
        const uint32_t cnt = min(kPreGroup, W - w0);
        Fe zs[kPreGroup], pre[kPreGroup];
#pragma unroll 1
        for (uint32_t t = 0; t < cnt; t++) {
#pragma unroll 1
            for (uint32_t d = 0; d < c; d++) jac_dbl_ni(cur);
            Affine *slot = &table[(size_t)(w0 + t) * wstride + i];
            store_fe(&slot->x, cur.x);
            store_fe(&slot->y, cur.y);
            zs[t] = cur.z;  // never zero: the group has prime order, so no doubling reaches the identity
            pre[t] = t ? Fq::mul(pre[t - 1], cur.z) : cur.z;
        }
        Fe inv = Fq::inv(pre[cnt - 1]);
#pragma unroll 1
        for (int t = (int)cnt - 1; t >= 0; t--) {
            const Fe zi = t ? Fq::mul(inv, pre[t - 1]) : inv;  // 1 / Z_t
            if (t) inv = Fq::mul(inv, zs[t]);
            const Fe zi2 = Fq::sqr(zi);
            Affine *slot = &table[(size_t)(w0 + t) * wstride + i];
            const Fe x = Fq::mul(load_fe(&slot->x), zi2);
            const Fe y = Fq::mul(Fq::mul(load_fe(&slot->y), zi2), zi);
            store_fe(&slot->x, x);
            store_fe(&slot->y, y);
            if (t == (int)cnt - 1) {  // the chain continues from the normalised point
                cur.x = x;
                cur.y = y;
                cur.z = Fq::one();
            }
        }
    }
}

// out[i] = [scalars[i]] * base, affine ((0,0) for the identity): the per-element fixed-base
// multiplication of ParamsKZG::setup (halo2_proofs @6b43b6b src/poly/kzg/commitment.rs:68-114,
// `g_projective[i] = g * s^i` under parallelize, then batch_normalize).  Used to build
// synthetic SRS / benchmark bases on the device.
static __device__ __noinline__ void xyzz_madd_ni(XYZZ &acc, const Affine &p) { xyzz_madd(acc, p); }

__global__ void __launch_bounds__(128)
g1_fixed_base_mul_kernel(const Fe *__restrict__ scalars, uint32_t n, Affine base, Affine *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe k = Fr::from_mont(load_fe_ro(&scalars[i]));
    uint32_t kl[8];
#pragma unroll
    for (int j = 0; j < 8; j++) kl[j] = k.l[j];
    XYZZ acc = xyzz_identity();
    const bool base_is_id = affine_is_identity(base);
#pragma unroll 1
    for (int bit = 253; bit >= 0; bit--) {
        acc = xyzz_dbl_ni(acc);
        if (!base_is_id && ((kl[bit >> 5] >> (bit & 31)) & 1)) xyzz_madd_ni(acc, base);
    }
    Affine r;
    if (xyzz_is_identity(acc)) {
        r.x = Fq::zero();
        r.y = Fq::zero();
    } else {
        Fe t = Fq::inv(Fq::mul(acc.zz, acc.zzz));
        r.x = Fq::mul(Fq::mul(acc.x, t), acc.zzz);  // X / ZZ
        r.y = Fq::mul(Fq::mul(acc.y, t), acc.zz);   // Y / ZZZ
    }
    store_fe(&out[i].x, r.x);
    store_fe(&out[i].y, r.y);
}

// out = sum of `count` projective points (multi-GPU fold of partial MSM results).
__global__ void g1_fold_kernel(const Projective *__restrict__ pts, uint32_t count, Projective *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    XYZZ acc = xyzz_identity();
    for (uint32_t i = 0; i < count; i++) {
        Projective p;
        p.x = load_fe(&pts[i].x);
        p.y = load_fe(&pts[i].y);
        p.z = load_fe(&pts[i].z);
        XYZZ q = projective_to_xyzz(p);
        xyzz_add(acc, q);
    }
    Projective j = xyzz_to_projective(acc);
    store_fe(&out->x, j.x);
    store_fe(&out->y, j.y);
    store_fe(&out->z, j.z);
}

}  // namespace h2b
