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
//   1. digits   : scalar -> canonical -> signed c-bit digits; per-(window,|digit|)
//                 bucket histogram with global atomics
//   2. scan     : exclusive scan of bucket sizes, and of per-bucket task counts
//   3. scatter  : counting-sort of (point index, sign) into bucket order
//   4. accumulate: one thread per task (<= T consecutive entries of one bucket),
//                 XYZZ mixed additions (8M + 2S) over gathered affine bases
//   5. combine  : buckets that were split into several tasks are folded
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
    uint32_t task;     // T: max entries per accumulation task
    uint32_t lgrp;     // log2 of buckets per reduction group
};

// Signed digit of window w.  `carry` is threaded from window 0 upwards.
H2B_DI int32_t next_digit(const uint32_t (&s)[9], uint32_t w, uint32_t c, uint32_t &carry,
                          bool last_window) {
    uint32_t off = w * c;
    uint32_t idx = off >> 5, sh = off & 31;
    uint64_t two = ((uint64_t)s[idx + 1 < 9 ? idx + 1 : 8] << 32) | s[idx];
    uint32_t d = (uint32_t)(two >> sh) & ((1u << c) - 1);
    d += carry;
    if (!last_window && d > (1u << (c - 1))) {
        carry = 1;
        return (int32_t)d - (int32_t)(1u << c);
    }
    carry = 0;
    return (int32_t)d;
}

// mode 0: histogram; mode 1: scatter (cursor holds the running write position per bucket)
template <int MODE>
__global__ void msm_digits_kernel(const Fe *__restrict__ scalars, MsmCfg cfg,
                                  uint32_t *__restrict__ counts_or_cursor,
                                  uint32_t *__restrict__ sorted) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cfg.n) return;
    Fe s = Fr::from_mont(load_fe_ro(&scalars[i]));
    uint32_t l[9];
#pragma unroll
    for (int k = 0; k < 8; k++) l[k] = s.l[k];
    l[8] = 0;
    uint32_t carry = 0;
    for (uint32_t w = 0; w < cfg.windows; w++) {
        int32_t d = next_digit(l, w, cfg.c, carry, w + 1 == cfg.windows);
        if (d == 0) continue;
        uint32_t neg = d < 0;
        uint32_t mag = neg ? (uint32_t)(-d) : (uint32_t)d;
        uint32_t bucket = w * cfg.bpw + mag - 1;
        if (MODE == 0) {
            atomicAdd(&counts_or_cursor[bucket], 1u);
        } else {
            uint32_t pos = atomicAdd(&counts_or_cursor[bucket], 1u);
            sorted[pos] = i | (neg << 31);
        }
    }
}

// Exclusive scans over the nb buckets in three small launches (block sums, scan of block sums,
// rescan with offsets):
//   offsets[b] / cursor[b] = sum_{b' < b} counts[b']
//   task_off[b]            = sum_{b' < b} ceil(counts[b'] / T);  task_off[nb] = total
// Blocks of 1024 threads, `ipt` consecutive buckets per thread, at most 1024 blocks.
__global__ void __launch_bounds__(1024)
msm_scan_sums_kernel(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t T, uint32_t ipt,
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
            sb += (cnt + T - 1) / T;
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
// One block: exclusive scan of the (<= 1024) block sums in place; totals to task_off[nb].
__global__ void __launch_bounds__(1024)
msm_scan_blocks_kernel(uint2 *__restrict__ block_sums, uint32_t nblocks, uint32_t nb,
                       uint32_t *__restrict__ task_off) {
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
    if (tid == 1023) task_off[nb] = sh_b[1023];
}
__global__ void __launch_bounds__(1024)
msm_scan_apply_kernel(const uint32_t *__restrict__ counts, uint32_t nb, uint32_t T, uint32_t ipt,
                      const uint2 *__restrict__ block_sums, uint32_t *__restrict__ offsets,
                      uint32_t *__restrict__ cursor, uint32_t *__restrict__ task_off,
                      uint32_t *__restrict__ heavy /* [0] = count, [1..] = bucket ids */) {
    __shared__ uint32_t sh_a[1024], sh_b[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t lo = (blockIdx.x * 1024 + tid) * ipt;
    uint32_t sa = 0, sb = 0;
    for (uint32_t k = 0; k < ipt; k++) {
        uint32_t b = lo + k;
        if (b < nb) {
            uint32_t cnt = counts[b];
            sa += cnt;
            sb += (cnt + T - 1) / T;
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
            offsets[b] = ra;
            cursor[b] = ra;
            task_off[b] = rb;
            ra += cnt;
            rb += (cnt + T - 1) / T;
            if (cnt > T) heavy[1 + atomicAdd(&heavy[0], 1u)] = b;  // split into several tasks
        }
    }
}

// One thread per task.  task_off is non-decreasing; the owning bucket is the last b with
// task_off[b] <= t among buckets that have tasks (binary search for upper bound).
__global__ void __launch_bounds__(128)
msm_accumulate_kernel(const Affine *__restrict__ bases, const uint32_t *__restrict__ sorted,
                      const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ counts,
                      const uint32_t *__restrict__ task_off, MsmCfg cfg,
                      XYZZ *__restrict__ bucket_sums, XYZZ *__restrict__ partials) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t total = task_off[cfg.nb];
    if (t >= total) return;
    uint32_t lo = 0, hi = cfg.nb;  // find largest b in [0, nb) with task_off[b] <= t
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (task_off[mid] <= t) lo = mid; else hi = mid;
    }
    // lo may sit on a run of empty buckets sharing the same offset; the owner is the last one
    // of that run, which is what the search returns (largest b with task_off[b] <= t).
    uint32_t b = lo;
    uint32_t cnt = counts[b];
    uint32_t chunk = t - task_off[b];
    uint32_t begin = offsets[b] + chunk * cfg.task;
    uint32_t end = min(offsets[b] + cnt, begin + cfg.task);

    XYZZ acc = xyzz_identity();
    uint32_t e = begin;
    uint32_t v = sorted[e];
    Affine p = load_affine(&bases[v & 0x7fffffffu]);
    while (true) {
        uint32_t vn = 0;
        Affine pn;
        bool more = e + 1 < end;
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
    uint32_t ntasks = (cnt + cfg.task - 1) / cfg.task;
    if (ntasks == 1) store_xyzz(&bucket_sums[b], acc);
    else store_xyzz(&partials[t], acc);
}

// Buckets that were split into several tasks ("heavy": skewed scalars, or the short top window):
// one block per heavy bucket folds its partial sums -- threads stride over the partials, then a
// shared-memory tree.  The grid is persistent and walks the heavy list built by the scan.
__global__ void __launch_bounds__(128)
msm_combine_kernel(const uint32_t *__restrict__ counts, const uint32_t *__restrict__ task_off,
                   const uint32_t *__restrict__ heavy, MsmCfg cfg, const XYZZ *__restrict__ partials,
                   XYZZ *__restrict__ bucket_sums) {
    __shared__ uint4 comb_smem[128 * 8];
    XYZZ *sh = reinterpret_cast<XYZZ *>(comb_smem);
    const uint32_t nheavy = heavy[0], tid = threadIdx.x;
    for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
        const uint32_t b = heavy[1 + h];
        const uint32_t ntasks = (counts[b] + cfg.task - 1) / cfg.task;
        const uint32_t t0 = task_off[b];
        XYZZ acc = xyzz_identity();
        for (uint32_t k = tid; k < ntasks; k += 128) {
            XYZZ q = load_xyzz(&partials[t0 + k]);
            xyzz_add(acc, q);
        }
        store_xyzz(&sh[tid], acc);
        __syncthreads();
        for (uint32_t stride = 64; stride > 0; stride >>= 1) {
            if (tid < stride && tid + stride < ntasks) {
                XYZZ a = load_xyzz(&sh[tid]);
                XYZZ b2 = load_xyzz(&sh[tid + stride]);
                xyzz_add(a, b2);
                store_xyzz(&sh[tid], a);
            }
            __syncthreads();
        }
        if (tid == 0) store_xyzz(&bucket_sums[b], load_xyzz(&sh[0]));
        __syncthreads();
    }
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

// Horner over windows, high to low; result as a homogeneous projective point (96 B).
__global__ void msm_final_kernel(const XYZZ *__restrict__ window_sums, MsmCfg cfg, Projective *out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    XYZZ acc = xyzz_identity();
    for (int w = (int)cfg.windows - 1; w >= 0; w--) {
        for (uint32_t d = 0; d < cfg.c; d++) acc = xyzz_dbl_ni(acc);
        XYZZ s = load_xyzz(&window_sums[w]);
        xyzz_add(acc, s);
    }
    Projective j = xyzz_to_projective(acc);
    store_fe(&out->x, j.x);
    store_fe(&out->y, j.y);
    store_fe(&out->z, j.z);
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
