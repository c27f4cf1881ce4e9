// msm_reduce.cuh -- bucket reduction  sum_b (b + 1) * B_b  as a tree over the BITS of the bucket index.
//
// The running-sum reduction (msm.cuh, msm_reduce_kernel) is work-optimal but serial: 2 * 2^lgrp chained additions
// per group plus a double-and-add for the group offset, ~60-150 dependent group operations at ~5 us each for a
// lone thread, i.e. 0.4-0.9 ms however small the MSM -- at 2^16..2^20 points that was 20-45 % of a commit.
//
// Here   sum_b (b + 1) B_b  =  S + sum_j 2^j S_j,   S = sum_b B_b,   S_j = sum over buckets whose index has bit j.
// A node over 2^l consecutive buckets carries (S, S_0 .. S_{l-1}); merging the left (bit l clear) and the right
// (bit l set) node costs l + 1 independent additions and S_l = S_right is a copy, so the whole tree is
// sum_l (l + 1) / 2^(l+1) = 2 additions per bucket -- the same work as running sums -- at a depth of one
// addition per level, and every addition is shared by a team of 4 lanes (msm_comb.cuh) because most lanes of
// the upper levels would idle.  One block folds 2^lgT buckets; the per-block S feed the same kernel again (the
// block index is the high part of the bucket index), the per-block S_j ride along as extra rows of that launch
// (they only need their plain total), and a last kernel runs Horner over the c - 1 bit sums with cooperative
// doublings.
#pragma once
#include "msm.cuh"
#include "msm_comb.cuh"

namespace h2b {

// src[row][count] -> next[row][nblk] (block totals) and, for the first `bit_rows` rows, part[row][lgT][nblk] (block
// bit sums); nblk = gridDim.x = count >> lgT, rows = gridDim.y.  Rows beyond bit_rows are plain sums (the bit sums
// of earlier levels on their way to a single value).
template <int kThreads>
__global__ void __launch_bounds__(kThreads)
msm_bit_tree_kernel(const XYZZ *__restrict__ src, uint32_t count, uint32_t lgT, uint32_t bit_rows,
                    XYZZ *__restrict__ next, XYZZ *__restrict__ part, uint32_t lone_levels) {
    extern __shared__ uint4 bt_smem[];
    const uint32_t T = 1u << lgT, w = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x, tid = threadIdx.x;
    XYZZ *in = reinterpret_cast<XYZZ *>(bt_smem), *out = in + T;
    const XYZZ *leaves = src + (size_t)w * count + (size_t)blk * T;
    for (uint32_t i = tid; i < T; i += kThreads) store_xyzz(&in[i], load_xyzz(&leaves[i]));
    __syncthreads();
    for (uint32_t l = 0; l < lgT; l++) {
        const uint32_t per = l + 1, total = (T >> (l + 1)) * per;
        if (l < lone_levels) {
            for (uint32_t a = tid; a < total; a += kThreads) {
                const uint32_t node = a / per, v = a - node * per;
                XYZZ x = load_xyzz(&in[(2 * node) * per + v]);
                const XYZZ y = load_xyzz(&in[(2 * node + 1) * per + v]);
                if (v == 0) store_xyzz(&out[node * (per + 1) + per], y);
                xyzz_add(x, y);
                store_xyzz(&out[node * (per + 1) + v], x);
            }
        } else {
            const uint32_t team = tid >> 2, teams = kThreads >> 2;
            for (uint32_t base = 0; base < total; base += teams) {
                const uint32_t a = base + team;
                const bool active = a < total;
                const uint32_t mask = __ballot_sync(0xffffffffu, active);
                if (active) {
                    const uint32_t node = a / per, v = a - node * per;
                    const XYZZ x = load_xyzz(&in[(2 * node) * per + v]);
                    const XYZZ y = load_xyzz(&in[(2 * node + 1) * per + v]);
                    const XYZZ z = xyzz_add_team4(x, y, tid & 3, mask);
                    __syncwarp(mask);
                    if ((tid & 3) == 0) {
                        store_xyzz(&out[node * (per + 1) + v], z);
                        if (v == 0) store_xyzz(&out[node * (per + 1) + per], y);
                    }
                }
            }
        }
        __syncthreads();
        XYZZ *t = in;
        in = out;
        out = t;
    }
    if (tid <= lgT && (tid == 0 || w < bit_rows)) {
        const XYZZ v = load_xyzz(&in[tid]);
        if (tid == 0) store_xyzz(&next[(size_t)w * nblk + blk], v);
        else store_xyzz(&part[((size_t)w * lgT + (tid - 1)) * nblk + blk], v);
    }
}

// Optional first stage for very wide windows (throughput-bound): thread t folds the run of Q = 2^q consecutive
// buckets [t * Q, (t + 1) * Q) by running sums into S[t] = sum B and Wp[t] = sum (k + 1) B_(tQ + k).  Then
// sum_b (b + 1) B_b = sum_t Wp[t] + Q * sum_t t * S[t], and the second sum is the tree above over S.
__global__ void __launch_bounds__(128)
msm_bucket_runs_kernel(const XYZZ *__restrict__ buckets, uint32_t runs, uint32_t q, XYZZ *__restrict__ S,
                       XYZZ *__restrict__ Wp) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= runs) return;
    const XYZZ *bk = buckets + ((size_t)t << q);
    XYZZ running = xyzz_identity(), acc = xyzz_identity();
    for (int k = (1 << q) - 1; k >= 0; k--) {
        XYZZ b = load_xyzz(&bk[k]);
        xyzz_add(running, b);
        xyzz_add(acc, running);
    }
    store_xyzz(&S[t], running);
    store_xyzz(&Wp[t], acc);
}

// The bit sums of up to three tree levels (lowest bits first) and the grand total of every window.
struct BitSums {
    const XYZZ *rows[3];  // rows[i][w * lg[i] + j] = S_(bit offset of level i + j) of window w
    uint32_t lg[3];
    uint32_t levels;
    const XYZZ *total;    // total[w]
    uint32_t shift;       // the bit sums are of (bucket index >> shift): doublings after the Horner
};

// Doubling shared by a team of 4 adjacent lanes (xyzz_dbl_coop4 of msm.cuh for any team of a warp); p is
// replicated in the team and is not the identity.
H2B_DI void xyzz_dbl_team4(XYZZ &p, uint32_t lane4, uint32_t mask) {
    const Fe u = Fq::dbl(p.y);
    const Fe a1 = fe_sel(lane4 == 0, u, p.x);
    const Fe r1 = Fq::mul(a1, a1);
    const Fe v = fe_team(r1, 0, mask), xx = fe_team(r1, 1, mask);
    const Fe m = Fq::add(Fq::dbl(xx), xx);
    const Fe a2 = fe_sel(lane4 == 0, u, fe_sel(lane4 == 1, p.x, fe_sel(lane4 == 2, m, v)));
    const Fe b2 = fe_sel(lane4 == 2, m, fe_sel(lane4 == 3, p.zz, v));
    const Fe r2 = Fq::mul(a2, b2);
    const Fe w = fe_team(r2, 0, mask), sx = fe_team(r2, 1, mask), mm = fe_team(r2, 2, mask), zz3 = fe_team(r2, 3, mask);
    const Fe x3 = Fq::sub(Fq::sub(mm, sx), sx);
    const Fe a3 = fe_sel(lane4 == 2, m, w);
    const Fe b3 = fe_sel(lane4 == 0, p.y, fe_sel(lane4 == 1, p.zzz, Fq::sub(sx, x3)));
    const Fe r3 = Fq::mul(a3, b3);
    const Fe wy = fe_team(r3, 0, mask), zzz3 = fe_team(r3, 1, mask), msx = fe_team(r3, 2, mask);
    p.x = x3;
    p.y = Fq::sub(msx, wy);
    p.zz = zz3;
    p.zzz = zzz3;
}

// window_sums[w] = total + 2^shift * sum_j 2^j S_j.  The c - 1 doublings are a latency chain whichever way the sum
// is bracketed, but the additions are not: warp m runs Horner over bits [4m, 4m + 4) (lanes 0..3, team
// operations), then warp 0 runs Horner over those chunk values (4 doublings and one addition per chunk).
constexpr uint32_t kHornerChunk = 4, kHornerWarps = 8;
__global__ void __launch_bounds__(32 * kHornerWarps)
msm_bit_horner_kernel(BitSums bs, XYZZ *__restrict__ window_sums) {
    __shared__ uint4 horner_smem[kHornerWarps * 8];
    XYZZ *sh = reinterpret_cast<XYZZ *>(horner_smem);
    const uint32_t w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t nbits = 0;
    for (uint32_t i = 0; i < bs.levels; i++) nbits += bs.lg[i];
    const uint32_t nchunks = (nbits + kHornerChunk - 1) / kHornerChunk;  // <= kHornerWarps (c <= 24)
    if (lane < 4 && warp < nchunks) {
        const uint32_t lo = warp * kHornerChunk, hi = min(lo + kHornerChunk, nbits);
        XYZZ acc = xyzz_identity();
#pragma unroll 1
        for (int b = (int)hi - 1; b >= (int)lo; b--) {
            uint32_t lv = 0, j = (uint32_t)b;
            while (j >= bs.lg[lv]) j -= bs.lg[lv++];
            if (!xyzz_is_identity(acc)) xyzz_dbl_team4(acc, lane, 0xfu);  // uniform across the team
            const XYZZ s = load_xyzz(&bs.rows[lv][(size_t)w * bs.lg[lv] + j]);
            acc = xyzz_add_team4(acc, s, lane, 0xfu);
        }
        if (lane == 0) store_xyzz(&sh[warp], acc);
    }
    __syncthreads();
    if (warp != 0 || lane >= 4) return;
    XYZZ acc = xyzz_identity();
#pragma unroll 1
    for (int m = (int)nchunks - 1; m >= 0; m--) {
        if (!xyzz_is_identity(acc)) {
#pragma unroll 1
            for (uint32_t d = 0; d < kHornerChunk; d++) xyzz_dbl_team4(acc, lane, 0xfu);
        }
        const XYZZ s = load_xyzz(&sh[m]);
        acc = xyzz_add_team4(acc, s, lane, 0xfu);
    }
    if (!xyzz_is_identity(acc)) {
#pragma unroll 1
        for (uint32_t d = 0; d < bs.shift; d++) xyzz_dbl_team4(acc, lane, 0xfu);
    }
    const XYZZ t = load_xyzz(&bs.total[w]);
    acc = xyzz_add_team4(acc, t, lane, 0xfu);
    if (lane == 0) store_xyzz(&window_sums[w], acc);
}

}  // namespace h2b
