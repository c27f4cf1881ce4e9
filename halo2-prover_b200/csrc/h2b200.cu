// h2b200.cu -- context, launch planning and the C ABI of libh2b200.so (see include/h2b200.h).
//
// Everything arithmetic runs on the device.  The host side only plans launches,
// owns the workspace / SRS / twiddle caches and moves bytes.  There is no CPU
// implementation of the MSM or the NTT in this library and no fallback: a missing
// or failing GPU surfaces as H2B_ERR_CUDA.
#include "../../include/h2b200.h"

#include <cuda_runtime.h>

#include <algorithm>

#include <cstdio>
#include <chrono>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "curve.cuh"
#include "ecntt.cuh"
#include "evalh.cuh"
#include "hostcopy.h"
#include "field.cuh"
#include "msm.cuh"
#include "msm_comb.cuh"
#include "msm_reduce.cuh"
#include "ntt.cuh"

using namespace h2b;

static_assert(sizeof(Fe) == 32, "Fr/Fq must be 32 bytes (4 x u64)");
static_assert(sizeof(Affine) == 64, "G1Affine must be 64 bytes");
static_assert(sizeof(Projective) == 96, "G1 must be 96 bytes");
static_assert(sizeof(XYZZ) == 128, "bucket must be 128 bytes");

namespace {

thread_local std::string g_err;
std::mutex g_mu;

enum BufId {
    BUF_SCALARS = 0, BUF_BASES, BUF_COUNTS, BUF_CURSOR, BUF_NEOFF, BUF_NEID, BUF_SORTED, BUF_DIGITS, BUF_HEAD, BUF_TAIL,
    BUF_TAILJ, BUF_TOTALS, BUF_ECNTT, BUF_EVALH, BUF_EVALH_SCRATCH,
    BUF_BUCKETS, BUF_BUCKETS2, BUF_WINDOWS, BUF_WPART, BUF_BLOCKSUMS, BUF_HEAVY, BUF_OUT, BUF_GATHER, BUF_NTT_A, BUF_NTT_T, BUF_NTT_T2, BUF_NTT_IN,
    BUF_NTT_OUT, BUF_MISC,
    BUF_TEST_A, BUF_TEST_B, BUF_TEST_O, BUF_COUNT
};

struct TwKey {
    uint64_t w[4];
    uint32_t log_n;
    bool operator<(const TwKey &o) const {
        int c = memcmp(w, o.w, sizeof w);
        if (c) return c < 0;
        return log_n < o.log_n;
    }
};
// One device's share of a registered base array.
struct TwEntry {
    Fe *W = nullptr;          // omega^e, e in [0, 2^log_n)
    Fe *pass[11] = {};        // inner twiddles of a radix-2^S pass, per level (ntt_pass_twiddles_kernel), by S
};
struct Srs {
    int dev = 0;              // index into g_all
    size_t off = 0;           // first point of the share inside the registered array
    Affine *d = nullptr;      // the bases of the share (n x 64 B)
    size_t n = 0;
    Affine *table = nullptr;  // precomputed windows: table[v * n + i] = 2^(c*tstride*v) * d[i], or null
    uint32_t c = 0, windows = 0;
    uint32_t tstride = 1;     // every tstride-th window power is tabulated (msm.cuh: t bucket sets)
    // small SRS: every multiple d * 2^(comb_c * w) * d[i], d <= 2^(comb_c - 1) (msm_comb.cuh), or null
    Affine *comb = nullptr;
    uint32_t comb_c = 0, comb_w = 0;
};

// One persistent host thread per context: the multi-device paths hand each device's share of a call to its worker
// so that the copies and launches for different devices are issued concurrently (each over its own PCIe link).
class Worker {
  public:
    void start() {
        th_ = std::thread([this] { loop(); });
    }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = std::move(f);
            state_ = 1;
        }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return state_ == 0; });
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return stop_ || state_ == 1; });
                if (stop_) return;
                f = std::move(job_);
                state_ = 2;
            }
            f();
            {
                std::lock_guard<std::mutex> lk(mu_);
                state_ = 0;
            }
            cv_.notify_all();
        }
    }
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::function<void()> job_;
    int state_ = 0;  // 0 idle, 1 submitted, 2 running
    bool stop_ = false;
};

// A registered base array: one share on a single device, or -- with several devices (h2b_init_devices) -- either
// the whole array replicated on every device (small SRS: whole columns are dealt to the devices) or contiguous
// point ranges, one per device (large SRS: every commit is split by point range and the partial sums are folded).
struct SrsSet {
    size_t n = 0;
    bool replicated = false;
    std::vector<Srs> parts;   // parts[0] lives on the primary device
};

struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    void *buf[BUF_COUNT] = {};
    size_t cap[BUF_COUNT] = {};
    std::map<TwKey, TwEntry> twiddles;
    size_t twiddle_bytes = 0;
    std::map<uint64_t, h2b_domain> domains;
    uint64_t launches = 0;
    uint32_t msm_window = 0;
    uint32_t reduce_lgrp = 0;  // tuning override (H2B_REDUCE_LGRP)
    uint32_t reduce_tree = 1;  // bucket reduction: 1 = bit tree (msm_reduce.cuh), 0 = running sums (H2B_REDUCE_TREE)
    uint32_t reduce_lone = 2;  // tree levels done by lone threads on large grids (H2B_REDUCE_LONE)
    uint32_t min_slice = 16;   // shortest accumulation slice (H2B_MIN_SLICE)
    uint32_t min_waves = 1;    // fewest accumulation waves (H2B_MIN_WAVES)
    uint64_t max_entries = 1ull << 31;  // sorted entries per pass (H2B_MAX_ENTRIES_LOG lowers it for tests)
    std::set<const void *> attr_done;  // kernels whose dynamic shared-memory limit was raised on this device
    int reduce_q = -1;         // first-stage run length 2^q of the tree reduction, -1 = automatic (H2B_REDUCE_Q)
    size_t comb_max_n = (size_t)1 << 14;  // registered SRS up to this length get the bucket-free table
    uint32_t comb_c = 8;
    double e2e_ratio = 0;      // growth of the host-path chunk sizes (0 = automatic)
    double h2d_gbs = 55.0;     // pinned host -> this device, GB/s, with every device of the library copying at once (measured by
                               // h2b_init_devices; one device alone: the PCIe Gen5 x16 figure of this host)
    uint32_t srs_window = 0;   // 0 = automatic
    uint32_t srs_table_stride = 0;  // 0 = automatic (1 unless HBM is short), else every t-th window power is tabulated
    int srs_precompute = 1;
    int timing = 0;
    cudaEvent_t last_done = nullptr, copy_fence = nullptr;
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> chunk_events;
    uint32_t e2e_chunks = 4;
    bool e2e_auto = true;      // pieces by size (two below 2^23 points) until h2b_set_e2e_chunking / H2B_E2E_CHUNKS names a count
    size_t e2e_min_n = (size_t)1 << 21;
    cudaStream_t last_stream = nullptr;
    std::vector<cudaEvent_t> tev0, tev1;  // timing event pairs
    uint32_t tev_used = 0;
    int sm_count = 148;
    HostCopier *copier = nullptr;  // pageable host memory <-> HBM through worker threads and a pinned ring
    Worker worker;
};
// `g` is the context the calling thread works on.  API entry points set it to the primary context (ensure_ctx);
// the multi-device paths run one worker thread per device, each with its own `g` (so every internal function
// below is device-agnostic), while the API thread holds g_mu.
thread_local Ctx *g = nullptr;
Ctx *g_primary = nullptr;
std::vector<Ctx *> g_all;  // every context of h2b_init / h2b_init_devices; g_all[0] == g_primary
std::map<uint64_t, SrsSet> g_srs;
uint64_t g_next_handle = 1;
size_t g_shard_min_n = (size_t)1 << 21;  // registered arrays from this length up are sharded by point range (H2B_SHARD_MIN_LOG)

int fail(int code, const char *what, cudaError_t e = cudaSuccess) {
    char tmp[512];
    if (e != cudaSuccess) snprintf(tmp, sizeof tmp, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(tmp, sizeof tmp, "%s", what);
    g_err = tmp;
    return code;
}
#define CU(call)                                                             \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return fail(H2B_ERR_CUDA, #call, e__);       \
    } while (0)
#define LAUNCHED()                                                           \
    do {                                                                     \
        g->launches++;                                                       \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) return fail(H2B_ERR_CUDA, "kernel launch", e__); \
    } while (0)
#define TRY(expr)                        \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != H2B_OK) return rc__; \
    } while (0)

int ensure_ctx() {
    g = g_primary;
    if (g) return H2B_OK;
    return fail(H2B_ERR_STATE, "h2b_init has not been called");
}

// host <-> device copies of caller buffers (pinned buffers go straight to the DMA engine)
// round_trip: the call will read a result of similar size back (NTT entry points): stage the input through the ring too
int copy_in(void *dev, const void *host, size_t bytes, cudaStream_t s, bool round_trip = false) {
    cudaError_t e = g->copier->h2d(dev, host, bytes, s, round_trip);
    if (e != cudaSuccess) return fail(H2B_ERR_CUDA, "host-to-device copy", e);
    return H2B_OK;
}
// ordered after the work on `s`; returns when the host buffer is complete and `s` is idle
int copy_out(void *host, const void *dev, size_t bytes, cudaStream_t s) {
    cudaError_t e = g->copier->d2h(host, dev, bytes, s);
    if (e != cudaSuccess) return fail(H2B_ERR_CUDA, "device-to-host copy", e);
    return H2B_OK;
}
int get_buf(BufId id, size_t bytes, void **out) {
    if (bytes == 0) bytes = 16;
    if (g->cap[id] < bytes) {
        if (g->buf[id]) {
            CU(cudaDeviceSynchronize());  // work on any stream may still reference it
            CU(cudaFree(g->buf[id]));
            g->buf[id] = nullptr;
            g->cap[id] = 0;
        }
        size_t want = bytes + bytes / 8;  // a little headroom against regrowth
        cudaError_t e = cudaMalloc(&g->buf[id], want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&g->buf[id], want);
        }
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(H2B_ERR_OOM, "cudaMalloc(workspace)", e);
        }
        g->cap[id] = want;
    }
    *out = g->buf[id];
    return H2B_OK;
}

// The workspace is shared by every call, so work submitted on different streams is chained:
// a call on stream B waits for the previous call's work on stream A.
int enter(cudaStream_t s) {
    if (g->last_stream && g->last_stream != s) CU(cudaStreamWaitEvent(s, g->last_done, 0));
    return H2B_OK;
}
// One per entry point: orders the call after the previous one (enter) and records the context's `last_done`
// event on EVERY exit path, error returns included, so that the next call on another stream -- and
// h2b_srs_release / workspace regrowth -- wait for whatever this call managed to enqueue.
struct Scope {
    Ctx *c = nullptr;
    cudaStream_t s = nullptr;
    int begin(cudaStream_t st) {
        TRY(enter(st));
        c = g;
        s = st;
        return H2B_OK;
    }
    ~Scope() {
        if (c && cudaEventRecord(c->last_done, s) == cudaSuccess) c->last_stream = s;
    }
};

constexpr uint32_t kMaxTimed = 256;
void time_begin(cudaStream_t s) {
    if (!g->timing || g->tev_used >= kMaxTimed) return;
    if (g->tev0.size() <= g->tev_used) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
        g->tev0.push_back(a);
        g->tev1.push_back(b);
    }
    cudaEventRecord(g->tev0[g->tev_used], s);
}
void time_end(cudaStream_t s) {
    if (!g->timing || g->tev_used >= kMaxTimed || g->tev0.size() <= g->tev_used) return;
    cudaEventRecord(g->tev1[g->tev_used], s);
    g->tev_used++;
}

// ------------------------------------------------------------------------------ MSM
uint32_t ceil_log2(size_t n) {
    uint32_t l = 0;
    while (((size_t)1 << l) < n) l++;
    return l;
}

uint32_t msm_windows_for(uint32_t c) {
    uint32_t W = (254 + c - 1) / c;
    uint32_t top_bits = 254 - (W - 1) * c;
    if (top_bits > c - 1) W += 1;  // the (unsigned) top digit plus carry must fit 2^(c-1)
    return W;
}

// Window geometry.  srs == nullptr: independent windows (one bucket set per window, Horner at the
// end).  srs with a precomputed table: its c, all windows share one bucket set.
MsmCfg msm_plan(size_t n, const Srs *srs = nullptr, uint32_t cols = 1) {
    MsmCfg cfg{};
    cfg.n = (uint32_t)n;
    cfg.cols = cols;
    uint32_t c;
    if (srs && srs->table) {
        c = srs->c;
        cfg.shared = srs->tstride;
        cfg.stride = (uint32_t)srs->n;
    } else {
        uint32_t lg = ceil_log2(n);
        // window: measured optimum on B200 is lg(n) - 5 for n >= 2^19 (17 at 2^22..2^24; beyond that the
        // bucket reduction and the scatter cost more than the saved window gains)
        c = g->msm_window ? g->msm_window : (lg >= 19 ? lg - 5 : (lg > 4 ? lg - 4 : 0));
        if (!g->msm_window) {
            if (c < 4) c = 4;
            if (c > 17) c = 17;
        }
        if (c < 2) c = 2;
        if (c > 22) c = 22;
    }
    cfg.c = c;
    uint32_t W = msm_windows_for(c);
    cfg.windows = W;
    cfg.bpw = 1u << (c - 1);
    cfg.nb = cfg.shared ? cols * cfg.shared * cfg.bpw : W * cfg.bpw;
    for (uint32_t w = 0; w + 1 < W; w++) {
        uint32_t bit = c * w + c - 1;
        cfg.half[bit >> 5] |= 1u << (bit & 31);
    }
    // reduction groups of 2^lgrp buckets: 16 per group once a window has >= 4096 buckets
    uint32_t lgrp = 0;
    while (lgrp < 4 && (cfg.bpw >> lgrp) > 256) lgrp++;
    // very wide windows (shared bucket sets of a window table): larger groups amortise the double-and-add of
    // the group offset while still leaving >= 2^15 group chains for the GPU
    while (lgrp < 6 && ((uint64_t)cfg.bpw * cols >> lgrp) > (1u << 15)) lgrp++;
    if (g->reduce_lgrp) lgrp = g->reduce_lgrp;
    cfg.lgrp = lgrp;
    return cfg;
}

// Shared-bucket window for a registered SRS of n points (tunable: H2B_SRS_WINDOW).
uint32_t srs_window_for(size_t n) {
    if (g->srs_window) return g->srs_window;
    // Measured on B200 (scripts/probe.py, PROBE_COMMIT / PROBE_SRS_C sweeps).  Only widths whose top window
    // keeps >= 12 of the scalar's 254 bits are used from 2^15 points up (15, 16, 17, 20, 22, 24): a top window
    // of one or two bits sends every point to the same few buckets (hot atomics, buckets spanning thousands
    // of slices).  Small SRS are latency-bound: few buckets keep the reduction chains short.
    const uint32_t lg = ceil_log2(n);
    if (lg <= 14) return lg >= 8 ? lg - 2 : 6;  // single commits are flat in c here (latency chains); batches want lg - 2
    if (lg <= 16) return 16;
    if (lg <= 19) return 17;  // 2^19: 1.82 ms at 17 bits, 1.92 ms at 20 (2^19 buckets for 6.8 M entries)
    if (lg <= 25) return 20;
    return 22;
}

// An MSM is: begin (clear the buckets) -> one or more chunks over contiguous point ranges, each
// adding into the same buckets -> finish (bucket reduction, window fold, Horner).  Chunking lets the
// host-buffer entry points overlap the H2D copy of chunk k+1 with the accumulation of chunk k.
struct MsmRun {
    MsmCfg cfg;       // window geometry from the TOTAL size
    XYZZ *buckets;    // running bucket sums (all-zero = identity)
    uint32_t chunks_done = 0;
};

int msm_identity_out(Projective *d_out, cudaStream_t s) {
    // G1::identity() = (0, R, 0)
    Projective id;
    memset(&id, 0, sizeof id);
    const uint32_t one[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                             0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    memcpy(id.y.l, one, sizeof one);
    CU(cudaMemcpyAsync(d_out, &id, sizeof id, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));  // `id` lives on this stack frame
    return H2B_OK;
}

int msm_begin(size_t n_total, MsmRun *run, cudaStream_t s, const Srs *srs = nullptr, uint32_t cols = 1) {
    if (n_total > (1u << 30)) return fail(H2B_ERR_ARG, "msm: n > 2^30 not supported");
    run->cfg = msm_plan(n_total, srs, cols);
    TRY(get_buf(BUF_BUCKETS, (size_t)run->cfg.nb * sizeof(XYZZ), (void **)&run->buckets));
    CU(cudaMemsetAsync(run->buckets, 0, (size_t)run->cfg.nb * sizeof(XYZZ), s));  // all-zero XYZZ = identity
    return H2B_OK;
}

// Points [0, m) of (d_scalars, d_bases): digits -> scan -> scatter -> accumulate -> fix-up.
// In shared-bucket mode d_bases is the SRS table and `ioff` the chunk's first point inside the SRS.
int msm_chunk(MsmRun &run, const Fe *d_scalars, const Affine *d_bases, size_t m, cudaStream_t s, size_t ioff = 0) {
    if (m == 0) return H2B_OK;
    MsmCfg cfg = run.cfg;
    // Positions in the sorted entry list (cursor, ne_off, totals) are 32-bit: a chunk of 2^31 entries or more
    // (from ~2^27 points at 15 windows) is split by point range; the halves add into the same buckets.
    if ((uint64_t)m * cfg.cols * cfg.windows >= g->max_entries) {
        if (cfg.cols > 1 || m < 2) return fail(H2B_ERR_ARG, "msm: batch too large for one pass");
        size_t half = m / 2;
        if (half >= 512) half &= ~(size_t)255;
        TRY(msm_chunk(run, d_scalars, d_bases, half, s, ioff));
        return msm_chunk(run, d_scalars + half, cfg.shared ? d_bases : d_bases + half, m - half, s, ioff + half);
    }
    cfg.n = (uint32_t)m;
    cfg.ioff = (uint32_t)ioff;
    size_t entries = m * cfg.cols * cfg.windows;
    // Slice length.  Every slice is the same amount of work and 4 blocks of 128 slices are resident per SM, so the
    // launch runs in lockstep waves and the slices are sized so that the last wave is full (2^24 x 13 entries:
    // 5.6 waves at L = 512, 6.0 at L = 480).  How many waves is a trade: more (shorter slices) let late blocks
    // backfill and speed the accumulation up by up to 10 %, but every slice leaves a head and a tail piece for the
    // fix-up tree.  Measured optimum on B200 (commits, 2^17..2^24 points): one wave up to ~5 M entries, then about
    // one wave per 64 entries of slice length, 6 at most; slices never longer than 640 entries.
    uint32_t L;
    {
        const size_t per_wave = (size_t)g->sm_count * 4 * 128;
        size_t waves = std::min<size_t>(6, (entries + per_wave * 32) / (per_wave * 64));
        waves = std::max<size_t>(waves, (entries + per_wave * 640 - 1) / (per_wave * 640));
        waves = std::max<size_t>(waves, g->min_waves);
        const size_t fit = (entries + waves * per_wave - 1) / (waves * per_wave);
        L = (uint32_t)std::min<size_t>(640, std::max<size_t>(g->min_slice, fit));
    }
    cfg.slice = L;
    size_t max_slices = entries / cfg.slice + 1;
    uint32_t *counts, *cursor, *ne_off, *ne_id, *sorted, *totals, *heavy, *digits;
    int32_t *tail_j, *head_j;
    XYZZ *head, *tail;
    uint2 *block_sums;
    TRY(get_buf(BUF_COUNTS, (size_t)cfg.nb * 4, (void **)&counts));
    TRY(get_buf(BUF_CURSOR, (size_t)cfg.nb * 4, (void **)&cursor));
    TRY(get_buf(BUF_NEOFF, ((size_t)cfg.nb + 1) * 4, (void **)&ne_off));
    TRY(get_buf(BUF_NEID, ((size_t)cfg.nb + 1) * 4, (void **)&ne_id));
    TRY(get_buf(BUF_SORTED, entries * 4, (void **)&sorted));
    TRY(get_buf(BUF_DIGITS, entries * 4, (void **)&digits));
    TRY(get_buf(BUF_HEAD, max_slices * sizeof(XYZZ), (void **)&head));
    TRY(get_buf(BUF_TAIL, max_slices * sizeof(XYZZ), (void **)&tail));
    TRY(get_buf(BUF_TAILJ, 2 * max_slices * 4, (void **)&tail_j));
    head_j = tail_j + max_slices;
    TRY(get_buf(BUF_BLOCKSUMS, 1024 * sizeof(uint2), (void **)&block_sums));
    TRY(get_buf(BUF_TOTALS, 16, (void **)&totals));
    TRY(get_buf(BUF_HEAVY, (max_slices + 2) * 4, (void **)&heavy));

    // the first chunk fills the running buckets directly; later chunks fill a scratch set that a
    // uniform merge kernel adds in (doing that addition inside the accumulation loop would stall
    // whole warps on every lane's bucket boundary)
    XYZZ *target = run.buckets;
    if (run.chunks_done > 0) {
        TRY(get_buf(BUF_BUCKETS2, (size_t)cfg.nb * sizeof(XYZZ), (void **)&target));
        CU(cudaMemsetAsync(target, 0, (size_t)cfg.nb * sizeof(XYZZ), s));
    }
    CU(cudaMemsetAsync(counts, 0, (size_t)cfg.nb * 4, s));
    CU(cudaMemsetAsync(heavy, 0, 4, s));
    CU(cudaMemsetAsync(tail_j, 0xff, 2 * max_slices * 4, s));
    uint32_t nblk = (uint32_t)((m * cfg.cols + 255) / 256);
    msm_digits_kernel<<<nblk, 256, 0, s>>>(d_scalars, cfg, counts, digits);
    LAUNCHED();
    uint32_t ipt = (cfg.nb + 1024 * 1024 - 1) / (1024 * 1024);
    uint32_t sblocks = (cfg.nb + 1024 * ipt - 1) / (1024 * ipt);
    msm_scan_sums_kernel<<<sblocks, 1024, 0, s>>>(counts, cfg.nb, ipt, block_sums);
    LAUNCHED();
    msm_scan_blocks_kernel<<<1, 1024, 0, s>>>(block_sums, sblocks, totals);
    LAUNCHED();
    msm_scan_apply_kernel<<<sblocks, 1024, 0, s>>>(counts, cfg.nb, ipt, block_sums, cursor, ne_off, ne_id);
    LAUNCHED();
    msm_scatter_kernel<<<dim3(nblk, cfg.windows), 256, 0, s>>>(digits, cfg, cursor, sorted);
    LAUNCHED();
    time_begin(s);
    uint32_t ablocks = (uint32_t)((max_slices + 127) / 128);
    msm_accumulate_kernel<<<ablocks, 128, 0, s>>>(d_bases, sorted, ne_off, ne_id, totals, cfg, target, head, tail,
                                                  tail_j, head_j);
    LAUNCHED();
    time_end(s);
    for (uint32_t r = 0; r < kFixupLevels; r++) {
        msm_fixup_level_kernel<<<ablocks, 128, 0, s>>>(ne_off, ne_id, totals, cfg, head, tail, tail_j, head_j, target,
                                                       heavy, r);
        LAUNCHED();
    }
    msm_fixup_heavy_kernel<<<g->sm_count * 4, 128, 0, s>>>(ne_off, ne_id, cfg, head, tail, tail_j, heavy, target);
    LAUNCHED();
    if (run.chunks_done > 0) {
        msm_merge_kernel<<<(cfg.nb + 127) / 128, 128, 0, s>>>(run.buckets, target, cfg.nb);
        LAUNCHED();
    }
    run.chunks_done++;
    return H2B_OK;
}

// Bucket reduction as a tree over the bits of the bucket index (msm_reduce.cuh): windows[w] = sum_b (b + 1) B_b.
// Every level is one launch over ROWS: the first nwin rows are the windows (they produce bit sums), the others
// are the bit sums of earlier levels (and the run sums of the optional first stage), which only need their total;
// a level appends its nwin * lg bit-sum rows after the rows it received, so the last level leaves one array of
// single values: [per-window totals | (run sums) | bit sums of level 1 | level 2 | ...].
int msm_reduce_tree(const XYZZ *buckets, uint32_t bpw, uint32_t nwin, XYZZ *windows, cudaStream_t s) {
    constexpr int kThreads = 256;
    constexpr uint32_t kLgT = 8;
    if (!g->attr_done.count((const void *)msm_bit_tree_kernel<kThreads>)) {  // per context: the attribute belongs to its device
        CU(cudaFuncSetAttribute(msm_bit_tree_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(2 * sizeof(XYZZ) << kLgT)));
        g->attr_done.insert((const void *)msm_bit_tree_kernel<kThreads>);
    }
    // wide windows are throughput-bound: fold runs of 2^q buckets by lone threads first (full lane efficiency)
    // (measured: commits at 2^18 / 2^20 / 2^22 points, q = 0 / 2 / 4: 1.23 / 1.22 / 1.36, 3.52 / 3.39 / 3.20,
    // 10.82 / 10.68 / 10.50 ms)
    const uint64_t all = (uint64_t)bpw * nwin;
    uint32_t q = g->reduce_q >= 0 ? (uint32_t)g->reduce_q : (all >= (1u << 19) ? 4 : all >= (1u << 16) ? 2 : 0);
    while (q && (bpw >> q) < 2) q--;
    const uint32_t leaves = bpw >> q;
    size_t need = q ? (size_t)2 * nwin * leaves : 0;
    {
        size_t rows = q ? 2 * nwin : nwin;
        for (uint32_t cnt = leaves; cnt > 1;) {
            const uint32_t lg = std::min(kLgT, ceil_log2(cnt)), nblk = cnt >> lg;
            rows += (size_t)nwin * lg;
            need += rows * nblk;
            cnt = nblk;
        }
        need += nwin;  // leaves == 1
    }
    XYZZ *scratch;
    TRY(get_buf(BUF_WPART, need * sizeof(XYZZ), (void **)&scratch));
    BitSums bs{};
    bs.shift = q;
    const XYZZ *cur = buckets;
    uint32_t rows = nwin;
    if (q) {
        const uint32_t runs = nwin * leaves;
        XYZZ *S = scratch, *Wp = S + runs;
        scratch = Wp + runs;
        msm_bucket_runs_kernel<<<(runs + 127) / 128, 128, 0, s>>>(buckets, runs, q, S, Wp);
        LAUNCHED();
        cur = S;
        rows = 2 * nwin;
    }
    uint32_t bit_off[3] = {0, 0, 0};
    for (uint32_t cnt = leaves; cnt > 1;) {
        const uint32_t lg = std::min(kLgT, ceil_log2(cnt)), nblk = cnt >> lg;
        if (bs.levels >= 3) return fail(H2B_ERR_ARG, "msm: bucket window too wide for the reduction tree");
        XYZZ *next = scratch, *part = next + (size_t)rows * nblk;
        scratch = part + (size_t)nwin * lg * nblk;
        // on large grids the two lowest levels have one addition per two leaves, which lone threads do with less
        // overhead than teams
        const uint32_t lone = (uint64_t)nblk * rows >= 4u * g->sm_count ? g->reduce_lone : 0;
        msm_bit_tree_kernel<kThreads><<<dim3(nblk, rows), kThreads, 2 * sizeof(XYZZ) << lg, s>>>(cur, cnt, lg, nwin, next,
                                                                                                  part, lone);
        LAUNCHED();
        bit_off[bs.levels] = rows;
        bs.lg[bs.levels] = lg;
        bs.levels++;
        rows += nwin * lg;
        cur = next;
        cnt = nblk;
    }
    for (uint32_t i = 0; i < bs.levels; i++) bs.rows[i] = cur + bit_off[i];
    bs.total = q ? cur + nwin : cur;
    msm_bit_horner_kernel<<<nwin, 32 * kHornerWarps, 0, s>>>(bs, windows);
    LAUNCHED();
    return H2B_OK;
}

int msm_finish(const MsmRun &run, Projective *d_out, cudaStream_t s) {
    MsmCfg cfg = run.cfg;
    if (cfg.shared) cfg.windows = cfg.cols * cfg.shared;  // t bucket sets per column (t = 1: sum_k k * B_k is the result)
    XYZZ *windows, *wpart;
    TRY(get_buf(BUF_WINDOWS, (size_t)cfg.windows * sizeof(XYZZ), (void **)&windows));
    if (g->reduce_tree) {
        TRY(msm_reduce_tree(run.buckets, cfg.bpw, cfg.windows, windows, s));
    } else {
        uint32_t G = cfg.bpw >> cfg.lgrp;          // groups per window
        uint32_t rthreads = G < 256 ? G : 256;      // power of two
        // few groups in total (small MSMs): narrower blocks put the serial group chains on different SMs
        while (rthreads > 32 && (uint64_t)(G / rthreads) * cfg.windows < 2u * g->sm_count) rthreads >>= 1;
        uint32_t per_window = G / rthreads;         // blocks (= partials) per window
        TRY(get_buf(BUF_WPART, (size_t)cfg.windows * per_window * sizeof(XYZZ), (void **)&wpart));
        msm_reduce_kernel<<<dim3(per_window, cfg.windows), rthreads, rthreads * sizeof(XYZZ), s>>>(run.buckets, cfg, wpart);
        LAUNCHED();
        msm_window_fold_kernel<<<cfg.windows, 32, 0, s>>>(wpart, per_window, windows);
        LAUNCHED();
    }
    if (cfg.shared && cfg.cols > 1) {
        msm_batch_out_kernel<<<(cfg.cols + 31) / 32, 32, 0, s>>>(windows, cfg.cols, cfg.shared, cfg.c, d_out);
        LAUNCHED();
        return H2B_OK;
    }
    if (cfg.shared) cfg.windows = cfg.shared;  // Horner over the t bucket-set sums, c doublings each
    msm_final_kernel<<<1, 32, 0, s>>>(windows, cfg, d_out);
    LAUNCHED();
    return H2B_OK;
}

int msm_run(const Fe *d_scalars, const Affine *d_bases, size_t n, Projective *d_out, cudaStream_t s) {
    if (n == 0) return msm_identity_out(d_out, s);
    MsmRun run;
    TRY(msm_begin(n, &run, s));
    TRY(msm_chunk(run, d_scalars, d_bases, n, s));
    return msm_finish(run, d_out, s);
}

// Host scalars (and optionally host bases) -> device in `chunks` pieces on the copy stream while
// the compute stream works on the pieces that have landed.
int msm_run_pipelined(const uint64_t *h_scalars, const uint64_t *h_bases, const Srs *srs, size_t n,
                      Projective *d_out) {
    cudaStream_t s = g->stream;
    if (n == 0) return msm_identity_out(d_out, s);
    // pieces: every piece pays the fixed latency of a sort + fix-up pass (~0.3 ms), so below 2^23 points -- a device's
    // share of a 2^24-point commit on 4 or 8 GPUs -- two pieces hide the copy better than four
    // Growth of the piece sizes = accumulation time per point over copy time per point (each piece copies while its
    // predecessor accumulates): ~4 when this device has the host's PCIe bandwidth to itself, less when the devices of
    // one process share it (8 x B200: 24 GB/s each instead of 55).  A slower copy wants more, more even pieces.
    const double copy_ns = 32.0 / g->h2d_gbs, auto_ratio = std::min(4.0, std::max(1.3, 2.3 / copy_ns));
    uint32_t chunks = n >= g->e2e_min_n ? ((g->e2e_auto && n < ((size_t)1 << 23)) ? (auto_ratio < 3.0 ? 3u : 2u) : g->e2e_chunks) : 1;
    Fe *ds;
    Affine *db = nullptr;
    TRY(get_buf(BUF_SCALARS, n * sizeof(Fe), (void **)&ds));
    if (h_bases) TRY(get_buf(BUF_BASES, n * sizeof(Affine), (void **)&db));
    const bool shared = srs && srs->table;
    const Affine *bases = h_bases ? db : (shared ? srs->table : srs->d);
    MsmRun run;
    TRY(msm_begin(n, &run, s, h_bases ? nullptr : srs));
    if (chunks <= 1) {
        TRY(copy_in(ds, h_scalars, n * sizeof(Fe), s));
        if (h_bases) TRY(copy_in(db, h_bases, n * sizeof(Affine), s));
        TRY(msm_chunk(run, ds, bases, n, s));
        return msm_finish(run, d_out, s);
    }
    if (g->copier->stages(h_scalars)) {
        // pageable scalars: the staged copy of a chunk is host work that overlaps the accumulation of the
        // previous chunk already running on the GPU
        std::vector<size_t> clo, chi;
        {
            // staged copies run at roughly half the pinned rate: chunks grow by 2 (measured: 44.6 ms for 2^24
            // pageable scalars against 45.8 ms from pinned memory and 89 ms through the driver's staging)
            const double ratio = g->e2e_ratio > 0 ? g->e2e_ratio : (h_bases ? 1.25 : 2.0);
            double denom = 0, pw = 1;
            for (uint32_t k = 0; k < chunks; k++) { denom += pw; pw *= ratio; }
            pw = 1;
            size_t at = 0;
            for (uint32_t k = 0; k < chunks && at < n; k++) {
                size_t m = k + 1 == chunks ? n - at : (((size_t)((double)n * pw / denom) + 255) & ~(size_t)255);
                if (m == 0) m = 256;
                if (at + m > n) m = n - at;
                clo.push_back(at);
                chi.push_back(at + m);
                at += m;
                pw *= ratio;
            }
            if (at < n) chi.back() = n;
        }
        for (size_t k = 0; k < clo.size(); k++) {
            const size_t m = chi[k] - clo[k];
            TRY(copy_in(ds + clo[k], h_scalars + 4 * clo[k], m * sizeof(Fe), s));
            if (h_bases) TRY(copy_in(db + clo[k], h_bases + 8 * clo[k], m * sizeof(Affine), s));
            if (shared) TRY(msm_chunk(run, ds + clo[k], bases, m, s, clo[k]));
            else TRY(msm_chunk(run, ds + clo[k], bases + clo[k], m, s));
        }
        return msm_finish(run, d_out, s);
    }
    // the copy stream must not overwrite staging that earlier work on `s` may still read
    CU(cudaEventRecord(g->copy_fence, s));
    CU(cudaStreamWaitEvent(g->copy_stream, g->copy_fence, 0));
    // Chunk sizes grow geometrically: the first copy is the only one nothing overlaps, so it is small, and
    // every later chunk is as large as the accumulation of its predecessor can hide (the ratio is compute
    // time per point over copy time per point: ~4 with resident bases, ~1.5 when the bases travel too).
    const double ratio = g->e2e_ratio > 0 ? g->e2e_ratio : (h_bases ? 1.5 : auto_ratio);
    double denom = 0, pw = 1;
    for (uint32_t k = 0; k < chunks; k++) { denom += pw; pw *= ratio; }
    std::vector<size_t> lo, hi;
    pw = 1;
    size_t at = 0;
    for (uint32_t k = 0; k < chunks && at < n; k++) {
        size_t m = k + 1 == chunks ? n - at : (size_t)((double)n * pw / denom);
        m = (m + 255) & ~(size_t)255;
        if (m == 0) m = 256;
        if (at + m > n) m = n - at;
        lo.push_back(at);
        hi.push_back(at + m);
        at += m;
        pw *= ratio;
    }
    if (at < n) hi.back() = n;
    while (g->chunk_events.size() < lo.size()) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g->chunk_events.push_back(e);
    }
    for (size_t k = 0; k < lo.size(); k++) {
        size_t m = hi[k] - lo[k];
        CU(cudaMemcpyAsync(ds + lo[k], h_scalars + 4 * lo[k], m * sizeof(Fe), cudaMemcpyHostToDevice, g->copy_stream));
        if (h_bases)
            CU(cudaMemcpyAsync(db + lo[k], h_bases + 8 * lo[k], m * sizeof(Affine), cudaMemcpyHostToDevice, g->copy_stream));
        CU(cudaEventRecord(g->chunk_events[k], g->copy_stream));
    }
    for (size_t k = 0; k < lo.size(); k++) {
        CU(cudaStreamWaitEvent(s, g->chunk_events[k], 0));
        if (shared) TRY(msm_chunk(run, ds + lo[k], bases, hi[k] - lo[k], s, lo[k]));
        else TRY(msm_chunk(run, ds + lo[k], bases + lo[k], hi[k] - lo[k], s));
    }
    return msm_finish(run, d_out, s);
}

// ------------------------------------------------------------------------------ NTT
int get_twiddles(const uint64_t omega[4], uint32_t log_n, cudaStream_t s, TwEntry **out) {
    TwKey key;
    memcpy(key.w, omega, 32);
    key.log_n = log_n;
    auto it = g->twiddles.find(key);
    if (it != g->twiddles.end()) {
        *out = &it->second;
        return H2B_OK;
    }
    size_t n = (size_t)1 << log_n;
    size_t bytes = n * sizeof(Fe);
    if (g->twiddle_bytes + bytes > ((size_t)24 << 30)) {  // bound the cache
        CU(cudaDeviceSynchronize());
        for (auto &kv : g->twiddles) {
            cudaFree(kv.second.W);
            for (Fe *t : kv.second.pass)
                if (t) cudaFree(t);
        }
        g->twiddles.clear();
        g->twiddle_bytes = 0;
    }
    Fe *W = nullptr;
    cudaError_t e = cudaMalloc(&W, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(H2B_ERR_OOM, "cudaMalloc(twiddles)", e);
    }
    uint32_t lo_bits = log_n < 10 ? log_n : 10;
    uint32_t n_lo = 1u << lo_bits, n_hi = 1u << (log_n - lo_bits);
    Fe *small;
    TRY(get_buf(BUF_MISC, ((size_t)n_lo + n_hi) * sizeof(Fe), (void **)&small));
    Fe w;
    memcpy(&w, omega, 32);
    uint32_t mx = n_lo > n_hi ? n_lo : n_hi;
    ntt_pow_small_kernel<<<(mx + 127) / 128, 128, 0, s>>>(w, lo_bits, n_lo, n_hi, small, small + n_lo);
    LAUNCHED();
    ntt_pow_table_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(small, small + n_lo, lo_bits,
                                                                     (uint32_t)n, W);
    LAUNCHED();
    TwEntry ent;
    ent.W = W;
    auto ins = g->twiddles.emplace(key, ent);
    g->twiddle_bytes += bytes;
    *out = &ins.first->second;
    return H2B_OK;
}
// The per-level inner twiddle tables of a radix-2^S pass of a 2^log_n transform (built on first use).
int get_pass_twiddles(TwEntry *ent, uint32_t log_n, uint32_t S, cudaStream_t s, const Fe **out) {
    if (!ent->pass[S]) {
        Fe *t = nullptr;
        cudaError_t e = cudaMalloc(&t, ((size_t)1 << S) * sizeof(Fe));
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(H2B_ERR_OOM, "cudaMalloc(pass twiddles)", e);
        }
        ntt_pass_twiddles_kernel<<<((1u << S) + 127) / 128, 128, 0, s>>>(ent->W, log_n, S, t);
        LAUNCHED();
        ent->pass[S] = t;
    }
    *out = ent->pass[S];
    return H2B_OK;
}

// Which exchanges of a pass stay inside a warp: bit r is set when, after round r, every thread reads only elements
// written by threads of its own warp (then __syncwarp orders them; otherwise the round ends with a block barrier).
// Enumerated once per tile shape: the row sets of the rounds are those of ntt_pass_kernel.
uint32_t ntt_warp_sync_mask(uint32_t S, uint32_t C, uint32_t EL) {
    static std::map<uint32_t, uint32_t> cache;
    static std::mutex cache_mu;  // the per-device workers of a dealt batch plan their passes concurrently
    std::lock_guard<std::mutex> lk(cache_mu);
    const uint32_t key = S | (C << 8) | (EL << 24);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    const uint32_t E = 1u << EL, groups = ((1u << S) * C) >> EL;
    std::vector<uint32_t> ss;  // s_r of every round
    for (uint32_t lvl = 0; lvl < S;) {
        const uint32_t er = std::min(EL, S - lvl);
        ss.push_back(er < EL ? 0u : S - lvl - EL);
        lvl += er;
    }
    auto elem = [&](uint32_t gidx, uint32_t t, uint32_t s) {
        const uint32_t col = gidx % C, rest = gidx / C, lo = rest & ((1u << s) - 1u), hi = rest >> s;
        return (((hi << (s + EL)) | (t << s) | lo) * C) + col;
    };
    uint32_t mask = 0;
    std::vector<uint32_t> owner((size_t)(1u << S) * C);
    for (size_t r = 0; r + 1 < ss.size(); r++) {
        for (uint32_t gidx = 0; gidx < groups; gidx++)
            for (uint32_t t = 0; t < E; t++) owner[elem(gidx, t, ss[r])] = gidx;
        bool local = true;
        for (uint32_t gidx = 0; gidx < groups && local; gidx++)
            for (uint32_t t = 0; t < E; t++)
                if ((owner[elem(gidx, t, ss[r + 1])] >> 5) != (gidx >> 5)) {
                    local = false;
                    break;
                }
        if (local) mask |= 1u << r;
    }
    cache[key] = mask;
    return mask;
}

template <int S, int C, int EL, int NT, int MINB = 1, int VAR = 0>
int launch_pass(const Fe *in, Fe *out, const Fe *W, const Fe *TW, uint32_t log_n, uint32_t log_ns, bool last,
                const NttIo &io, cudaStream_t s, uint32_t batch) {
    static_assert(S >= EL && S <= 10, "a round needs EL levels; tiles hold at most 2^10 rows");
    constexpr int R = 1 << S;
    size_t smem = ((size_t)2 * R * C + 2 * (R >> EL)) * sizeof(uint4);  // tile + the later rounds' twiddles
    if (VAR == 1) smem += (size_t)2 * R * C * sizeof(uint4) + 16;           // + the TMA staging area and its mbarrier
    if (smem > 48 * 1024 && !g->attr_done.count((const void *)ntt_pass_kernel<S, C, EL, NT, MINB, VAR>)) {
        CU(cudaFuncSetAttribute(ntt_pass_kernel<S, C, EL, NT, MINB, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)smem));
        g->attr_done.insert((const void *)ntt_pass_kernel<S, C, EL, NT, MINB, VAR>);
    }
    const uint32_t M = 1u << (log_n - S);
    const uint32_t blocks = M / C;
    if (blocks == 0) return fail(H2B_ERR_ARG, "ntt: tile wider than the pass");
    ntt_pass_kernel<S, C, EL, NT, MINB, VAR><<<dim3(blocks, batch), NT, smem, s>>>(in, out, W, TW, log_n, log_ns, last ? 1u : 0u,
                                                                      ntt_warp_sync_mask(S, C, EL), io);
    LAUNCHED();
    return H2B_OK;
}

// Tile shapes.  A transform of up to 2^10 elements is one single-column tile.  Multi-pass transforms use tiles of
// 2^TL elements with 2^EL elements per thread: TL = 10 is the throughput shape (EL = 3 / 2 / 1 on 128 / 256 / 512
// threads), TL = 9 and 8 on 128 threads give more and smaller blocks to transforms that would otherwise leave most
// SMs idle (those are latency-bound, not throughput-bound).
extern int g_ntt_dense, g_ntt_variant;
int dispatch_pass(uint32_t S, uint32_t EL, uint32_t TL, bool single, const Fe *in, Fe *out, const Fe *W, const Fe *TW,
                  uint32_t log_n, uint32_t log_ns, bool last, const NttIo &io, cudaStream_t s, uint32_t batch) {
#define H2B_PASS(S_, C_, EL_, NT_) return launch_pass<S_, C_, EL_, NT_>(in, out, W, TW, log_n, log_ns, last, io, s, batch)
#define H2B_PASSB(S_, C_, EL_, NT_, B_) return launch_pass<S_, C_, EL_, NT_, B_>(in, out, W, TW, log_n, log_ns, last, io, s, batch)
#define H2B_TILE10(EL_, NT_, B_)                      \
    switch (S) {                                      \
        case 5: H2B_PASSB(5, 32, EL_, NT_, B_);       \
        case 6: H2B_PASSB(6, 16, EL_, NT_, B_);       \
        case 7: H2B_PASSB(7, 8, EL_, NT_, B_);        \
        case 8: H2B_PASSB(8, 4, EL_, NT_, B_);        \
        case 9: H2B_PASSB(9, 2, EL_, NT_, B_);        \
        case 10: H2B_PASSB(10, 1, EL_, NT_, B_);      \
    }
    if (single) {
        switch (S) {
            case 1: H2B_PASS(1, 1, 1, 32);
            case 2: H2B_PASS(2, 1, 2, 32);
            case 3: H2B_PASS(3, 1, 3, 32);
            case 4: H2B_PASS(4, 1, 3, 32);
            case 5: H2B_PASS(5, 1, 3, 32);
            case 6: H2B_PASS(6, 1, 3, 32);
            case 7: H2B_PASS(7, 1, 3, 32);
            case 8: H2B_PASS(8, 1, 3, 32);
            case 9: H2B_PASS(9, 1, 3, 64);
            case 10: H2B_PASS(10, 1, 3, 128);
        }
    } else if (TL == 10 && EL == 3) {
        if (g_ntt_dense) { H2B_TILE10(3, 128, 5) } else { H2B_TILE10(3, 128, 4) }
    } else if (TL == 10 && EL == 2) {
        // measured alternatives (H2B_NTT_VARIANT): 1 = TMA bulk staging of the first round (plain transforms), 2 = shuffle exchanges
        const bool plain = !io.pro && io.n_in == (1u << log_n);
        if (g_ntt_variant == 1 && plain && S == 10) return launch_pass<10, 1, 2, 256, 3, 1>(in, out, W, TW, log_n, log_ns, last, io, s, batch);
        if (g_ntt_variant == 1 && plain && S == 8) return launch_pass<8, 4, 2, 256, 3, 1>(in, out, W, TW, log_n, log_ns, last, io, s, batch);
        if (g_ntt_variant == 2 && S == 10) return launch_pass<10, 1, 2, 256, 3, 2>(in, out, W, TW, log_n, log_ns, last, io, s, batch);
        if (g_ntt_dense == 2) { H2B_TILE10(2, 256, 4) } else if (g_ntt_dense) { H2B_TILE10(2, 256, 3) } else { H2B_TILE10(2, 256, 2) }
    } else if (TL == 10 && EL == 1) {
        H2B_TILE10(1, 512, 2)
    } else if (TL == 9 && EL == 2) {
        switch (S) {
            case 5: H2B_PASS(5, 16, 2, 128);
            case 6: H2B_PASS(6, 8, 2, 128);
            case 7: H2B_PASS(7, 4, 2, 128);
            case 8: H2B_PASS(8, 2, 2, 128);
            case 9: H2B_PASS(9, 1, 2, 128);
        }
    } else if (TL == 8 && EL == 1) {
        switch (S) {
            case 4: H2B_PASS(4, 16, 1, 128);
            case 5: H2B_PASS(5, 8, 1, 128);
            case 6: H2B_PASS(6, 4, 1, 128);
            case 7: H2B_PASS(7, 2, 1, 128);
            case 8: H2B_PASS(8, 1, 1, 128);
        }
    }
#undef H2B_TILE10
#undef H2B_PASSB
#undef H2B_PASS
    return fail(H2B_ERR_ARG, "ntt: unsupported radix / tile configuration");
}

uint32_t g_ntt_max_radix = 10;  // two passes up to 2^20, three up to 2^28 (H2B_NTT_MAX_RADIX)
int g_ntt_el_big = 2;           // elements per thread (log2) of the 2^10-element throughput tiles (H2B_NTT_EL_BIG): measured
                                // on B200 at k = 20 / 22 / 24: 4 per thread on 256 threads, 3 blocks per SM (24 warps) 0.200 / 0.874 /
                                // 3.53 ms; 8 per thread on 128 threads, 4 blocks (16 warps) 0.220 / 0.919 / 3.62; 2 per thread on 512
                                // threads (32 warps, but ten exchanges) 0.225 / 0.955 / 4.05
int g_ntt_tile = 0;             // forced tile size as log2, 8..10 (H2B_NTT_TILE); 0 = by size
int g_ntt_variant = 0;          // 0 product path, 1 TMA-staged first round, 2 shuffle exchanges (H2B_NTT_VARIANT; A/B only)
int g_ntt_dense = 1;            // registers held to more blocks per SM (H2B_NTT_DENSE): 0 = 4 (EL = 3) / 2 (EL = 2) blocks, 1 = 5 / 3
                                // (EL = 2: 80 registers, 24 warps per SM -- the default), 2 = 5 / 4 (64 registers, 32 warps, small spills)

// Transform `src` (n_in valid elements of a 2^log_n domain) into `dst`; `dst` may equal `src`.
// dst_full: `dst` holds 2^log_n elements and may carry intermediate passes; otherwise (truncated
// output) intermediates stay in library scratch and only the last pass touches `dst`.
// batch > 1: `batch` independent transforms, transform b at src + b * io.bin / dst + b * io.bout.
int ntt_run(const Fe *src, Fe *dst, uint32_t log_n, const uint64_t omega[4], NttIo io, cudaStream_t s,
            bool dst_full = true, uint32_t batch = 1) {
    if (log_n > 28) return fail(H2B_ERR_ARG, "ntt: log_n > 28 (Fr::S)");
    if (batch == 0) return H2B_OK;
    if (log_n == 0 && batch > 1) {
        for (uint32_t b = 0; b < batch; b++) {
            NttIo one = io;
            one.bin = one.bout = 0;
            TRY(ntt_run(src + (size_t)b * io.bin, dst + (size_t)b * io.bout, 0, omega, one, s, dst_full, 1));
        }
        return H2B_OK;
    }
    if (log_n == 0) {
        // length-1 transform is the identity; only the fused scalings remain (index 0: pro is a no-op)
        if (io.n_out == 0) return H2B_OK;
        if (src != dst) CU(cudaMemcpyAsync(dst, src, sizeof(Fe), cudaMemcpyDeviceToDevice, s));
        if (io.epi) {
            Fe *c;
            TRY(get_buf(BUF_MISC, sizeof(Fe), (void **)&c));
            CU(cudaMemcpyAsync(c, &io.epi_c[0], sizeof(Fe), cudaMemcpyHostToDevice, s));
            fr_scale_cyclic_kernel<<<1, 32, 0, s>>>(dst, 1, c, 1);
            LAUNCHED();
            CU(cudaStreamSynchronize(s));
        }
        return H2B_OK;
    }
    TwEntry *tw;
    TRY(get_twiddles(omega, log_n, s, &tw));
    uint32_t P, radix[4], EL = 3, TL = 10;
    bool single = log_n <= 10;
    if (single) {
        P = 1;
        radix[0] = log_n;
    } else {
        // 2^10-element tiles when that already gives every SM two of them; smaller tiles (more blocks) for transforms
        // that would leave SMs idle
        const uint64_t tiles10 = ((uint64_t)batch << log_n) >> 10;
        if (g_ntt_tile) TL = (uint32_t)g_ntt_tile;
        else if (tiles10 >= 2u * (uint32_t)g->sm_count) TL = 10;
        else if (tiles10 >= (uint32_t)g->sm_count) TL = 9;
        else TL = 8;
        auto passes = [&](uint32_t tl) {
            const uint32_t mr = std::min(g_ntt_max_radix, tl);
            return std::max(2u, (log_n + mr - 1) / mr);
        };
        while (TL < 10 && passes(TL) > passes(10)) TL++;  // never more passes than the throughput shape needs
        EL = TL == 10 ? (uint32_t)g_ntt_el_big : TL - 7;
        P = passes(TL);
        if (P > 4) return fail(H2B_ERR_ARG, "ntt: too many passes");
        const uint32_t base = log_n / P, rem = log_n % P;  // log_n >= 11: every radix >= 5
        for (uint32_t i = 0; i < P; i++) radix[i] = base + (i < rem ? 1 : 0);
    }
    Fe *tmp = nullptr, *tmp2 = dst;
    const uint32_t N = 1u << log_n;
    uint32_t tmp2_stride = io.bout;
    if (P > 1) TRY(get_buf(BUF_NTT_T, (size_t)batch * N * sizeof(Fe), (void **)&tmp));
    if (P > 2 && !dst_full) {
        TRY(get_buf(BUF_NTT_T2, (size_t)batch * N * sizeof(Fe), (void **)&tmp2));
        tmp2_stride = N;
    }
    const Fe *TW[4];
    for (uint32_t i = 0; i < P; i++) TRY(get_pass_twiddles(tw, log_n, radix[i], s, &TW[i]));
    time_begin(s);
    const Fe *cur = src;
    uint32_t log_ns = 0, cur_stride = io.bin;
    for (uint32_t i = 0; i < P; i++) {
        bool last = (i + 1 == P);
        Fe *to = last ? dst : ((i & 1) == 0 ? tmp : tmp2);
        NttIo pio = io;
        if (i != 0) { pio.pro = 0; pio.n_in = 1u << log_n; pio.bin = cur_stride; }
        if (!last) { pio.epi = 0; pio.n_out = 1u << log_n; pio.bout = (to == tmp) ? N : tmp2_stride; }
        TRY(dispatch_pass(radix[i], EL, TL, single, cur, to, tw->W, TW[i], log_n, log_ns, last, pio, s, batch));
        cur = to;
        cur_stride = pio.bout;
        log_ns += radix[i];
    }
    time_end(s);
    return H2B_OK;
}

NttIo io_plain(uint32_t log_n) {
    NttIo io;
    memset(&io, 0, sizeof io);
    io.n_in = io.n_out = 1u << log_n;
    return io;
}

// ------------------------------------------------------------------------------ domain constants
__device__ const uint32_t kRootOfUnityMont[8] = {0xb639feb8u, 0x9632c7c5u, 0x0d0ff299u, 0x985ce340u,
                                                 0x01b0ecd8u, 0xb2dd8800u, 0x6d98ce29u, 0x1d69070du};
// Fr::ZETA = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd in Montgomery form.  (SURVEY.md quotes the other
// primitive cube root, ZETA^2 = g_coset_inv; the reference's recorded coeff_to_extended calls decide: coefficient
// 1 is multiplied by this value.)
__device__ const uint32_t kZetaMont[8] = {0x4a0329b3u, 0x93e7cedeu, 0x7a96c167u, 0x7d4fdca7u,
                                          0xb19a750au, 0x8be4ba08u, 0xa5661c25u, 0x1cbd5653u};

struct DomainDev {
    Fe omega, omega_inv, ext_omega, ext_omega_inv, g_coset, g_coset_inv, ifft_div, ext_ifft_div;
    Fe t_eval[32];
    Fe ext_coset[3];
    uint32_t t_count;
};

// One block of 64 threads: thread 0 derives the forward constants, then 4 + n_t threads invert
// in parallel (EvaluationDomain::new batches these inversions; the values are the same).
__global__ void domain_new_kernel(uint32_t k, uint32_t ext_k, DomainDev *out) {
    __shared__ Fe to_inv[36];
    __shared__ uint32_t t_count;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        Fe w, zeta;
#pragma unroll
        for (int i = 0; i < 8; i++) { w.l[i] = kRootOfUnityMont[i]; zeta.l[i] = kZetaMont[i]; }
        for (uint32_t i = ext_k; i < 28; i++) w = Fr::sqr(w);
        Fe ext_w = w;
        for (uint32_t i = k; i < ext_k; i++) w = Fr::sqr(w);
        out->omega = w;
        out->ext_omega = ext_w;
        out->g_coset = zeta;
        out->g_coset_inv = Fr::sqr(zeta);
        uint64_t n = 1ull << k;
        Fe orig = Fr::pow_u64(zeta, n);
        Fe step = Fr::pow_u64(ext_w, n);
        Fe cur = orig;
        uint32_t cnt = 0;
        do {
            if (cnt < 32) to_inv[4 + cnt] = Fr::sub(cur, Fr::one());
            cnt++;
            cur = Fr::mul(cur, step);
        } while (!Fr::eq(cur, orig) && cnt < 64);
        t_count = cnt;
        to_inv[0] = w;
        to_inv[1] = ext_w;
        Fe two_k = Fr::zero();
        two_k.l[k >> 5] = 1u << (k & 31);
        to_inv[2] = Fr::to_mont(two_k);
        Fe two_e = Fr::zero();
        two_e.l[ext_k >> 5] = 1u << (ext_k & 31);
        to_inv[3] = Fr::to_mont(two_e);
    }
    __syncthreads();
    uint32_t cnt = t_count < 32 ? t_count : 32;
    if (tid < 4 + cnt) {
        Fe v = Fr::inv(to_inv[tid]);
        if (tid == 0) out->omega_inv = v;
        else if (tid == 1) out->ext_omega_inv = v;
        else if (tid == 2) out->ifft_div = v;
        else if (tid == 3) {
            out->ext_ifft_div = v;
            Fe zeta;
#pragma unroll
            for (int i = 0; i < 8; i++) zeta.l[i] = kZetaMont[i];
            out->ext_coset[0] = v;
            out->ext_coset[1] = Fr::mul(v, Fr::sqr(zeta));
            out->ext_coset[2] = Fr::mul(v, zeta);
        } else out->t_eval[tid - 4] = v;
    }
    if (tid == 0) out->t_count = t_count;
}

int domain_build(uint32_t j, uint32_t k, h2b_domain *out) {
    if (j < 2) return fail(H2B_ERR_ARG, "domain: j < 2");
    uint64_t qdeg = j - 1;
    uint32_t ext_k = k;
    if (k > 28) return fail(H2B_ERR_ARG, "domain: k > 28");
    while (((uint64_t)1 << ext_k) < ((uint64_t)1 << k) * qdeg) ext_k++;
    if (ext_k > 28) return fail(H2B_ERR_ARG, "domain: extended_k > Fr::S");
    if (ext_k - k > 5) return fail(H2B_ERR_ARG, "domain: extended_k - k > 5 not supported");
    uint64_t key = ((uint64_t)j << 32) | k;
    auto it = g->domains.find(key);
    if (it != g->domains.end()) {
        *out = it->second;
        return H2B_OK;
    }
    DomainDev *dd;
    TRY(get_buf(BUF_MISC, sizeof(DomainDev), (void **)&dd));
    domain_new_kernel<<<1, 64, 0, g->stream>>>(k, ext_k, dd);
    LAUNCHED();
    DomainDev h;
    CU(cudaMemcpyAsync(&h, dd, sizeof h, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    if (h.t_count != (1u << (ext_k - k)))  // domain.rs:101
        return fail(H2B_ERR_ARG, "domain: t_evaluations.len() != 1 << (extended_k - k)");
    h2b_domain d;
    memset(&d, 0, sizeof d);
    d.k = k; d.extended_k = ext_k; d.j = j; d.n_t = h.t_count;
    memcpy(d.omega, &h.omega, 32);
    memcpy(d.omega_inv, &h.omega_inv, 32);
    memcpy(d.extended_omega, &h.ext_omega, 32);
    memcpy(d.extended_omega_inv, &h.ext_omega_inv, 32);
    memcpy(d.g_coset, &h.g_coset, 32);
    memcpy(d.g_coset_inv, &h.g_coset_inv, 32);
    memcpy(d.ifft_divisor, &h.ifft_div, 32);
    memcpy(d.extended_ifft_divisor, &h.ext_ifft_div, 32);
    memcpy(d.t_evaluations, h.t_eval, 32 * d.n_t);
    memcpy(d.extended_ifft_coset, h.ext_coset, 96);
    g->domains[key] = d;
    *out = d;
    return H2B_OK;
}

int check_domain(const h2b_domain *d) {
    if (!d) return fail(H2B_ERR_ARG, "domain is null");
    if (d->k > d->extended_k || d->extended_k > 28 || d->j < 2) return fail(H2B_ERR_ARG, "domain is malformed");
    return H2B_OK;
}

// batch columns lie one after the other: a + b * 2^k
int dev_lagrange_to_coeff(const h2b_domain *d, Fe *a, cudaStream_t s, uint32_t batch = 1) {
    NttIo io = io_plain(d->k);
    io.epi = 1;
    for (int i = 0; i < 3; i++) memcpy(&io.epi_c[i], d->ifft_divisor, 32);
    io.bin = io.bout = 1u << d->k;
    return ntt_run(a, a, d->k, d->omega_inv, io, s, true, batch);
}
// batch: in + b * 2^k -> out + b * 2^extended_k
int dev_coeff_to_extended(const h2b_domain *d, const Fe *in, Fe *out, cudaStream_t s, uint32_t batch = 1) {
    NttIo io = io_plain(d->extended_k);
    io.n_in = 1u << d->k;
    io.pro = 1;
    memcpy(&io.pro_c[1], d->g_coset, 32);      // i % 3 == 1 -> zeta
    memcpy(&io.pro_c[2], d->g_coset_inv, 32);  // i % 3 == 2 -> zeta^2
    io.bin = 1u << d->k;
    io.bout = 1u << d->extended_k;
    return ntt_run(in, out, d->extended_k, d->extended_omega, io, s, true, batch);
}
int dev_extended_to_coeff(const h2b_domain *d, const Fe *in, Fe *out, cudaStream_t s) {
    // ifft with extended_omega_inv; the last pass multiplies by 1/2^ext_k * {1, zeta^2, zeta}[i % 3]
    // and stores only the first n * (j - 1) coefficients (the truncation of domain.rs:~325).
    size_t en = (size_t)1 << d->extended_k;
    size_t keep = (size_t)(d->j - 1) << d->k;
    if (keep > en) return fail(H2B_ERR_ARG, "extended_to_coeff: n*(j-1) > extended length");
    NttIo io = io_plain(d->extended_k);
    io.epi = 1;
    memcpy(io.epi_c, d->extended_ifft_coset, 96);
    io.n_out = (uint32_t)keep;
    return ntt_run(in, out, d->extended_k, d->extended_omega_inv, io, s, /*dst_full=*/false);
}

// G1Affine::to_bytes of halo2curves 0.3.2 (the 32-byte form the transcript writes for every commitment):
// x as a canonical little-endian integer, bit 6 of byte 31 = y mod 2 (confirmed on the reference's own proof
// bytes, tests/test_wasm_golden.py); the identity is written as 32 zero bytes (no identity occurs in those
// proofs, so that case rests on the upstream source as remembered, not on a recorded byte).
__global__ void g1_to_bytes_kernel(const Projective *__restrict__ pts, uint32_t m, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const Fe X = load_fe(&pts[i].x), Y = load_fe(&pts[i].y), Z = load_fe(&pts[i].z);
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!Fq::is_zero(Z)) {
        const Fe zi = Fq::inv(Z);
        const Fe x = Fq::from_mont(Fq::mul(X, zi)), y = Fq::from_mont(Fq::mul(Y, zi));
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = x.l[k];
        w[7] |= (y.l[0] & 1u) << 30;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[(size_t)i * 8 + k] = w[k];
}

// ------------------------------------------------------------------------------ test kernels
template <class F>
__global__ void test_field_kernel(int op, const Fe *a, const Fe *b, Fe *o, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fe x = load_fe(&a[i]);
    Fe y = b ? load_fe(&b[i]) : F::zero();
    Fe r;
    switch (op) {
        case 0: r = F::mul(x, y); break;
        case 1: r = F::add(x, y); break;
        case 2: r = F::sub(x, y); break;
        case 3: r = F::mul_portable(x, y); break;
        case 4: r = F::inv(x); break;
        case 6: r = F::sqr(x); break;
        case 7: r = F::mul2_add(x, y, F::add(x, y), F::sub(x, y)); break;   // xy + x^2 - y^2
        case 8: r = F::mul2_sub(x, y, y, x); break;                         // 0, through neg()
        case 9: r = F::reduce_once(F::mul_lazy(x, y)); break;               // x < 4N (raw limbs), y < N
        case 10: r = F::reduce_once(F::reduce_2n(F::sub_2n(x, y))); break;  // x, y < 2N (raw limbs)
        case 11: r = F::reduce_once(F::add_2n(x, y)); break;                // x, y < 2N (raw limbs)
        default: r = F::from_mont(x); break;
    }
    store_fe(&o[i], r);
}
__global__ void test_g1_add_kernel(const Affine *a, const Affine *b, Projective *o, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine p = load_affine(&a[i]), q = load_affine(&b[i]);
    XYZZ acc = xyzz_from_affine(p);
    if (!affine_is_identity(q)) xyzz_madd(acc, q);
    Projective j = xyzz_to_projective(acc);
    store_fe(&o[i].x, j.x);
    store_fe(&o[i].y, j.y);
    store_fe(&o[i].z, j.z);
}

// Integer-pipe microbenchmarks: 8 independent chains per thread, fully unrolled bodies.
__global__ void imad_bench_kernel(uint32_t *sink, uint32_t iters, uint32_t seed) {
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11,
             a6 = a0 * 13, a7 = a0 * 15;
    uint32_t b = seed | 1, c = seed ^ 0x9e3779b9u;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a4) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a5) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a6) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a7) : "r"(b), "r"(c));
        }
    }
    uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678u) sink[0] = r;
}
__global__ void imad_wide_bench_kernel(uint64_t *sink, uint32_t iters, uint32_t seed) {
    uint64_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 9, a5 = a0 * 11,
             a6 = a0 * 13, a7 = a0 * 15;
    uint32_t b = seed | 1;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"((uint32_t)a0), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"((uint32_t)a1), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"((uint32_t)a2), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"((uint32_t)a3), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"((uint32_t)a4), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"((uint32_t)a5), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"((uint32_t)a6), "r"(b));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"((uint32_t)a7), "r"(b));
        }
    }
    uint64_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x12345678u) sink[0] = r;
}

// host staging helpers
int stage_in(BufId id, const void *host, size_t bytes, void **dev, bool round_trip = false) {
    TRY(get_buf(id, bytes, dev));
    return copy_in(*dev, host, bytes, g->stream, round_trip);
}

}  // namespace

// ============================================================================== C ABI
extern "C" {

uint32_t h2b_abi_version(void) { return 2; }
const char *h2b_last_error(void) { return g_err.c_str(); }

// One context per device: stream, copy stream, workspace, caches, tunables from the environment.
static int ctx_create(int device, Ctx **out) {
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(H2B_ERR_CUDA, "h2b_init: kernels are built for sm_100a only");
    Ctx *c = new Ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->copy_fence, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->last_done, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return fail(H2B_ERR_CUDA, "h2b_init: stream/event creation failed", cudaGetLastError());
    }
    {
        int threads = (int)std::min(6u, std::max(2u, std::thread::hardware_concurrency() / 2));
        const char *ct = getenv("H2B_COPY_THREADS");  // 0: leave pageable copies to the driver
        if (ct) threads = atoi(ct);
        c->copier = new HostCopier(threads);
    }
    const char *sw = getenv("H2B_SRS_WINDOW");
    if (sw) {
        int v = atoi(sw);
        if (v >= 2 && v <= 24) c->srs_window = (uint32_t)v;
    }
    const char *ts = getenv("H2B_SRS_TABLE_STRIDE");
    if (ts && atoi(ts) >= 0 && atoi(ts) <= 8) c->srs_table_stride = (uint32_t)atoi(ts);
    const char *rl = getenv("H2B_REDUCE_LGRP");
    if (rl) c->reduce_lgrp = (uint32_t)atoi(rl);
    const char *rt = getenv("H2B_REDUCE_TREE");
    if (rt) c->reduce_tree = (uint32_t)atoi(rt);
    const char *ro = getenv("H2B_REDUCE_LONE");
    if (ro) c->reduce_lone = (uint32_t)atoi(ro);
    const char *ms = getenv("H2B_MIN_SLICE");
    if (ms && atoi(ms) >= 1) c->min_slice = (uint32_t)atoi(ms);
    const char *mw = getenv("H2B_MIN_WAVES");
    if (mw && atoi(mw) >= 1) c->min_waves = (uint32_t)atoi(mw);
    const char *rq = getenv("H2B_REDUCE_Q");
    if (rq) c->reduce_q = atoi(rq);
    const char *er = getenv("H2B_E2E_RATIO");
    if (er) c->e2e_ratio = atof(er);
    const char *sp = getenv("H2B_SRS_PRECOMPUTE");
    if (sp) c->srs_precompute = atoi(sp);
    const char *ec = getenv("H2B_E2E_CHUNKS");
    if (ec) {
        int v = atoi(ec);
        if (v >= 1 && v <= 64) { c->e2e_chunks = (uint32_t)v; c->e2e_auto = false; }
    }
    c->worker.start();
    *out = c;
    return H2B_OK;
}
static void ctx_destroy(Ctx *c) {
    c->worker.stop();
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < BUF_COUNT; i++)
        if (c->buf[i]) cudaFree(c->buf[i]);
    for (auto &kv : c->twiddles) {
        cudaFree(kv.second.W);
        for (Fe *t : kv.second.pass)
            if (t) cudaFree(t);
    }
    for (auto e : c->tev0) cudaEventDestroy(e);
    for (auto e : c->tev1) cudaEventDestroy(e);
    for (auto e : c->chunk_events) cudaEventDestroy(e);
    delete c->copier;
    cudaEventDestroy(c->copy_fence);
    cudaStreamDestroy(c->copy_stream);
    cudaEventDestroy(c->last_done);
    cudaStreamDestroy(c->stream);
    delete c;
}

static int init_locked(const int *devices, int count) {
    if (!devices || count < 1 || count > 64) return fail(H2B_ERR_ARG, "h2b_init: bad device list");
    if (g_primary) {
        bool same = (int)g_all.size() == count;
        for (int i = 0; same && i < count; i++) same = g_all[i]->device == devices[i];
        if (same) return H2B_OK;
        return fail(H2B_ERR_STATE, "h2b_init: already initialised on other devices");
    }
    int have = 0;
    cudaError_t e = cudaGetDeviceCount(&have);
    if (e != cudaSuccess || have == 0) {
        (void)cudaGetLastError();
        return fail(H2B_ERR_CUDA, "h2b_init: no CUDA device (this library has no CPU fallback)", e);
    }
    for (int i = 0; i < count; i++) {
        if (devices[i] < 0 || devices[i] >= have) return fail(H2B_ERR_ARG, "h2b_init: bad device index");
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) return fail(H2B_ERR_ARG, "h2b_init: device listed twice");
    }
    const char *sm = getenv("H2B_SHARD_MIN_LOG");
    if (sm && atoi(sm) >= 8 && atoi(sm) <= 30) g_shard_min_n = (size_t)1 << atoi(sm);
    const char *el = getenv("H2B_NTT_EL_BIG");
    if (el && atoi(el) >= 1 && atoi(el) <= 3) g_ntt_el_big = atoi(el);
    const char *nv = getenv("H2B_NTT_VARIANT");
    if (nv) g_ntt_variant = atoi(nv);
    const char *dn = getenv("H2B_NTT_DENSE");
    if (dn) g_ntt_dense = atoi(dn);
    const char *tl = getenv("H2B_NTT_TILE");
    if (tl && atoi(tl) >= 8 && atoi(tl) <= 10) g_ntt_tile = atoi(tl);
    const char *mr = getenv("H2B_NTT_MAX_RADIX");
    if (mr) {
        int v = atoi(mr);
        if (v >= 5 && v <= 10) g_ntt_max_radix = (uint32_t)v;
    }
    std::vector<Ctx *> made;
    for (int i = 0; i < count; i++) {
        Ctx *c = nullptr;
        int rc = ctx_create(devices[i], &c);
        if (rc != H2B_OK) {
            for (Ctx *m : made) ctx_destroy(m);
            return rc;
        }
        made.push_back(c);
    }
    // the partial sums of a sharded commit travel device-to-device (96 bytes each): enable direct peer access where
    // the topology offers it (NVLink / NVSwitch on an HGX board); without it the peer copies are staged by the driver
    for (int i = 0; i < count && count > 1; i++) {
        cudaSetDevice(devices[i]);
        for (int j = 0; j < count; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(devices[j], 0);
                if (pe != cudaSuccess) (void)cudaGetLastError();  // already enabled (e.g. by the host framework)
            }
        }
    }
    // what the host delivers to every device when all of them copy at once sizes the copy pieces of a sharded commit
    if (count > 1) {
        const size_t probe = (size_t)32 << 20;
        void *hbuf = nullptr;
        std::vector<void *> dbuf(count, nullptr);
        bool ok = cudaHostAlloc(&hbuf, probe, cudaHostAllocDefault) == cudaSuccess;
        for (int i = 0; ok && i < count; i++) {
            cudaSetDevice(devices[i]);
            ok = cudaMalloc(&dbuf[i], probe) == cudaSuccess;
        }
        if (ok) {
            double best = 0;
            for (int rep = 0; rep < 3; rep++) {
                for (int i = 0; i < count; i++) {
                    cudaSetDevice(devices[i]);
                    cudaStreamSynchronize(made[i]->copy_stream);
                }
                const auto t0 = std::chrono::steady_clock::now();
                for (int i = 0; i < count; i++) {
                    cudaSetDevice(devices[i]);
                    cudaMemcpyAsync(dbuf[i], hbuf, probe, cudaMemcpyHostToDevice, made[i]->copy_stream);
                }
                for (int i = 0; i < count; i++) {
                    cudaSetDevice(devices[i]);
                    cudaStreamSynchronize(made[i]->copy_stream);
                }
                const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (rep) best = std::max(best, (double)probe / sec / 1e9);  // per device, all copying at once
            }
            if (best > 1.0)
                for (Ctx *c : made) c->h2d_gbs = std::min(55.0, best);
        }
        (void)cudaGetLastError();
        for (int i = 0; i < count; i++)
            if (dbuf[i]) {
                cudaSetDevice(devices[i]);
                cudaFree(dbuf[i]);
            }
        if (hbuf) cudaFreeHost(hbuf);
    }
    cudaSetDevice(devices[0]);
    g_all = made;
    g_primary = made[0];
    g = g_primary;
    return H2B_OK;
}

int h2b_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(&device, 1);
}
int h2b_init_devices(const int *devices, int count) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(devices, count);
}
int h2b_device_count(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    return (int)g_all.size();
}

static void srs_free(Srs &sr);
void h2b_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_primary) return;
    for (auto &kv : g_srs)
        for (Srs &pt : kv.second.parts) srs_free(pt);
    g_srs.clear();
    for (Ctx *c : g_all) ctx_destroy(c);
    g_all.clear();
    g_primary = nullptr;
    g = nullptr;
}

int h2b_set_msm_window(uint32_t c) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (c != 0 && (c < 2 || c > 22)) return fail(H2B_ERR_ARG, "msm window must be 0 or in [2, 22]");
    for (Ctx *x : g_all) x->msm_window = c;
    return H2B_OK;
}
int h2b_set_srs_precompute(int enabled, uint32_t c) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (c != 0 && (c < 2 || c > 24)) return fail(H2B_ERR_ARG, "srs window must be 0 or in [2, 24]");
    for (Ctx *x : g_all) {
        x->srs_precompute = enabled;  // 0: none, 1: automatic (bucket-free table up to 2^14 points), 2: window table only
        x->srs_window = c;
    }
    return H2B_OK;
}
int h2b_set_h2d_bandwidth(double gbs) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (gbs < 0 || gbs > 1000) return fail(H2B_ERR_ARG, "h2d bandwidth must be 0 (default) or a positive GB/s figure");
    for (Ctx *x : g_all) x->h2d_gbs = gbs > 0 ? std::min(55.0, gbs) : 55.0;
    return H2B_OK;
}
int h2b_set_srs_table_stride(uint32_t t) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (t > 8) return fail(H2B_ERR_ARG, "srs table stride must be 0 (automatic) or in [1, 8]");
    for (Ctx *x : g_all) x->srs_table_stride = t;
    return H2B_OK;
}
int h2b_set_e2e_chunking(uint32_t chunks, size_t min_n) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (chunks > 64) return fail(H2B_ERR_ARG, "e2e chunks must be 0 (by size) or in [1, 64]");
    for (Ctx *x : g_all) {
        x->e2e_chunks = chunks ? chunks : 4;
        x->e2e_auto = chunks == 0;
        x->e2e_min_n = min_n;
    }
    return H2B_OK;
}
uint64_t h2b_kernel_launches(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    uint64_t sum = 0;
    for (Ctx *x : g_all) sum += x->launches;
    return sum;
}
int h2b_set_kernel_timing(int enabled) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    for (Ctx *x : g_all) {
        x->timing = enabled;
        x->tev_used = 0;
    }
    return H2B_OK;
}
int h2b_kernel_time_collect(double *total_ms, uint32_t *calls) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    double sum = 0;
    uint32_t n = 0;
    for (Ctx *x : g_all) {  // every device's pairs: a sharded commit contributes one accumulation launch per device
        CU(cudaSetDevice(x->device));
        for (uint32_t i = 0; i < x->tev_used; i++) {
            CU(cudaEventSynchronize(x->tev1[i]));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, x->tev0[i], x->tev1[i]));
            sum += ms;
        }
        n += x->tev_used;
        x->tev_used = 0;
    }
    CU(cudaSetDevice(g->device));
    if (total_ms) *total_ms = sum;
    if (calls) *calls = n;
    return H2B_OK;
}

// ---- MSM
int h2b_dev_msm(const void *d_coeffs, const void *d_bases, size_t n, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!d_out || (n && (!d_coeffs || !d_bases))) return fail(H2B_ERR_ARG, "msm: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return msm_run((const Fe *)d_coeffs, (const Affine *)d_bases, n, (Projective *)d_out, s);
}

int commit_many_device(const Srs &sr, const Fe *d_scalars, size_t n, size_t m, Projective *d_out, cudaStream_t s);

// Bucket-free commit of `cols` columns against a small SRS (msm_comb.cuh): digits -> table indices, slice sums,
// a binary tree over the slice sums.
int comb_commit(const Srs &sr, const Fe *d_scalars, size_t n, size_t cols, Projective *d_out, cudaStream_t s) {
    MsmCfg cfg{};
    cfg.n = (uint32_t)n;
    cfg.cols = (uint32_t)cols;
    cfg.c = sr.comb_c;
    cfg.windows = sr.comb_w;
    cfg.bpw = 1u << (sr.comb_c - 1);
    cfg.stride = (uint32_t)sr.n;
    for (uint32_t w = 0; w + 1 < cfg.windows; w++) {
        const uint32_t bit = cfg.c * w + cfg.c - 1;
        cfg.half[bit >> 5] |= 1u << (bit & 31);
    }
    const size_t per_col = n * cfg.windows, total = per_col * cols;
    uint32_t *entries;
    TRY(get_buf(BUF_SORTED, total * 4, (void **)&entries));
    msm_comb_digits_kernel<<<(uint32_t)((n * cols + 255) / 256), 256, 0, s>>>(d_scalars, cfg, entries);
    LAUNCHED();
    // ~2^17 slice sums in flight: short enough chains, few enough tree levels
    size_t L = (total + (1u << 17) - 1) >> 17;
    if (L < 2) L = 2;
    if (L > 64) L = 64;
    const size_t slices = (per_col + L - 1) / L;
    XYZZ *pa, *pb;
    TRY(get_buf(BUF_HEAD, cols * slices * sizeof(XYZZ), (void **)&pa));
    time_begin(s);
    msm_comb_sum_kernel<<<dim3((uint32_t)((slices + 127) / 128), (uint32_t)cols), 128, 0, s>>>(
        sr.comb, entries, (uint32_t)per_col, (uint32_t)L, (uint32_t)slices, pa);
    LAUNCHED();
    time_end(s);
    size_t count = slices;
    TRY(get_buf(BUF_TAIL, cols * ((count + 255) / 256 + 1) * sizeof(XYZZ), (void **)&pb));
    XYZZ *src = pa, *dst = pb;
    while (count > 1) {
        uint32_t nt = 256, pt = 1;
        size_t blocks;
        if (count <= 256) {
            nt = 32;
            while (nt < count) nt <<= 1;
            blocks = 1;
        } else {
            pt = (uint32_t)((count + 65535) / 65536);  // at most 256 blocks remain after this level group
            blocks = (count + (size_t)nt * pt - 1) / ((size_t)nt * pt);
        }
        msm_comb_tree_kernel<<<dim3((uint32_t)blocks, (uint32_t)cols), nt, nt * sizeof(XYZZ), s>>>(src, (uint32_t)count, pt, dst);
        LAUNCHED();
        count = blocks;
        std::swap(src, dst);
    }
    msm_batch_out_kernel<<<(uint32_t)((cols + 31) / 32), 32, 0, s>>>(src, (uint32_t)cols, 1, 0, d_out);
    LAUNCHED();
    return H2B_OK;
}

// m polynomials of n scalars each (device, one after the other) against bases[0..n] of one SRS.  With a window
// table every column gets its own bucket set inside ONE pass of the kernels (columns take the place of
// windows in the reduction), so the fixed latency of a small MSM is paid once per batch, not per column.
int commit_many_device(const Srs &sr, const Fe *d_scalars, size_t n, size_t m, Projective *d_out, cudaStream_t s) {
    if (m == 0) return H2B_OK;
    if (n == 0) {
        for (size_t q = 0; q < m; q++) TRY(msm_identity_out(d_out + q, s));
        return H2B_OK;
    }
    if (sr.comb) {
        const size_t per_col = n * sr.comb_w;
        size_t step = m;
        while (step > 1 && step * per_col > ((size_t)1 << 27)) step = (step + 1) / 2;
        for (size_t q0 = 0; q0 < m; q0 += step)
            TRY(comb_commit(sr, d_scalars + q0 * n, n, std::min(step, m - q0), d_out + q0, s));
        return H2B_OK;
    }
    if (!sr.table) {  // no table: one MSM per column
        for (size_t q = 0; q < m; q++) TRY(msm_run(d_scalars + q * n, sr.d, n, d_out + q, s));
        return H2B_OK;
    }
    // bound one pass: sorted entries < 2^28 and at most 1 GiB of buckets
    const uint64_t per_col_entries = (uint64_t)n * sr.windows, per_col_buckets = (uint64_t)sr.tstride << (sr.c - 1);
    size_t step = m;
    while (step > 1 && (step * per_col_entries > (1ull << 28) || step * per_col_buckets * sizeof(XYZZ) > (1ull << 30)))
        step = (step + 1) / 2;
    for (size_t q0 = 0; q0 < m; q0 += step) {
        const size_t cols = std::min(step, m - q0);
        MsmRun run;
        TRY(msm_begin(n, &run, s, &sr, (uint32_t)cols));
        TRY(msm_chunk(run, d_scalars + q0 * n, sr.table, n, s, 0));
        TRY(msm_finish(run, d_out + q0, s));
    }
    return H2B_OK;
}

// ---- several devices ---------------------------------------------------------------------------------------
// fn(i) runs for devs[i] on that context's worker thread with `g` bound to it (so every internal function above
// works unchanged on any device) while the API thread, which holds g_mu, waits.  First error wins.
int on_devices(const std::vector<int> &devs, const std::function<int(size_t)> &fn) {
    Ctx *self = g;
    std::vector<int> rcs(devs.size(), H2B_OK);
    std::vector<std::string> errs(devs.size());
    auto body = [&](size_t i) {
        g = g_all[devs[i]];
        if (cudaSetDevice(g->device) != cudaSuccess) {
            rcs[i] = H2B_ERR_CUDA;
            errs[i] = "cudaSetDevice";
            return;
        }
        rcs[i] = fn(i);
        if (rcs[i] != H2B_OK) errs[i] = g_err;
    };
    if (devs.size() == 1 && g_all[devs[0]] == self) {
        body(0);
    } else {
        for (size_t i = 0; i < devs.size(); i++) g_all[devs[i]]->worker.submit([&body, i] { body(i); });
        for (size_t i = 0; i < devs.size(); i++) g_all[devs[i]]->worker.wait();
    }
    g = self;
    cudaSetDevice(self->device);
    for (size_t i = 0; i < devs.size(); i++)
        if (rcs[i] != H2B_OK) {
            g_err = errs[i];
            return rcs[i];
        }
    return H2B_OK;
}

// out[q] = sum over p of pts[p * cols + q]: the fold of per-device partial commitments, column by column
// (arithmetic.rs:~176, `results.iter().fold(identity, |a, b| a + b)` with a device in place of a thread).
__global__ void g1_fold_cols_kernel(const Projective *__restrict__ pts, uint32_t parts, uint32_t cols, Projective *__restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cols) return;
    XYZZ acc = xyzz_identity();
    for (uint32_t p = 0; p < parts; p++) {
        Projective v;
        v.x = load_fe(&pts[(size_t)p * cols + q].x);
        v.y = load_fe(&pts[(size_t)p * cols + q].y);
        v.z = load_fe(&pts[(size_t)p * cols + q].z);
        XYZZ t = projective_to_xyzz(v);
        xyzz_add(acc, t);
    }
    const Projective r = xyzz_to_projective(acc);
    store_fe(&out[q].x, r.x);
    store_fe(&out[q].y, r.y);
    store_fe(&out[q].z, r.z);
}

// The polynomials of a commit: host columns, or columns one after the other in the PRIMARY device's memory.
struct ColumnSrc {
    const uint64_t *const *host = nullptr;
    const Fe *dev = nullptr;
};

// One device's share of a commit: columns [q0, q0 + m) restricted to points [off, off + n) of polynomials of
// length `n_poly`, against `part`; the m results land in d_out (device memory of the calling context).
int commit_part(const Srs &part, const ColumnSrc &src, size_t n_poly, size_t off, size_t n, size_t q0, size_t m,
                Projective *d_out, cudaStream_t s, int src_device) {
    if (src.host && m == 1 && !part.comb) {  // large single commits: the copy is pipelined against the accumulation
        (void)s;
        return msm_run_pipelined(src.host[q0] + 4 * off, nullptr, &part, n, d_out);
    }
    Fe *ds;
    TRY(get_buf(BUF_SCALARS, m * n * sizeof(Fe), (void **)&ds));
    if (src.host) {
        std::vector<HostCopier::Seg> segs;
        for (size_t q = 0; q < m; q++)
            segs.push_back({ds + q * n, const_cast<uint64_t *>(src.host[q0 + q] + 4 * off), n * sizeof(Fe)});
        cudaError_t ce = g->copier->h2d(segs, s);
        if (ce != cudaSuccess) return fail(H2B_ERR_CUDA, "host-to-device copy", ce);
    } else if (src_device == g->device && (m == 1 || (off == 0 && n == n_poly))) {
        ds = const_cast<Fe *>(src.dev + q0 * n_poly + off);  // already here, contiguous
    } else {
        for (size_t q = 0; q < m; q++) {
            const Fe *from = src.dev + (q0 + q) * n_poly + off;
            if (src_device == g->device) CU(cudaMemcpyAsync(ds + q * n, from, n * sizeof(Fe), cudaMemcpyDeviceToDevice, s));
            else CU(cudaMemcpyPeerAsync(ds + q * n, g->device, from, src_device, n * sizeof(Fe), s));
        }
    }
    if (m == 1 && !part.comb) {
        MsmRun run;
        TRY(msm_begin(n, &run, s, &part));
        TRY(msm_chunk(run, ds, part.table ? part.table : part.d, n, s, 0));
        return msm_finish(run, d_out, s);
    }
    return commit_many_device(part, ds, n, m, d_out, s);
}

// ParamsKZG::commit / commit_lagrange of m polynomials of n scalars against a registered set.  Results go to the
// host (h_out, m x 12 u64) or to the primary device (d_out).  `s` is the stream of the primary device the call is
// ordered on (the caller's for the _dev entry points).
int commit_set(const SrsSet &set, const ColumnSrc &src, size_t n, size_t m, uint64_t *h_out, Projective *d_out, cudaStream_t s) {
    if (m == 0) return H2B_OK;
    Ctx *prim = g;
    Scope sc;
    TRY(sc.begin(s));
    Projective *dres = d_out;
    if (!dres) TRY(get_buf(BUF_OUT, m * sizeof(Projective), (void **)&dres));
    if (n == 0) {
        for (size_t q = 0; q < m; q++) TRY(msm_identity_out(dres + q, s));
        if (h_out) TRY(copy_out(h_out, dres, m * sizeof(Projective), s));
        return H2B_OK;
    }
    const size_t nparts = set.parts.size();
    const bool deal = nparts > 1 && set.replicated && src.host && m >= 2;
    const bool shard = nparts > 1 && !set.replicated;
    if (!deal && !shard) {  // one device does it all
        TRY(commit_part(set.parts[0], src, n, 0, n, 0, m, dres, s, prim->device));
        if (h_out) TRY(copy_out(h_out, dres, m * sizeof(Projective), s));
        return H2B_OK;
    }
    // the work list: (part, columns [q0, q0 + mq), points [off, off + np))
    struct Job { size_t part, q0, mq, off, np; };
    std::vector<Job> jobs;
    if (deal) {  // whole columns, contiguous blocks per device
        const size_t use = std::min(nparts, m);
        for (size_t i = 0, q0 = 0; i < use; i++) {
            const size_t mq = m / use + (i < m % use ? 1 : 0);
            jobs.push_back({i, q0, mq, 0, n});
            q0 += mq;
        }
    } else {
        for (size_t i = 0; i < nparts; i++) {
            const Srs &pt = set.parts[i];
            if (pt.off >= n) break;
            jobs.push_back({i, 0, m, pt.off, std::min(pt.n, n - pt.off)});
        }
    }
    // partial results of job j arrive at gather + j * m (sharded) / straight at their columns (dealt)
    Projective *gather = dres;
    if (shard) TRY(get_buf(BUF_GATHER, jobs.size() * m * sizeof(Projective), (void **)&gather));
    cudaEvent_t ready = nullptr;
    if (src.dev) {  // the producers of the device columns run on `s`: the other devices' peer copies wait for them
        ready = prim->copy_fence;
        CU(cudaEventRecord(ready, s));
    }
    std::vector<int> devs;
    for (const Job &j : jobs) devs.push_back(set.parts[j.part].dev);
    TRY(on_devices(devs, [&](size_t i) -> int {
        const Job &j = jobs[i];
        const Srs &pt = set.parts[j.part];
        const bool here = g == prim;
        cudaStream_t ws = here ? s : g->stream;
        Scope wsc;
        if (!here) TRY(wsc.begin(ws));
        if (ready && !here) CU(cudaStreamWaitEvent(ws, ready, 0));
        Projective *target = gather + (shard ? i * m : j.q0);
        Projective *local = target;
        if (!here) TRY(get_buf(BUF_OUT, j.mq * sizeof(Projective), (void **)&local));
        TRY(commit_part(pt, src, n, j.off, j.np, j.q0, j.mq, local, ws, prim->device));
        if (!here) CU(cudaMemcpyPeerAsync(target, prim->device, local, g->device, j.mq * sizeof(Projective), ws));
        if (!here) CU(cudaStreamSynchronize(ws));  // the primary folds what has arrived
        return H2B_OK;
    }));
    if (shard) {
        g1_fold_cols_kernel<<<(uint32_t)((m + 31) / 32), 32, 0, s>>>(gather, (uint32_t)jobs.size(), (uint32_t)m, dres);
        LAUNCHED();
    }
    if (h_out) TRY(copy_out(h_out, dres, m * sizeof(Projective), s));
    return H2B_OK;
}

int find_set(uint64_t handle, size_t n, const char *what, const SrsSet **out) {
    auto it = g_srs.find(handle);
    if (it == g_srs.end()) return fail(H2B_ERR_STATE, "unknown SRS handle");
    if (n > it->second.n) return fail(H2B_ERR_ARG, what);  // commitment.rs:319/:363 assert!(bases.len() >= size)
    *out = &it->second;
    return H2B_OK;
}

int h2b_dev_commit(uint64_t srs, const void *d_coeffs, size_t n, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    const SrsSet *set;
    TRY(find_set(srs, n, "commit: bases.len() < size", &set));
    if (!d_out || (n && !d_coeffs)) return fail(H2B_ERR_ARG, "commit: null pointer");
    CU(cudaSetDevice(g->device));
    ColumnSrc src;
    src.dev = (const Fe *)d_coeffs;
    return commit_set(*set, src, n, 1, nullptr, (Projective *)d_out, stream ? (cudaStream_t)stream : g->stream);
}
int h2b_dev_commit_many(uint64_t srs, const void *d_coeffs, size_t n, size_t m, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    const SrsSet *set;
    TRY(find_set(srs, n, "commit_many: bases.len() < size", &set));
    if (m && (!d_out || (n && !d_coeffs))) return fail(H2B_ERR_ARG, "commit_many: null pointer");
    CU(cudaSetDevice(g->device));
    ColumnSrc src;
    src.dev = (const Fe *)d_coeffs;
    return commit_set(*set, src, n, m, nullptr, (Projective *)d_out, stream ? (cudaStream_t)stream : g->stream);
}
int h2b_commit_many(uint64_t srs, const uint64_t *const *polys, size_t n, size_t m, uint64_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    const SrsSet *set;
    TRY(find_set(srs, n, "commit_many: bases.len() < size", &set));
    if (m == 0) return H2B_OK;
    if (!out || !polys) return fail(H2B_ERR_ARG, "commit_many: null pointer");
    for (size_t q = 0; q < m && n; q++)
        if (!polys[q]) return fail(H2B_ERR_ARG, "commit_many: null column");
    CU(cudaSetDevice(g->device));
    ColumnSrc src;
    src.host = polys;
    return commit_set(*set, src, n, m, out, nullptr, g->stream);
}
int h2b_commit(uint64_t srs, const uint64_t *scalars, size_t n, uint64_t out[12]) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    const SrsSet *set;
    TRY(find_set(srs, n, "commit: bases.len() < size", &set));
    if (!out || (n && !scalars)) return fail(H2B_ERR_ARG, "commit: null pointer");
    CU(cudaSetDevice(g->device));
    ColumnSrc src;
    src.host = &scalars;
    return commit_set(*set, src, n, 1, out, nullptr, g->stream);
}

int h2b_best_multiexp(const uint64_t *coeffs, const uint64_t *bases, size_t n, uint64_t out[12]) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!out || (n && (!coeffs || !bases))) return fail(H2B_ERR_ARG, "best_multiexp: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *dout;
    TRY(get_buf(BUF_OUT, 96, &dout));
    TRY(msm_run_pipelined(coeffs, bases, nullptr, n, (Projective *)dout));
    TRY(copy_out(out, dout, 96, g->stream));
    return H2B_OK;
}

static void srs_free(Srs &sr) {
    cudaSetDevice(g_all[sr.dev]->device);
    cudaDeviceSynchronize();  // commits on caller streams may still read the bases / tables
    if (sr.d) cudaFree(sr.d);
    if (sr.table) cudaFree(sr.table);
    if (sr.comb) cudaFree(sr.comb);
    sr.d = sr.table = sr.comb = nullptr;
}

// Builds one share on the calling thread's context: points [off, off + n) of the source array (host memory, or
// memory of device `src_device`) go to HBM, then the static-base precomputation for a share of that length.
static int srs_part_create(const void *src, bool src_on_device, int src_device, size_t off, size_t n, Srs *out) {
    Srs s;
    s.off = off;
    s.n = n;
    Scope sc;
    TRY(sc.begin(g->stream));
    // H2B_TRACE=1: where the time of a registration goes (allocation calls vs kernels), one stderr line per phase
    static const bool trace = getenv("H2B_TRACE") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "h2b trace: srs n=%zu %s %.3f ms\n", n, what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
    };
    cudaError_t e = cudaMalloc(&s.d, n * sizeof(Affine));
    phase("cudaMalloc(points)");
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(H2B_ERR_OOM, "cudaMalloc(srs)", e);
    }
    const Affine *from = (const Affine *)src + off;
    if (!src_on_device) e = g->copier->h2d(s.d, from, n * sizeof(Affine), g->stream);
    else if (src_device == g->device) e = cudaMemcpyAsync(s.d, from, n * sizeof(Affine), cudaMemcpyDeviceToDevice, g->stream);
    else e = cudaMemcpyPeerAsync(s.d, g->device, from, src_device, n * sizeof(Affine), g->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
    phase("copy points");
    if (e != cudaSuccess) {
        cudaFree(s.d);
        return fail(H2B_ERR_CUDA, "cudaMemcpy(srs)", e);
    }
    // Static bases: precompute 2^(c*w) * P_i once so that every window of a commit feeds ONE bucket set
    // (fewer, wider windows; no Horner).  Skipped when the table would not fit comfortably.
    if (g->srs_precompute == 1 && g->srs_window == 0 && n <= g->comb_max_n) {
        // small SRS (the reference's circuits: k <= 14): bucket-free table of all window multiples
        const uint32_t c = g->comb_c, W = msm_windows_for(c), M = 1u << (c - 1);
        const size_t count = (size_t)W * M * n + 1, bytes = count * sizeof(Affine);
        const size_t scratch_bytes = (size_t)(M - 1) * n * W * sizeof(Fe);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        if (bytes + scratch_bytes <= free_b / 3 && count < (1u << 31)) {
            Fe *scratch = nullptr;
            e = cudaMalloc(&s.comb, bytes);
            if (e == cudaSuccess && (e = cudaMalloc(&scratch, scratch_bytes ? scratch_bytes : 16)) != cudaSuccess) {
                cudaFree(s.comb);
                s.comb = nullptr;
            }
            phase("cudaMalloc(multiples table + scratch)");
            if (e == cudaSuccess) {
                e = cudaMemsetAsync(s.comb + (count - 1), 0, sizeof(Affine), g->stream);  // the identity entry
                msm_precompute_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, g->stream>>>(s.d, (uint32_t)n, c, W, (size_t)M * n, s.comb);
                msm_comb_multiples_kernel<<<(uint32_t)((n * W + 127) / 128), 128, 0, g->stream>>>((uint32_t)n, c, W, s.comb, scratch);
                g->launches += 2;
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
                phase("multiples table kernels");
                cudaFree(scratch);
                phase("cudaFree(scratch)");
                if (e != cudaSuccess) {
                    cudaFree(s.comb);
                    cudaFree(s.d);
                    return fail(H2B_ERR_CUDA, "srs comb table", e);
                }
                s.comb_c = c;
                s.comb_w = W;
            } else {
                (void)cudaGetLastError();
                s.comb = nullptr;
            }
        }
    }
    if (!s.comb && g->srs_precompute && n >= 2) {
        uint32_t c = srs_window_for(n), W = msm_windows_for(c);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        // every t-th window power: the whole table (t = 1) while it fits a sixth of the free HBM (2^24 points: 14 GB), else
        // the thinnest stride up to 4 that does (2^26 points: 51.5 GB at t = 1, 27.5 GB at t = 2 -- two such arrays stay
        // resident).  Measured on B200 (commit, ms, t = 1 / 2 / 3): 2^20 2.91 / 3.34 / 3.81, 2^22 9.59 / 10.25 / 10.72,
        // 2^24 36.81 / 36.90 / 37.27 device-resident -- but 37.9-38.2 / 38.8 ms end to end from host memory: every copy
        // piece of the pipelined path sorts into t times the buckets (profiles/r02_e2e_chunk_probe.txt), so the thinner
        // table is for when HBM is short, not the default.
        uint32_t t = g->srs_table_stride ? g->srs_table_stride : 1;
        auto table_windows = [&](uint32_t tt) { return (W + tt - 1) / tt; };
        if (!g->srs_table_stride)
            while (t < 4 && (size_t)table_windows(t) * n * sizeof(Affine) > free_b / 6) t++;
        if (t > W) t = W;
        const uint32_t Wt = table_windows(t);
        const size_t bytes = (size_t)Wt * n * sizeof(Affine);
        if (bytes <= free_b / 3 && (size_t)Wt * n < (1u << 31)) {
            e = cudaMalloc(&s.table, bytes);
            phase("cudaMalloc(window table)");
            if (e == cudaSuccess) {
                s.tstride = t;
                msm_precompute_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, g->stream>>>(s.d, (uint32_t)n, c * t, Wt, n, s.table);
                g->launches++;
                e = cudaGetLastError();
                if (e == cudaSuccess) e = cudaStreamSynchronize(g->stream);
                phase("window table kernel");
                if (e != cudaSuccess) {
                    cudaFree(s.table);
                    cudaFree(s.d);
                    return fail(H2B_ERR_CUDA, "srs precompute", e);
                }
                s.c = c;
                s.windows = W;
            } else {
                (void)cudaGetLastError();
                s.table = nullptr;
            }
        }
    }
    *out = s;
    return H2B_OK;
}

// Registers a base array on every device of the library: replicated below g_shard_min_n points, sharded by
// contiguous point range (each device precomputes the table of its own share) from there up.
static int srs_register_locked(const void *bases, size_t n, uint64_t *handle, bool on_device = false) {
    TRY(ensure_ctx());
    if (!bases || !handle || n == 0) return fail(H2B_ERR_ARG, "srs_register: bad argument");
    CU(cudaSetDevice(g->device));
    if (on_device) CU(cudaDeviceSynchronize());  // the caller's producer of d_bases may run on any stream
    const size_t D = g_all.size();
    SrsSet set;
    set.n = n;
    set.replicated = D > 1 && n < g_shard_min_n;
    const size_t nparts = D;
    set.parts.resize(nparts);
    std::vector<int> devs;
    std::vector<size_t> lo(nparts), cnt(nparts);
    for (size_t i = 0; i < nparts; i++) {
        devs.push_back((int)i);
        if (D == 1 || set.replicated) {
            lo[i] = 0;
            cnt[i] = n;
        } else {  // equal ranges, boundaries on multiples of 256 points
            const size_t per = ((n + D - 1) / D + 255) & ~(size_t)255;
            lo[i] = std::min(n, i * per);
            cnt[i] = std::min(n, (i + 1) * per) - lo[i];
        }
    }
    while (!cnt.empty() && cnt.back() == 0) {  // fewer shares than devices (tiny tail)
        cnt.pop_back();
        lo.pop_back();
        devs.pop_back();
        set.parts.pop_back();
    }
    const int src_device = g->device;
    int rc = on_devices(devs, [&](size_t i) -> int {
        TRY(srs_part_create(bases, on_device, src_device, lo[i], cnt[i], &set.parts[i]));
        set.parts[i].dev = devs[i];
        return H2B_OK;
    });
    if (rc != H2B_OK) {
        std::string keep = g_err;
        for (Srs &pt : set.parts) {
            pt.dev = pt.dev < (int)D ? pt.dev : 0;
            srs_free(pt);
        }
        cudaSetDevice(g->device);
        g_err = keep;
        return rc;
    }
    *handle = g_next_handle++;
    g_srs[*handle] = set;
    return H2B_OK;
}
int h2b_srs_register(const uint64_t *bases, size_t n, uint64_t *handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    return srs_register_locked(bases, n, handle);
}
int h2b_dev_srs_register(const void *d_bases, size_t n, uint64_t *handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    return srs_register_locked(d_bases, n, handle, /*on_device=*/true);
}
// ParamsKZG::read, SerdeFormat::RawBytes (what ParamsKZG::write produces; the reference moves its SRS
// between setup and prover in this form, /root/reference/circuits/src/wasm.rs:52, :79, :126):
//   k: u32 LE | g[2^k] x 64 B | g_lagrange[2^k] x 64 B | g2 128 B | s_g2 128 B
// with every coordinate as its four Montgomery limbs -- the body of the file IS the base array, so both
// arrays go to HBM straight from the caller's buffer.
int h2b_params_read(const uint8_t *bytes, size_t len, uint32_t *k_out, uint64_t *g_handle, uint64_t *g_lagrange_handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!bytes || !g_handle || !g_lagrange_handle) return fail(H2B_ERR_ARG, "params_read: null pointer");
    if (len < 4) return fail(H2B_ERR_ARG, "params_read: truncated header");
    const uint32_t k = (uint32_t)bytes[0] | ((uint32_t)bytes[1] << 8) | ((uint32_t)bytes[2] << 16) | ((uint32_t)bytes[3] << 24);
    if (k > 28) return fail(H2B_ERR_ARG, "params_read: k > 28");
    const size_t n = (size_t)1 << k;
    if (len != 4 + 128 * n + 256) return fail(H2B_ERR_ARG, "params_read: length is not 4 + 128 * 2^k + 256");
    uint64_t hg = 0, hl = 0;
    TRY(srs_register_locked(bytes + 4, n, &hg));
    int rc = srs_register_locked(bytes + 4 + 64 * n, n, &hl);
    if (rc != H2B_OK) {
        std::string keep = g_err;
        auto it = g_srs.find(hg);
        if (it != g_srs.end()) {
            for (Srs &pt : it->second.parts) srs_free(pt);
            g_srs.erase(it);
        }
        cudaSetDevice(g->device);
        g_err = keep;
        return rc;
    }
    if (k_out) *k_out = k;
    *g_handle = hg;
    *g_lagrange_handle = hl;
    return H2B_OK;
}
// One registered array back to host memory, whatever its layout over the devices (unified addressing: a device-to-host copy
// needs no current-device switch).
static int srs_read_back(const SrsSet &set, uint8_t *out) {
    if (set.replicated || set.parts.size() == 1) {
        CU(cudaMemcpy(out, set.parts[0].d, set.n * sizeof(Affine), cudaMemcpyDeviceToHost));
        return H2B_OK;
    }
    for (const Srs &pt : set.parts)
        CU(cudaMemcpy(out + pt.off * sizeof(Affine), pt.d, pt.n * sizeof(Affine), cudaMemcpyDeviceToHost));
    return H2B_OK;
}
int h2b_params_write(uint32_t k, uint64_t g_handle, uint64_t g_lagrange_handle, const uint8_t *g2_and_s_g2, uint8_t *out, size_t len) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!g2_and_s_g2 || !out) return fail(H2B_ERR_ARG, "params_write: null pointer");
    if (k > 28) return fail(H2B_ERR_ARG, "params_write: k > 28");
    const size_t n = (size_t)1 << k;
    if (len != 4 + 128 * n + 256) return fail(H2B_ERR_ARG, "params_write: length is not 4 + 128 * 2^k + 256");
    auto ig = g_srs.find(g_handle), il = g_srs.find(g_lagrange_handle);
    if (ig == g_srs.end() || il == g_srs.end()) return fail(H2B_ERR_STATE, "params_write: unknown handle");
    if (ig->second.n != n || il->second.n != n) return fail(H2B_ERR_ARG, "params_write: a base array is not 2^k points long");
    CU(cudaStreamSynchronize(g->stream));
    for (int i = 0; i < 4; i++) out[i] = (uint8_t)(k >> (8 * i));
    TRY(srs_read_back(ig->second, out + 4));
    TRY(srs_read_back(il->second, out + 4 + 64 * n));
    memcpy(out + 4 + 128 * n, g2_and_s_g2, 256);
    return H2B_OK;
}
int h2b_srs_release(uint64_t handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    auto it = g_srs.find(handle);
    if (it == g_srs.end()) return fail(H2B_ERR_STATE, "srs_release: unknown handle");
    for (Srs &pt : it->second.parts) srs_free(pt);
    g_srs.erase(it);
    CU(cudaSetDevice(g->device));
    return H2B_OK;
}
int h2b_srs_device_ptr(uint64_t srs, void **d_bases, size_t *n) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    auto it = g_srs.find(srs);
    if (it == g_srs.end()) return fail(H2B_ERR_STATE, "srs_device_ptr: unknown handle");
    if (d_bases) *d_bases = it->second.parts[0].d;
    if (n) *n = it->second.parts[0].n;
    return H2B_OK;
}
int h2b_srs_info(uint64_t srs, size_t *n, uint32_t *window_bits, uint32_t *windows, size_t *table_bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    auto it = g_srs.find(srs);
    if (it == g_srs.end()) return fail(H2B_ERR_STATE, "srs_info: unknown handle");
    const Srs &sr = it->second.parts[0];
    if (n) *n = it->second.n;
    if (window_bits) *window_bits = sr.comb ? sr.comb_c : (sr.table ? sr.c : 0);
    if (windows) *windows = sr.comb ? sr.comb_w : (sr.table ? sr.windows : 0);
    if (table_bytes)
        *table_bytes = sr.comb ? ((size_t)sr.comb_w * (1u << (sr.comb_c - 1)) * sr.n + 1) * sizeof(Affine)
                               : (sr.table ? (size_t)((sr.windows + sr.tstride - 1) / sr.tstride) * sr.n * sizeof(Affine) : 0);
    return H2B_OK;
}
int h2b_srs_layout(uint64_t srs, uint32_t *parts, uint32_t *replicated, size_t *part_n /* parts entries, may be null */) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    auto it = g_srs.find(srs);
    if (it == g_srs.end()) return fail(H2B_ERR_STATE, "srs_layout: unknown handle");
    if (part_n)
        for (size_t i = 0; i < it->second.parts.size(); i++) part_n[i] = it->second.parts[i].n;
    if (parts) *parts = (uint32_t)it->second.parts.size();
    if (replicated) *replicated = it->second.replicated ? 1u : 0u;
    return H2B_OK;
}

int h2b_dev_g1_fold(const void *d_points, size_t count, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!d_out || (count && !d_points)) return fail(H2B_ERR_ARG, "g1_fold: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    g1_fold_kernel<<<1, 32, 0, s>>>((const Projective *)d_points, (uint32_t)count, (Projective *)d_out);
    LAUNCHED();
    return H2B_OK;
}
int h2b_dev_fixed_base_mul(const void *d_scalars, size_t n, const uint64_t base[8], void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!base || (n && (!d_scalars || !d_out))) return fail(H2B_ERR_ARG, "fixed_base_mul: null pointer");
    if (n > (1u << 30)) return fail(H2B_ERR_ARG, "fixed_base_mul: n > 2^30");
    if (n == 0) return H2B_OK;
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    Affine b;
    memcpy(&b, base, 64);
    g1_fixed_base_mul_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, s>>>((const Fe *)d_scalars, (uint32_t)n, b,
                                                                         (Affine *)d_out);
    LAUNCHED();
    return H2B_OK;
}
// ---- evaluate_h
static bool evalh_source_ok(uint64_t src, const h2b_eval_h *a) {
    const uint32_t kind = (uint32_t)(src & 0xff), x = (uint32_t)((src >> 8) & 0xfffffff), b = (uint32_t)(src >> 36);
    switch (kind) {
        case VS_CONSTANT: return x < a->num_constants;
        case VS_INTERMEDIATE: return x < a->num_intermediates;
        case VS_FIXED: return x < a->num_fixed && b < a->num_rotations;
        case VS_ADVICE: return x < a->num_advice && b < a->num_rotations;
        case VS_INSTANCE: return x < a->num_instance && b < a->num_rotations;
        case VS_CHALLENGE: return x < a->num_challenges;
        default: return kind <= VS_PREVIOUS;
    }
}
struct EvalhLookupPtrs {
    const void *product, *permuted_input, *permuted_table;
};
static int evalh_run(const h2b_domain *d, const h2b_eval_h *a, const EvalhLookupPtrs *lookup, void *d_values, void *stream);
int h2b_dev_evaluate_h(const h2b_domain *d, const h2b_eval_h *a, void *d_values, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    return evalh_run(d, a, nullptr, d_values, stream);
}
int h2b_dev_evaluate_h_lookup(const h2b_domain *d, const h2b_eval_h *a, const void *d_product, const void *d_permuted_input,
                              const void *d_permuted_table, void *d_values, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!d_product || !d_permuted_input || !d_permuted_table) return fail(H2B_ERR_ARG, "evaluate_h_lookup: null pointer");
    if (a && (!a->l0 || !a->l_last || !a->l_active_row)) return fail(H2B_ERR_ARG, "evaluate_h_lookup: null l0 / l_last / l_active_row");
    const EvalhLookupPtrs lp{d_product, d_permuted_input, d_permuted_table};
    return evalh_run(d, a, &lp, d_values, stream);
}
static int evalh_run(const h2b_domain *d, const h2b_eval_h *a, const EvalhLookupPtrs *lookup, void *d_values, void *stream) {
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!a || !d_values) return fail(H2B_ERR_ARG, "evaluate_h: null pointer");
    if (a->num_rotations > kMaxRotations) return fail(H2B_ERR_ARG, "evaluate_h: more than 32 distinct rotations");
    if ((a->num_fixed && !a->fixed) || (a->num_advice && !a->advice) || (a->num_instance && !a->instance) ||
        (a->num_challenges && !a->challenges) || (a->num_constants && !a->constants) || (a->num_rotations && !a->rotations) ||
        (a->num_calcs && !a->calcs))
        return fail(H2B_ERR_ARG, "evaluate_h: null array");
    // validate the serialised graph before it is trusted with device pointers
    {
        size_t w = 0;
        for (uint32_t c = 0; c < a->num_calcs; c++) {
            if (w >= a->calc_words) return fail(H2B_ERR_ARG, "evaluate_h: truncated calculation list");
            const uint64_t hdr = a->calcs[w++];
            const uint32_t op = (uint32_t)(hdr & 0xff), target = (uint32_t)((hdr >> 8) & 0xffffffffu), nparts = (uint32_t)(hdr >> 40);
            if (op > CALC_STORE || target >= a->num_intermediates) return fail(H2B_ERR_ARG, "evaluate_h: bad calculation");
            const size_t nsrc = op == CALC_HORNER ? 2 + (size_t)nparts : (op <= CALC_MUL ? 2 : 1);
            if (w + nsrc > a->calc_words) return fail(H2B_ERR_ARG, "evaluate_h: truncated calculation list");
            for (size_t k = 0; k < nsrc; k++)
                if (!evalh_source_ok(a->calcs[w + k], a)) return fail(H2B_ERR_ARG, "evaluate_h: value source out of range");
            w += nsrc;
        }
        if (w != a->calc_words) return fail(H2B_ERR_ARG, "evaluate_h: calc_words does not match the calculation list");
    }
    const uint32_t P = lookup ? 0 : a->num_perm_columns;
    uint32_t sets = 0;
    if (P) {
        if (a->chunk_len == 0) return fail(H2B_ERR_ARG, "evaluate_h: chunk_len = 0");
        sets = (P + a->chunk_len - 1) / a->chunk_len;
        if (!a->perm_kind || !a->perm_index || !a->sigma_cosets || !a->z_cosets || !a->l0 || !a->l_last || !a->l_active_row)
            return fail(H2B_ERR_ARG, "evaluate_h: null permutation data");
        for (uint32_t c = 0; c < P; c++) {
            const uint32_t lim = a->perm_kind[c] == 0 ? a->num_advice : (a->perm_kind[c] == 1 ? a->num_fixed : a->num_instance);
            if (a->perm_kind[c] > 2 || a->perm_index[c] >= lim) return fail(H2B_ERR_ARG, "evaluate_h: permutation column out of range");
        }
    }
    // Intermediates are renumbered by liveness: a slot is free again once its value has been read for the last
    // time, so the working set of a row is the graph's width, not its length (upstream's GraphEvaluator numbers every
    // calculation; a few hundred of them at extended_k = 22 would otherwise need tens of GiB of scratch).
    std::vector<uint64_t> calcs(a->calcs, a->calcs + a->calc_words);
    uint32_t slots = 0;
    {
        const uint32_t NI = a->num_intermediates, NC = a->num_calcs;
        std::vector<size_t> at(NC);                 // word offset of every calculation
        std::vector<uint32_t> last_use(NI, 0);      // last calculation that reads intermediate t (as numbered by the caller)
        size_t w = 0;
        auto nsrc_of = [&](uint64_t hdr) -> size_t {
            const uint32_t op = (uint32_t)(hdr & 0xff), nparts = (uint32_t)(hdr >> 40);
            return op == CALC_HORNER ? 2 + (size_t)nparts : (op <= CALC_MUL ? 2 : 1);
        };
        for (uint32_t c = 0; c < NC; c++) {
            at[c] = w;
            const size_t ns = nsrc_of(calcs[w]);
            for (size_t k = 1; k <= ns; k++)
                if ((calcs[w + k] & 0xff) == VS_INTERMEDIATE) last_use[(uint32_t)((calcs[w + k] >> 8) & 0xfffffff)] = c;
            w += 1 + ns;
        }
        std::vector<int64_t> slot_of(NI, -1);
        std::vector<uint32_t> free_slots;
        for (uint32_t c = 0; c < NC; c++) {
            const size_t o = at[c], ns = nsrc_of(calcs[o]);
            std::vector<uint32_t> dying;
            for (size_t k = 1; k <= ns; k++) {
                uint64_t &src = calcs[o + k];
                if ((src & 0xff) != VS_INTERMEDIATE) continue;
                const uint32_t t = (uint32_t)((src >> 8) & 0xfffffff);
                if (slot_of[t] < 0) return fail(H2B_ERR_ARG, "evaluate_h: intermediate read before it is written");
                src = (src & ~((uint64_t)0xfffffff << 8)) | ((uint64_t)slot_of[t] << 8);
                if (last_use[t] == c) dying.push_back(t);
            }
            for (uint32_t t : dying)
                if (slot_of[t] >= 0) {
                    free_slots.push_back((uint32_t)slot_of[t]);
                    slot_of[t] = -1;
                }
            const uint32_t target = (uint32_t)((calcs[o] >> 8) & 0xffffffffu);
            if (slot_of[target] >= 0) free_slots.push_back((uint32_t)slot_of[target]);  // rewritten: the old value is dead
            uint32_t sl;
            if (!free_slots.empty()) {
                sl = free_slots.back();
                free_slots.pop_back();
            } else {
                sl = slots++;
            }
            slot_of[target] = sl;
            calcs[o] = (calcs[o] & ~((uint64_t)0xffffffffu << 8)) | ((uint64_t)sl << 8);
            if (last_use[target] <= c) {  // never read afterwards (the row's value is taken from the last calculation)
                free_slots.push_back(sl);
                slot_of[target] = -1;
            }
        }
    }
    const bool local = slots <= kLocalSlots;
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    const uint32_t size = 1u << d->extended_k;
    // one staging blob: pointer tables, scalars, graph
    std::vector<uint64_t> blob;
    auto put = [&](const void *src, size_t bytes) {
        if (blob.size() & 1) blob.push_back(0);  // field elements are read with 128-bit loads
        const size_t off = blob.size();
        blob.resize(off + (bytes + 7) / 8);
        if (bytes) memcpy(&blob[off], src, bytes);
        return off;
    };
    const size_t o_fixed = put(a->fixed, a->num_fixed * sizeof(void *)), o_adv = put(a->advice, a->num_advice * sizeof(void *));
    const size_t o_inst = put(a->instance, a->num_instance * sizeof(void *));
    const size_t o_chal = put(a->challenges, (size_t)a->num_challenges * 32), o_const = put(a->constants, (size_t)a->num_constants * 32);
    const size_t o_rot = put(a->rotations, a->num_rotations * sizeof(int32_t)), o_calc = put(calcs.data(), calcs.size() * 8);
    std::vector<const void *> cols(P);
    for (uint32_t c = 0; c < P; c++)
        cols[c] = a->perm_kind[c] == 0 ? a->advice[a->perm_index[c]] : (a->perm_kind[c] == 1 ? a->fixed[a->perm_index[c]] : a->instance[a->perm_index[c]]);
    const size_t o_cols = put(cols.data(), P * sizeof(void *)), o_sig = put(a->sigma_cosets, P * sizeof(void *));
    const size_t o_z = put(a->z_cosets, sets * sizeof(void *));
    uint64_t *dblob;
    TRY(get_buf(BUF_EVALH, blob.size() * 8 + 8, (void **)&dblob));
    // (a copy from pageable memory has left the source buffer when cudaMemcpyAsync returns: `blob` may go out of scope)
    if (!blob.empty()) CU(cudaMemcpyAsync(dblob, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice, s));
    Fe *scratch = nullptr;
    if (!local) TRY(get_buf(BUF_EVALH_SCRATCH, (size_t)slots * size * sizeof(Fe), (void **)&scratch));
    EvalGates eg;
    eg.fixed = (const Fe *const *)(dblob + o_fixed);
    eg.advice = (const Fe *const *)(dblob + o_adv);
    eg.instance = (const Fe *const *)(dblob + o_inst);
    eg.challenges = (const Fe *)(dblob + o_chal);
    eg.constants = (const Fe *)(dblob + o_const);
    eg.rotations = (const int32_t *)(dblob + o_rot);
    eg.calcs = dblob + o_calc;
    eg.scratch = scratch;
    eg.num_rotations = a->num_rotations;
    eg.num_calcs = a->num_calcs;
    eg.size = size;
    eg.rot_scale = 1u << (d->extended_k - d->k);
    memcpy(&eg.beta, a->beta, 32);
    memcpy(&eg.gamma, a->gamma, 32);
    memcpy(&eg.theta, a->theta, 32);
    memcpy(&eg.y, a->y, 32);
    const uint32_t blocks = (size + 127) / 128;
    if (lookup) {
        EvalLookup el;
        el.product = (const Fe *)lookup->product;
        el.permuted_input = (const Fe *)lookup->permuted_input;
        el.permuted_table = (const Fe *)lookup->permuted_table;
        el.l0 = (const Fe *)a->l0;
        el.l_last = (const Fe *)a->l_last;
        el.l_active = (const Fe *)a->l_active_row;
        if (local) evalh_lookup_kernel<true><<<blocks, 128, 0, s>>>(eg, el, (Fe *)d_values);
        else evalh_lookup_kernel<false><<<blocks, 128, 0, s>>>(eg, el, (Fe *)d_values);
        LAUNCHED();
        return H2B_OK;
    }
    EvalPerm ep;
    memset(&ep, 0, sizeof ep);
    if (P) {
        static const uint32_t kDeltaMont[8] = {0xefd78855u, 0x9a0c322bu, 0x249b563cu, 0x46e82d14u,
                                               0xe0b0b7a7u, 0x5983a663u, 0xaaa111adu, 0x22ab452bu};  // Fr::DELTA = 7^(2^28)
        TwEntry *tw;  // extended_omega^i: the table the extended-domain transforms use anyway
        TRY(get_twiddles(d->extended_omega, d->extended_k, s, &tw));
        ep.columns = (const Fe *const *)(dblob + o_cols);
        ep.sigma = (const Fe *const *)(dblob + o_sig);
        ep.z = (const Fe *const *)(dblob + o_z);
        ep.l0 = (const Fe *)a->l0;
        ep.l_last = (const Fe *)a->l_last;
        ep.l_active = (const Fe *)a->l_active_row;
        ep.ext_pows = tw->W;
        ep.num_columns = P;
        ep.chunk_len = a->chunk_len;
        ep.num_sets = sets;
        ep.size = size;
        ep.rot_scale = eg.rot_scale;
        ep.last_rotation = a->last_rotation;
        ep.beta = eg.beta;
        ep.gamma = eg.gamma;
        ep.y = eg.y;
        memcpy(&ep.delta, kDeltaMont, 32);
        memcpy(&ep.zeta, d->g_coset, 32);
    }
    if (local) evalh_fused_kernel<true><<<blocks, 128, 0, s>>>(eg, ep, (Fe *)d_values, a->flags & H2B_EVALH_ACCUMULATE ? 1u : 0u);
    else evalh_fused_kernel<false><<<blocks, 128, 0, s>>>(eg, ep, (Fe *)d_values, a->flags & H2B_EVALH_ACCUMULATE ? 1u : 0u);
    LAUNCHED();
    return H2B_OK;
}

int h2b_g_to_lagrange(const uint64_t *g_bases, uint32_t k, uint64_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!g_bases || !out) return fail(H2B_ERR_ARG, "g_to_lagrange: null pointer");
    if (k > 26) return fail(H2B_ERR_ARG, "g_to_lagrange: k > 26");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = g->stream;
    Scope sc;
    TRY(sc.begin(s));
    const size_t n = (size_t)1 << k;
    h2b_domain d;
    TRY(domain_build(3, k, &d));  // omega_inv and 1/2^k do not depend on j
    void *din;
    XYZZ *work;
    Affine *dout;
    TRY(stage_in(BUF_BASES, g_bases, n * sizeof(Affine), &din));
    TRY(get_buf(BUF_ECNTT, n * sizeof(XYZZ), (void **)&work));
    TRY(get_buf(BUF_TEST_O, n * sizeof(Affine), (void **)&dout));
    const Fe *W = nullptr;
    if (k > 0) {
        TwEntry *tw;
        TRY(get_twiddles(d.omega_inv, k, s, &tw));
        W = tw->W;
    }
    const uint32_t blocks = (uint32_t)((n + 127) / 128);
    ec_ntt_load_kernel<<<blocks, 128, 0, s>>>((const Affine *)din, k, work);
    LAUNCHED();
    for (uint32_t st = 0; st < k; st++) {
        ec_ntt_stage_kernel<<<(uint32_t)((n / 2 + 127) / 128), 128, 0, s>>>(work, k, st, W);
        LAUNCHED();
    }
    Fe scale;
    memcpy(&scale, d.ifft_divisor, 32);
    ec_ntt_finish_kernel<<<blocks, 128, 0, s>>>(work, (uint32_t)n, scale, dout);
    LAUNCHED();
    TRY(copy_out(out, dout, n * sizeof(Affine), s));
    return H2B_OK;
}
int h2b_g1_to_bytes(const uint64_t *points, size_t m, uint8_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (m == 0) return H2B_OK;
    if (!points || !out) return fail(H2B_ERR_ARG, "g1_to_bytes: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *dp = nullptr, *dout;
    TRY(stage_in(BUF_TEST_A, points, m * 96, &dp));
    TRY(get_buf(BUF_TEST_O, m * 32, &dout));
    g1_to_bytes_kernel<<<(uint32_t)((m + 63) / 64), 64, 0, g->stream>>>((const Projective *)dp, (uint32_t)m, (uint32_t *)dout);
    LAUNCHED();
    TRY(copy_out(out, dout, m * 32, g->stream));
    return H2B_OK;
}
int h2b_g1_fold(const uint64_t *points, size_t count, uint64_t out[12]) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!out || (count && !points)) return fail(H2B_ERR_ARG, "g1_fold: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *dp = nullptr, *dout;
    TRY(stage_in(BUF_TEST_A, points, count * 96, &dp));
    TRY(get_buf(BUF_OUT, 96, &dout));
    g1_fold_kernel<<<1, 32, 0, g->stream>>>((const Projective *)dp, (uint32_t)count, (Projective *)dout);
    LAUNCHED();
    TRY(copy_out(out, dout, 96, g->stream));
    return H2B_OK;
}

// ---- NTT
int h2b_dev_best_fft(void *d_a, const uint64_t omega[4], uint32_t log_n, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!d_a || !omega) return fail(H2B_ERR_ARG, "best_fft: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return ntt_run((const Fe *)d_a, (Fe *)d_a, log_n, omega, io_plain(log_n), s);
}
int h2b_best_fft(uint64_t *a, const uint64_t omega[4], uint32_t log_n) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!a || !omega) return fail(H2B_ERR_ARG, "best_fft: null pointer");
    if (log_n > 28) return fail(H2B_ERR_ARG, "best_fft: log_n > 28");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    size_t bytes = ((size_t)1 << log_n) * 32;
    void *da;
    TRY(stage_in(BUF_NTT_A, a, bytes, &da, /*round_trip=*/true));
    TRY(ntt_run((const Fe *)da, (Fe *)da, log_n, omega, io_plain(log_n), g->stream));
    TRY(copy_out(a, da, bytes, g->stream));
    return H2B_OK;
}

int h2b_domain_new(uint32_t j, uint32_t k, h2b_domain *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!out) return fail(H2B_ERR_ARG, "domain_new: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    return domain_build(j, k, out);
}

int h2b_dev_lagrange_to_coeff(const h2b_domain *d, void *d_a, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!d_a) return fail(H2B_ERR_ARG, "lagrange_to_coeff: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return dev_lagrange_to_coeff(d, (Fe *)d_a, s);
}
int h2b_lagrange_to_coeff(const h2b_domain *d, uint64_t *a) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!a) return fail(H2B_ERR_ARG, "lagrange_to_coeff: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    size_t bytes = ((size_t)1 << d->k) * 32;
    void *da;
    TRY(stage_in(BUF_NTT_A, a, bytes, &da, /*round_trip=*/true));
    TRY(dev_lagrange_to_coeff(d, (Fe *)da, g->stream));
    TRY(copy_out(a, da, bytes, g->stream));
    return H2B_OK;
}
int h2b_dev_coeff_to_extended(const h2b_domain *d, const void *d_in, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!d_in || !d_out) return fail(H2B_ERR_ARG, "coeff_to_extended: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return dev_coeff_to_extended(d, (const Fe *)d_in, (Fe *)d_out, s);
}
int h2b_coeff_to_extended(const h2b_domain *d, const uint64_t *in, uint64_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!in || !out) return fail(H2B_ERR_ARG, "coeff_to_extended: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    size_t in_bytes = ((size_t)1 << d->k) * 32, out_bytes = ((size_t)1 << d->extended_k) * 32;
    void *din, *dout;
    TRY(stage_in(BUF_NTT_IN, in, in_bytes, &din, /*round_trip=*/true));
    TRY(get_buf(BUF_NTT_A, out_bytes, &dout));
    TRY(dev_coeff_to_extended(d, (const Fe *)din, (Fe *)dout, g->stream));
    TRY(copy_out(out, dout, out_bytes, g->stream));
    return H2B_OK;
}
// The columns [0, m) on the calling thread's context.
static int lagrange_to_coeff_many_local(const h2b_domain *d, uint64_t *const *cols, size_t m) {
    cudaStream_t s = g->stream;
    Scope sc;
    TRY(sc.begin(s));
    const size_t n = (size_t)1 << d->k, bytes = n * 32;
    size_t step = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(m, 65535), ((size_t)1 << 30) / bytes));
    for (size_t q0 = 0; q0 < m; q0 += step) {
        const size_t cnt = std::min(step, m - q0);
        Fe *da;
        TRY(get_buf(BUF_NTT_A, cnt * bytes, (void **)&da));
        std::vector<HostCopier::Seg> segs;
        for (size_t q = 0; q < cnt; q++) segs.push_back({da + q * n, cols[q0 + q], bytes});
        cudaError_t ce = g->copier->h2d(segs, s);
        if (ce != cudaSuccess) return fail(H2B_ERR_CUDA, "host-to-device copy", ce);
        TRY(dev_lagrange_to_coeff(d, da, s, (uint32_t)cnt));
        ce = g->copier->d2h(segs, s);
        if (ce != cudaSuccess) return fail(H2B_ERR_CUDA, "device-to-host copy", ce);
    }
    return H2B_OK;
}
static int coeff_to_extended_many_local(const h2b_domain *d, const uint64_t *const *in, uint64_t *const *out, size_t m) {
    cudaStream_t s = g->stream;
    Scope sc;
    TRY(sc.begin(s));
    const size_t n = (size_t)1 << d->k, en = (size_t)1 << d->extended_k;
    size_t step = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(m, 65535), ((size_t)1 << 30) / (en * 32)));
    for (size_t q0 = 0; q0 < m; q0 += step) {
        const size_t cnt = std::min(step, m - q0);
        Fe *din, *dout;
        TRY(get_buf(BUF_NTT_IN, cnt * n * 32, (void **)&din));
        TRY(get_buf(BUF_NTT_A, cnt * en * 32, (void **)&dout));
        std::vector<HostCopier::Seg> sin, sout;
        for (size_t q = 0; q < cnt; q++) {
            sin.push_back({din + q * n, const_cast<uint64_t *>(in[q0 + q]), n * 32});
            sout.push_back({dout + q * en, out[q0 + q], en * 32});
        }
        cudaError_t ce = g->copier->h2d(sin, s);
        if (ce != cudaSuccess) return fail(H2B_ERR_CUDA, "host-to-device copy", ce);
        TRY(dev_coeff_to_extended(d, din, dout, s, (uint32_t)cnt));
        ce = g->copier->d2h(sout, s);
        if (ce != cudaSuccess) return fail(H2B_ERR_CUDA, "device-to-host copy", ce);
    }
    return H2B_OK;
}
// Whole columns are independent transforms: with several devices they are dealt in contiguous blocks, one block per
// device (a single NTT stays on one GPU), each over its own PCIe link.  fn(first column, count) runs per device.
static int deal_columns(size_t m, const std::function<int(size_t, size_t)> &fn) {
    const size_t use = std::min(g_all.size(), m);
    if (use <= 1) return fn(0, m);
    std::vector<int> devs;
    std::vector<size_t> q0(use), cnt(use);
    for (size_t i = 0, at = 0; i < use; i++) {
        devs.push_back((int)i);
        cnt[i] = m / use + (i < m % use ? 1 : 0);
        q0[i] = at;
        at += cnt[i];
    }
    return on_devices(devs, [&](size_t i) -> int { return fn(q0[i], cnt[i]); });
}
// Batched column transforms (create_proof runs lagrange_to_coeff over every advice / instance / product
// column and coeff_to_extended over every column inside evaluate_h, one call each): the
// columns are copied in one after the other, transformed by ONE launch per pass (gridDim.y = columns),
// and copied back.  Sub-batches keep the device staging below 1 GiB.
int h2b_lagrange_to_coeff_many(const h2b_domain *d, uint64_t *const *cols, size_t m) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (m == 0) return H2B_OK;
    if (!cols) return fail(H2B_ERR_ARG, "lagrange_to_coeff_many: null pointer");
    for (size_t q = 0; q < m; q++)
        if (!cols[q]) return fail(H2B_ERR_ARG, "lagrange_to_coeff_many: null column");
    CU(cudaSetDevice(g->device));
    return deal_columns(m, [&](size_t q0, size_t cnt) { return lagrange_to_coeff_many_local(d, cols + q0, cnt); });
}
int h2b_coeff_to_extended_many(const h2b_domain *d, const uint64_t *const *in, uint64_t *const *out, size_t m) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (m == 0) return H2B_OK;
    if (!in || !out) return fail(H2B_ERR_ARG, "coeff_to_extended_many: null pointer");
    for (size_t q = 0; q < m; q++)
        if (!in[q] || !out[q]) return fail(H2B_ERR_ARG, "coeff_to_extended_many: null column");
    CU(cudaSetDevice(g->device));
    return deal_columns(m, [&](size_t q0, size_t cnt) { return coeff_to_extended_many_local(d, in + q0, out + q0, cnt); });
}
int h2b_dev_lagrange_to_coeff_many(const h2b_domain *d, void *d_a, size_t m, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (m == 0) return H2B_OK;
    if (!d_a) return fail(H2B_ERR_ARG, "lagrange_to_coeff_many: null pointer");
    if (m > 65535) return fail(H2B_ERR_ARG, "lagrange_to_coeff_many: more than 65535 columns");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return dev_lagrange_to_coeff(d, (Fe *)d_a, s, (uint32_t)m);
}
int h2b_dev_coeff_to_extended_many(const h2b_domain *d, const void *d_in, void *d_out, size_t m, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (m == 0) return H2B_OK;
    if (!d_in || !d_out) return fail(H2B_ERR_ARG, "coeff_to_extended_many: null pointer");
    if (m > 65535) return fail(H2B_ERR_ARG, "coeff_to_extended_many: more than 65535 columns");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return dev_coeff_to_extended(d, (const Fe *)d_in, (Fe *)d_out, s, (uint32_t)m);
}
int h2b_dev_extended_to_coeff(const h2b_domain *d, const void *d_in, void *d_out, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!d_in || !d_out) return fail(H2B_ERR_ARG, "extended_to_coeff: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    return dev_extended_to_coeff(d, (const Fe *)d_in, (Fe *)d_out, s);
}
int h2b_extended_to_coeff(const h2b_domain *d, const uint64_t *in, uint64_t *out) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!in || !out) return fail(H2B_ERR_ARG, "extended_to_coeff: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    size_t in_bytes = ((size_t)1 << d->extended_k) * 32;
    size_t out_bytes = ((size_t)(d->j - 1) << d->k) * 32;
    void *din, *dout;
    TRY(stage_in(BUF_NTT_IN, in, in_bytes, &din, /*round_trip=*/true));
    TRY(get_buf(BUF_NTT_OUT, out_bytes, &dout));
    TRY(dev_extended_to_coeff(d, (const Fe *)din, (Fe *)dout, g->stream));
    TRY(copy_out(out, dout, out_bytes, g->stream));
    return H2B_OK;
}
int h2b_divide_by_vanishing_poly(const h2b_domain *d, uint64_t *a) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!a) return fail(H2B_ERR_ARG, "divide_by_vanishing_poly: null pointer");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    size_t n = (size_t)1 << d->extended_k;
    void *da, *dt;
    TRY(stage_in(BUF_NTT_A, a, n * 32, &da, /*round_trip=*/true));
    TRY(stage_in(BUF_MISC, d->t_evaluations, (size_t)d->n_t * 32, &dt));
    fr_scale_cyclic_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, g->stream>>>((Fe *)da, (uint32_t)n,
                                                                               (const Fe *)dt, d->n_t);
    LAUNCHED();
    TRY(copy_out(a, da, n * 32, g->stream));
    return H2B_OK;
}

int h2b_dev_divide_by_vanishing_poly(const h2b_domain *d, void *d_a, void *stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    TRY(check_domain(d));
    if (!d_a) return fail(H2B_ERR_ARG, "divide_by_vanishing_poly: null pointer");
    CU(cudaSetDevice(g->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : g->stream;
    Scope sc;
    TRY(sc.begin(s));
    const size_t n = (size_t)1 << d->extended_k;
    void *dt;
    TRY(get_buf(BUF_MISC, (size_t)d->n_t * 32, &dt));
    CU(cudaMemcpyAsync(dt, d->t_evaluations, (size_t)d->n_t * 32, cudaMemcpyHostToDevice, s));
    fr_scale_cyclic_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>((Fe *)d_a, (uint32_t)n, (const Fe *)dt, d->n_t);
    LAUNCHED();
    return H2B_OK;
}

// ---- test hooks
int h2b_test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!a || !out || n == 0) return fail(H2B_ERR_ARG, "test_field_op: bad argument");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *da, *db = nullptr, *dout;
    TRY(stage_in(BUF_TEST_A, a, n * 32, &da));
    if (b) TRY(stage_in(BUF_TEST_B, b, n * 32, &db));
    TRY(get_buf(BUF_TEST_O, n * 32, &dout));
    uint32_t blocks = (uint32_t)((n + 127) / 128);
    if (field == 0)
        test_field_kernel<Fr><<<blocks, 128, 0, g->stream>>>(op, (Fe *)da, (Fe *)db, (Fe *)dout, (uint32_t)n);
    else
        test_field_kernel<Fq><<<blocks, 128, 0, g->stream>>>(op, (Fe *)da, (Fe *)db, (Fe *)dout, (uint32_t)n);
    LAUNCHED();
    TRY(copy_out(out, dout, n * 32, g->stream));
    return H2B_OK;
}
int h2b_test_g1_add_affine(const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (!a || !b || !out || n == 0) return fail(H2B_ERR_ARG, "test_g1_add_affine: bad argument");
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *da, *db, *dout;
    TRY(stage_in(BUF_TEST_A, a, n * 64, &da));
    TRY(stage_in(BUF_TEST_B, b, n * 64, &db));
    TRY(get_buf(BUF_TEST_O, n * 96, &dout));
    test_g1_add_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, g->stream>>>((Affine *)da, (Affine *)db,
                                                                           (Projective *)dout, (uint32_t)n);
    LAUNCHED();
    TRY(copy_out(out, dout, n * 96, g->stream));
    return H2B_OK;
}

int h2b_test_set_max_entries(uint32_t log2_entries) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    if (log2_entries != 0 && (log2_entries < 10 || log2_entries > 31)) return fail(H2B_ERR_ARG, "max entries must be 0 or 2^10..2^31");
    for (Ctx *x : g_all) x->max_entries = 1ull << (log2_entries ? log2_entries : 31);
    return H2B_OK;
}

int h2b_imad_peak(double *imad_gops, double *imad_wide_gops, double *sm_mhz) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(ensure_ctx());
    CU(cudaSetDevice(g->device));
    Scope sc;
    TRY(sc.begin(g->stream));
    void *sink;
    TRY(get_buf(BUF_MISC, 64, &sink));
    const uint32_t iters = 2000, threads = 512;
    const uint32_t blocks = (uint32_t)g->sm_count * 4;
    const double ops = (double)blocks * threads * iters * 16.0 * 8.0;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    float best[2] = {1e30f, 1e30f};
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, g->stream));
        imad_bench_kernel<<<blocks, threads, 0, g->stream>>>((uint32_t *)sink, iters, 12345u + rep);
        LAUNCHED();
        CU(cudaEventRecord(e1, g->stream));
        CU(cudaEventSynchronize(e1));
        float ms;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best[0]) best[0] = ms;
        CU(cudaEventRecord(e0, g->stream));
        imad_wide_bench_kernel<<<blocks, threads, 0, g->stream>>>((uint64_t *)sink, iters, 12345u + rep);
        LAUNCHED();
        CU(cudaEventRecord(e1, g->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best[1]) best[1] = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (imad_gops) *imad_gops = ops / (best[0] * 1e-3) / 1e9;
    if (imad_wide_gops) *imad_wide_gops = ops / (best[1] * 1e-3) / 1e9;
    if (sm_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g->device);
        *sm_mhz = khz / 1000.0;
    }
    return H2B_OK;
}

}  // extern "C"
