// hostcopy.h -- pageable host memory <-> HBM at PCIe speed.
//
// The reference's polynomials live in ordinary Rust Vecs (pageable memory).  cudaMemcpy from/to such memory
// is staged by the driver through one thread and reaches ~12-15 GB/s on this host, a quarter of what the
// link does from pinned memory, and at k <= 20 those copies cost more than the transforms they feed.
// Here the staging is done by a few worker threads into a pinned ring, chunk by chunk, with the DMA of a
// chunk overlapping the memcpy of the next ones.  Pinned caller buffers are detected and copied directly.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace h2b {

class HostCopier {
  public:
    static constexpr size_t kChunk = (size_t)512 << 10;  // bytes per staged chunk
    static constexpr size_t kMinSlots = 8, kSlots = 128;  // pinned ring: 4 MiB at first, grown on demand up to 64 MiB
    struct Seg {
        void *dev;
        void *host;
        size_t bytes;
    };
    // Below these totals the driver's own staging is used.  Measured on B200 / PCIe Gen5 (scripts/hostpath_probe.py):
    //  * device-to-host: the ring pays from 512 KiB when the call's input went through it too;
    //  * host-to-device alone (a commit: scalars in, 96 bytes out): the ring LOSES below 8 MiB (8 commits of 2 MiB each
    //    5.1 -> 7.1 ms, commit_many of 8 x 512 KiB 1.09 -> 1.6 ms), so only round trips (the NTT entry points, `round_trip`)
    //    stage their input from 512 KiB: best_fft on a pageable buffer at 2^14 / 2^16 / 2^17 0.169 / 0.412 / 0.710 ->
    //    0.104 / 0.258 / 0.518 ms.
    size_t kMinStagedH2D = (size_t)8 << 20, kMinStagedD2H = (size_t)512 << 10, kMinStagedRoundTrip = (size_t)512 << 10;
    // Below this many staged bytes the calling thread fills / drains the ring alone: waking the worker threads costs more
    // than their help is worth when they have been asleep (a prover doing host work between calls), and a process that
    // never moves this much never creates them (each costs ~10 ms to start: 58 / 111 ms on the first staged call of a
    // Poseidon proof at k = 12 / 14 with six of them, scripts/proof_hostpath_ab.py).
    size_t kMinWorkers = (size_t)8 << 20;

    explicit HostCopier(int threads) : nthreads_(threads) {
        if (const char *e = getenv("H2B_MIN_STAGED_H2D_KB")) kMinStagedH2D = (size_t)atol(e) << 10;
        if (const char *e = getenv("H2B_MIN_STAGED_D2H_KB")) kMinStagedD2H = (size_t)atol(e) << 10;
        if (const char *e = getenv("H2B_MIN_STAGED_RT_KB")) kMinStagedRoundTrip = (size_t)atol(e) << 10;
        if (const char *e = getenv("H2B_COPY_WORKERS_MIN_KB")) kMinWorkers = (size_t)atol(e) << 10;
    }
    ~HostCopier() { shutdown(); }
    int threads() const { return nthreads_; }
    // true when copies from/to `host` go through the pinned ring (pageable memory, staging enabled)
    bool stages(const void *host) const { return staged(host, 0); }

    // dev <- host for every segment.  On return every DMA has been enqueued and `s` waits for them.
    cudaError_t h2d(const std::vector<Seg> &segs, cudaStream_t s, bool round_trip = false) {
        std::vector<Piece> pieces;
        cudaError_t e = plan(segs, cudaMemcpyHostToDevice, s, pieces, round_trip ? kMinStagedRoundTrip : kMinStagedH2D);
        if (e != cudaSuccess || pieces.empty()) return e;
        if ((e = ensure(pieces.size())) != cudaSuccess) return e;
        const bool helpers = pieces.size() * kChunk >= kMinWorkers;
        for (size_t base = 0; base < pieces.size(); base += slots_) {
            const size_t cnt = std::min(slots_, pieces.size() - base);
            if (base && (e = cudaStreamSynchronize(copy_)) != cudaSuccess) return e;  // ring reused: drain its DMAs
            std::atomic<int> err{0};
            run(cnt, helpers, [&](size_t i) {
                const Piece &p = pieces[base + i];
                std::memcpy(ring_ + i * kChunk, p.host, p.len);
                if (cudaMemcpyAsync(p.dev, ring_ + i * kChunk, p.len, cudaMemcpyHostToDevice, copy_) != cudaSuccess) err = 1;
            });
            if (err) return cudaErrorUnknown;
        }
        if ((e = cudaEventRecord(done_, copy_)) != cudaSuccess) return e;
        busy_ = true;
        return cudaStreamWaitEvent(s, done_, 0);
    }
    cudaError_t h2d(void *dev, const void *host, size_t bytes, cudaStream_t s, bool round_trip = false) {
        return h2d(std::vector<Seg>{{dev, const_cast<void *>(host), bytes}}, s, round_trip);
    }

    // host <- dev for every segment, ordered after the work already enqueued on `s`.  Blocks until the host
    // buffers are complete (and `s` is idle).
    cudaError_t d2h(const std::vector<Seg> &segs, cudaStream_t s) {
        std::vector<Piece> pieces;
        cudaError_t e = plan(segs, cudaMemcpyDeviceToHost, s, pieces, kMinStagedD2H);
        if (e != cudaSuccess) return e;
        if (pieces.empty()) return cudaStreamSynchronize(s);
        if ((e = ensure(pieces.size())) != cudaSuccess) return e;
        if ((e = cudaEventRecord(fence_, s)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(copy_, fence_, 0)) != cudaSuccess) return e;
        const bool helpers = pieces.size() * kChunk >= kMinWorkers;
        for (size_t base = 0; base < pieces.size(); base += slots_) {
            const size_t cnt = std::min(slots_, pieces.size() - base);
            for (size_t i = 0; i < cnt; i++) {
                const Piece &p = pieces[base + i];
                if ((e = cudaMemcpyAsync(ring_ + i * kChunk, p.dev, p.len, cudaMemcpyDeviceToHost, copy_)) != cudaSuccess) return e;
                if ((e = cudaEventRecord(slot_ev_[i], copy_)) != cudaSuccess) return e;
            }
            std::atomic<int> err{0};
            run(cnt, helpers, [&](size_t i) {
                const Piece &p = pieces[base + i];
                if (cudaEventSynchronize(slot_ev_[i]) != cudaSuccess) err = 1;
                std::memcpy(p.host, ring_ + i * kChunk, p.len);
            });
            if (err) return cudaErrorUnknown;
        }
        return cudaStreamSynchronize(s);
    }
    cudaError_t d2h(void *host, const void *dev, size_t bytes, cudaStream_t s) {
        return d2h(std::vector<Seg>{{const_cast<void *>(dev), host, bytes}}, s);
    }

    // Wait until staged host-to-device copies issued so far have left the pinned ring.
    cudaError_t drain() {
        if (!busy_ || !copy_) return cudaSuccess;
        busy_ = false;
        return cudaStreamSynchronize(copy_);
    }

    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
        workers_.clear();
        stop_ = false;
        if (ring_) cudaFreeHost(ring_);
        ring_ = nullptr;
        slots_ = 0;
        for (auto e : slot_ev_) cudaEventDestroy(e);
        slot_ev_.clear();
        if (done_) cudaEventDestroy(done_);
        if (fence_) cudaEventDestroy(fence_);
        if (copy_) cudaStreamDestroy(copy_);
        done_ = fence_ = nullptr;
        copy_ = nullptr;
    }

  private:
    struct Piece {
        char *dev;
        char *host;
        size_t len;
    };
    // Segments that are small or pinned are copied directly on `s`; the rest is cut into ring-sized pieces.
    cudaError_t plan(const std::vector<Seg> &segs, cudaMemcpyKind kind, cudaStream_t s, std::vector<Piece> &pieces, size_t min_total) {
        size_t total = 0;
        for (const Seg &g : segs) total += g.bytes;
        for (const Seg &g : segs) {
            if (g.bytes == 0) continue;
            if (total < min_total || !staged(g.host, g.bytes)) {
                cudaError_t e = kind == cudaMemcpyHostToDevice ? cudaMemcpyAsync(g.dev, g.host, g.bytes, kind, s)
                                                               : cudaMemcpyAsync(g.host, g.dev, g.bytes, kind, s);
                if (e != cudaSuccess) return e;
                continue;
            }
            for (size_t off = 0; off < g.bytes; off += kChunk)
                pieces.push_back({(char *)g.dev + off, (char *)g.host + off, std::min(kChunk, g.bytes - off)});
        }
        return cudaSuccess;
    }
    bool staged(const void *host, size_t bytes) const {
        (void)bytes;
        if (nthreads_ <= 0) return false;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
            (void)cudaGetLastError();
            return true;
        }
        return a.type == cudaMemoryTypeUnregistered;  // pinned / managed memory goes to the DMA engine directly
    }
    // Pinning host memory costs ~0.4 ms per MiB, so a process that only ever moves a few MiB (one proof at k <= 14)
    // should not pay for the full ring: it starts at what the first transfer needs and doubles up to kSlots.
    cudaError_t ensure(size_t want) {
        cudaError_t e;
        if (ring_ && busy_ && (e = drain_keep()) != cudaSuccess) return e;
        size_t need = kMinSlots;
        while (need < want && need < kSlots) need *= 2;
        if (ring_ && need <= slots_) return cudaSuccess;
        if (!ring_) {
            int dev = 0;
            cudaGetDevice(&dev);
            device_ = dev;
            if ((e = cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&done_, cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&fence_, cudaEventDisableTiming)) != cudaSuccess) return e;
        } else {
            if ((e = cudaStreamSynchronize(copy_)) != cudaSuccess) return e;
            cudaFreeHost(ring_);
            ring_ = nullptr;
            slots_ = 0;
        }
        if ((e = cudaHostAlloc((void **)&ring_, kChunk * need, cudaHostAllocDefault)) != cudaSuccess) return e;
        while (slot_ev_.size() < need) {
            cudaEvent_t ev;
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
            slot_ev_.push_back(ev);
        }
        slots_ = need;
        return cudaSuccess;
    }
    // the ring still holds chunks of an earlier h2d whose DMAs may be in flight
    cudaError_t drain_keep() {
        busy_ = false;
        return cudaStreamSynchronize(copy_);
    }
    // run fn(0..count-1) on the calling thread, helped by the workers when `helpers`, and wait
    void run(size_t count, bool helpers, const std::function<void(size_t)> &fn) {
        if (!helpers) {
            for (size_t i = 0; i < count; i++) fn(i);
            return;
        }
        if (workers_.empty())
            for (int t = 0; t < nthreads_; t++) workers_.emplace_back([this] { loop(); });
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_ = &fn;
            next_ = 0;
            count_ = count;
            pending_ = count;
            epoch_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }
    void work() {
        for (;;) {
            size_t i;
            const std::function<void(size_t)> *fn;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (!job_ || next_ >= count_) return;
                i = next_++;
                fn = job_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> lk(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void loop() {
        cudaSetDevice(device_);
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && job_ && next_ < count_); });
                if (stop_) return;
                seen = epoch_;
            }
            work();
        }
    }

    int nthreads_;
    int device_ = 0;
    char *ring_ = nullptr;
    size_t slots_ = 0;
    cudaStream_t copy_ = nullptr;
    cudaEvent_t done_ = nullptr, fence_ = nullptr;
    std::vector<cudaEvent_t> slot_ev_;
    bool busy_ = false;
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)> *job_ = nullptr;
    size_t next_ = 0, count_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

}  // namespace h2b
