// Host-memory plumbing of the host entry points (no device code here):
//   * prefault_for_write  -- fault in the pages of a fresh result buffer while the
//                            GPU is still computing
//   * upload              -- host -> device copies that stage PAGEABLE sources
//                            through a small pinned pool on several threads
// Everything lives in an anonymous namespace: this header is included by
// lic_api.cu only.
#pragma once

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <sys/mman.h>

namespace {

// Fault in the pages of a host buffer that is about to be overwritten by a
// device-to-host copy.  A freshly allocated NumPy result has no physical pages
// yet; letting the copy engine's staging path fault them in one by one costs
// ~12 ms per 64 MiB, while doing it here overlaps with the passes still running
// on the GPU.  Best effort: hints only, the contents are never read.
void prefault_for_write(void *ptr, size_t bytes)
{
    if (bytes < (size_t)1 << 20)
        return;
    const uintptr_t page = 4096, huge = (uintptr_t)2 << 20;
    const uintptr_t begin = (uintptr_t)ptr, end = begin + bytes;
#ifdef MADV_HUGEPAGE
    const uintptr_t hb = (begin + huge - 1) & ~(huge - 1), he = end & ~(huge - 1);
    if (he > hb)
        madvise((void *)hb, he - hb, MADV_HUGEPAGE);
#endif
    const uintptr_t pb = begin & ~(page - 1), pe = (end + page - 1) & ~(page - 1);
    const unsigned nthreads = bytes >= ((size_t)16 << 20) ? 4 : 1;
    auto populate = [=](uintptr_t a, uintptr_t b) {
#ifdef MADV_POPULATE_WRITE
        if (madvise((void *)a, b - a, MADV_POPULATE_WRITE) == 0)
            return;
#endif
        for (uintptr_t q = a < begin ? begin : a; q < b && q < end; q += page)
            *(volatile char *)q = 0;   // we own every byte of [begin, end)
    };
    std::vector<std::thread> helpers;
    const uintptr_t chunk = (((pe - pb) / nthreads) + huge - 1) & ~(huge - 1);
    for (unsigned t = 1; t < nthreads; ++t) {
        const uintptr_t a = pb + t * chunk, b = std::min(pe, a + chunk);
        if (a < b)
            helpers.emplace_back(populate, a, b);
    }
    populate(pb, std::min(pe, pb + chunk));
    for (auto &h : helpers)
        h.join();
}

// ---------------------------------------------------------------------------
// Host -> device copies from PAGEABLE memory (ordinary NumPy arrays).
//
// cudaMemcpyAsync from pageable memory is staged by the driver through one
// bounce buffer on the calling thread (~10-12 GB/s measured here).  Pinned
// sources go straight to the copy engine.  For pageable sources we do the
// staging ourselves: a few worker threads memcpy 1 MiB chunks into a small pool
// of pinned buffers and enqueue each chunk's DMA as soon as it is filled, so
// the memcpy of one chunk overlaps the DMA of the others.
class Workers {
public:
    // runs fn(i) for i in [0, n) on the pool (the caller takes part) and returns when all are done
    void parallel_for(size_t n, const std::function<void(size_t)> &fn)
    {
        if (n == 0)
            return;
        std::lock_guard<std::mutex> one_job(job_mu_);   // jobs from concurrent callers take turns
        start_threads();
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            n_ = n;
            next_.store(0);
            pending_ = threads_.size();
            ++generation_;
        }
        cv_.notify_all();
        run_items(fn, n);
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }
private:
    void run_items(const std::function<void(size_t)> &fn, size_t n)
    {
        for (size_t i = next_.fetch_add(1); i < n; i = next_.fetch_add(1))
            fn(i);
    }
    void start_threads()
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (!threads_.empty())
            return;
        // the caller takes part too: up to eight threads copy (one thread moves ~10 GB/s, the host
        // link takes ~55)
        unsigned hw = std::thread::hardware_concurrency();
        unsigned count = std::min(7u, hw > 2 ? hw / 2 - 1 : 1u);
        for (unsigned t = 0; t < count; ++t)
            threads_.emplace_back([this] { loop(); }).detach();
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)> *fn;
            size_t n;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                fn = fn_;
                n = n_;
            }
            run_items(*fn, n);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0)
                    done_cv_.notify_all();
            }
        }
    }
    std::mutex job_mu_, mu_;
    std::condition_variable cv_, done_cv_;
    std::vector<std::thread> threads_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    std::atomic<size_t> next_{0};
};

Workers &workers()
{
    static Workers *w = new Workers;   // lives for the process: worker threads are detached
    return *w;
}

class PinnedPool {
public:
    static constexpr size_t kChunk = (size_t)1 << 20;   // small enough that one band of one array (4 MiB at
    static constexpr int kSlots = 48;                    // 4096^2 / 16 bands) keeps several threads busy
    struct Slot {
        void *host = nullptr;
        cudaEvent_t drained = nullptr;   // recorded after the DMA that reads this slot
        std::mutex mu;
    };
    // next slot in rotation, locked, its previous DMA complete; nullptr if pinned memory is unavailable
    Slot *acquire()
    {
        Slot &s = slots_[cursor_.fetch_add(1) % kSlots];
        s.mu.lock();
        if (!s.host) {
            // the event belongs to the device that is current now: a pool serves ONE device
            // (pinned_pool(device)), because an event cannot be recorded on another device's stream
            if (cudaHostAlloc(&s.host, kChunk, cudaHostAllocPortable) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.drained, cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                if (s.host) { cudaFreeHost(s.host); s.host = nullptr; }
                s.mu.unlock();
                return nullptr;
            }
        } else {
            cudaEventSynchronize(s.drained);
        }
        return &s;
    }

private:
    Slot slots_[kSlots];
    std::atomic<unsigned> cursor_{0};
};

// One pool per device ordinal (created on first use, lives for the process): a slot's
// `drained` event is tied to the device it was created on, and the batch entry point
// runs lanes for several devices at once.
PinnedPool &pinned_pool(int device)
{
    static std::mutex mu;
    static std::vector<PinnedPool *> pools;
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0)
        device = 0;
    if ((size_t)device >= pools.size())
        pools.resize((size_t)device + 1, nullptr);
    if (!pools[(size_t)device])
        pools[(size_t)device] = new PinnedPool;
    return *pools[(size_t)device];
}

bool is_pageable(const void *ptr)
{
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// Copy up to three host ranges to the device on `stream`, in one parallel sweep.
struct HostToDevice { void *dst; const void *src; size_t bytes; };

cudaError_t upload(const HostToDevice *jobs, int njobs, cudaStream_t stream)
{
    struct Piece { char *dst; const char *src; size_t bytes; };
    std::vector<Piece> pieces;
    for (int k = 0; k < njobs; ++k) {
        const HostToDevice &j = jobs[k];
        if (j.bytes == 0)
            continue;
        if (j.bytes < ((size_t)1 << 20) || !is_pageable(j.src)) {
            cudaError_t e = cudaMemcpyAsync(j.dst, j.src, j.bytes, cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess)
                return e;
            continue;
        }
        for (size_t off = 0; off < j.bytes; off += PinnedPool::kChunk)
            pieces.push_back({(char *)j.dst + off, (const char *)j.src + off,
                              std::min(PinnedPool::kChunk, j.bytes - off)});
    }
    if (pieces.empty())
        return cudaSuccess;
    int device = 0;
    cudaGetDevice(&device);
    std::atomic<int> failed{(int)cudaSuccess};
    PinnedPool &pool = pinned_pool(device);
    workers().parallel_for(pieces.size(), [&](size_t i) {
        cudaSetDevice(device);   // worker threads start on device 0
        const Piece &p = pieces[i];
        cudaError_t e;
        if (PinnedPool::Slot *slot = pool.acquire()) {
            std::memcpy(slot->host, p.src, p.bytes);
            e = cudaMemcpyAsync(p.dst, slot->host, p.bytes, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess)
                e = cudaEventRecord(slot->drained, stream);
            if (e != cudaSuccess)
                cudaStreamSynchronize(stream);   // never hand the slot on with its DMA unfenced
            slot->mu.unlock();
        } else {
            e = cudaMemcpyAsync(p.dst, p.src, p.bytes, cudaMemcpyHostToDevice, stream);
        }
        if (e != cudaSuccess)
            failed.store((int)e);
    });
    return (cudaError_t)failed.load();
}

// ---------------------------------------------------------------------------
// Page-locked result buffers.  A fresh NumPy result has no physical pages and is
// pageable: faulting it in and letting the driver bounce the download through
// its own staging costs ~4 ms per 64 MiB.  The Python binding may instead ask
// for a page-locked block, wrap it in the ndarray it returns, and give it back
// when that array is garbage collected.  Blocks are cached by size class and the
// total (cached + handed out) is capped; past the cap the caller falls back to
// ordinary memory.  Page-locking is itself expensive (~0.4 ms per MiB), far more
// than it saves on one download, so a block is only allocated once a size class
// has been asked for before: one-shot calls stay on ordinary memory, repeated
// workloads get the fast path from their second call on.
class ResultBlocks {
public:
    static constexpr size_t kCap = (size_t)4 << 30;   // bytes of page-locked memory this pool may hold
    static constexpr size_t kLargest = (size_t)256 << 20;   // bigger results stay on ordinary memory
    void *acquire(size_t bytes)
    {
        if (bytes > kLargest)
            return nullptr;   // page-locking would cost more than many downloads save
        const size_t cls = size_class(bytes);
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < free_.size(); ++i)
            if (free_[i].second == cls) {
                void *p = free_[i].first;
                free_.erase(free_.begin() + (long)i);
                return p;
            }
        if (++requests_[cls] < 2)
            return nullptr;
        // make room by dropping cached blocks of other sizes, oldest first
        while (total_ + cls > kCap && !free_.empty()) {
            void *gone = free_.front().first;
            cudaFreeHost(gone);
            total_ -= free_.front().second;
            free_.erase(free_.begin());
            // forget its size class too: the allocator may hand the address out again for another
            sizes_.erase(std::remove_if(sizes_.begin(), sizes_.end(),
                                        [gone](const std::pair<void *, size_t> &e) { return e.first == gone; }),
                         sizes_.end());
        }
        if (total_ + cls > kCap)
            return nullptr;
        void *p = nullptr;
        if (cudaHostAlloc(&p, cls, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        total_ += cls;
        sizes_.push_back({p, cls});
        return p;
    }
    void release(void *p)
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (auto &e : sizes_)
            if (e.first == p) {
                free_.push_back(e);
                return;
            }
    }

private:
    static size_t size_class(size_t bytes)
    {
        const size_t mib = (size_t)1 << 20;
        return std::max(mib, (bytes + mib - 1) / mib * mib);
    }
    std::mutex mu_;
    std::vector<std::pair<void *, size_t>> free_, sizes_;
    std::map<size_t, int> requests_;
    size_t total_ = 0;
};

ResultBlocks &result_blocks()
{
    static ResultBlocks *r = new ResultBlocks;
    return *r;
}

}  // namespace
