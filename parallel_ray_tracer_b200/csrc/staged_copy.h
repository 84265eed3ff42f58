// staged_copy.h — host <-> device copies of large PAGEABLE arrays at PCIe rate.
//
// The reference uploads its scene with synchronous cudaMemcpy from malloc'ed memory (gpu/src/gpu.cu:143-175).  From
// pageable memory the driver stages through its own small pinned buffer on the calling thread: 1.7 GB/s measured for the
// 1.8 GB of raw triangles of the 50 M-triangle scene (profiles/r01_s2_five_configs_b200.jsonl).  Here the copy is staged
// through a ring of two 64 MB page-locked buffers owned by the library: several host threads memcpy chunk k+1 into one
// buffer while the DMA engine moves chunk k out of the other, one cudaMemcpyAsync per chunk.  Arrays below 4 MB, and the
// case where pinned memory cannot be had, take the plain path.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace rt {

struct StagingRing {
    static constexpr size_t kChunk = 64u << 20;
    void* buf[2] = {nullptr, nullptr};
    std::mutex mu;
    bool tried = false, ok = false;

    bool ready()
    {
        if (tried) return ok;
        tried = true;
        for (int i = 0; i < 2; i++) {
            if (cudaHostAlloc(&buf[i], kChunk, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return false; }
        }
        ok = true;
        return true;
    }
    static StagingRing& get() { static StagingRing r; return r; }
};

// two events on the CURRENT device (an event can only be recorded on a stream of the device it was created on)
struct EventPair {
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaError_t init()
    {
        for (int i = 0; i < 2; i++) {
            cudaError_t e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    ~EventPair() { for (int i = 0; i < 2; i++) if (ev[i]) cudaEventDestroy(ev[i]); }
};

inline void parallel_memcpy(void* dst, const void* src, size_t bytes)
{
    static const int hw = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    const int t = (int)std::min<size_t>((size_t)hw, bytes >> 22);
    if (t <= 1) { std::memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < t; i++) {
        const size_t lo = bytes * i / t, hi = bytes * (i + 1) / t;
        th.emplace_back([=] { std::memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    for (auto& x : th) x.join();
}

// Host (pageable) -> device, ordered on `st`.  Returns when the source may be reused; the device copy completes on `st`.
inline cudaError_t staged_h2d(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    StagingRing& R = StagingRing::get();
    std::unique_lock<std::mutex> lock(R.mu);
    if (bytes < (4u << 20) || !R.ready()) {
        lock.unlock();
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st); // pageable source: do not return before it has been read
    }
    EventPair P;
    cudaError_t e = P.init();
    if (e != cudaSuccess) return e;
    int k = 0;
    for (size_t off = 0; off < bytes; off += StagingRing::kChunk, k ^= 1) {
        const size_t len = std::min(StagingRing::kChunk, bytes - off);
        if ((e = cudaEventSynchronize(P.ev[k])) != cudaSuccess) return e; // the DMA out of this buffer (two chunks ago) is done
        parallel_memcpy(R.buf[k], (const char*)src + off, len);
        if ((e = cudaMemcpyAsync((char*)dst + off, R.buf[k], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(P.ev[k], st)) != cudaSuccess) return e;
    }
    for (int i = 0; i < 2; i++)
        if ((e = cudaEventSynchronize(P.ev[i])) != cudaSuccess) return e; // the ring is free for the next caller
    return cudaSuccess;
}

// Device -> host (pageable), blocking; `st` must have no later work the caller cares to overlap.
inline cudaError_t staged_d2h(void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    StagingRing& R = StagingRing::get();
    std::unique_lock<std::mutex> lock(R.mu);
    if (bytes < (4u << 20) || !R.ready()) {
        lock.unlock();
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    EventPair P;
    cudaError_t e = P.init();
    if (e != cudaSuccess) return e;
    const size_t n_chunks = (bytes + StagingRing::kChunk - 1) / StagingRing::kChunk;
    auto len_of = [&](size_t c) { return std::min(StagingRing::kChunk, bytes - c * StagingRing::kChunk); };
    if ((e = cudaMemcpyAsync(R.buf[0], src, len_of(0), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(P.ev[0], st)) != cudaSuccess) return e;
    for (size_t c = 0; c < n_chunks; c++) {
        const int k = (int)(c & 1);
        if (c + 1 < n_chunks) { // next chunk into the other buffer while this one is copied out on the host
            if ((e = cudaMemcpyAsync(R.buf[k ^ 1], (const char*)src + (c + 1) * StagingRing::kChunk, len_of(c + 1), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(P.ev[k ^ 1], st)) != cudaSuccess) return e;
        }
        if ((e = cudaEventSynchronize(P.ev[k])) != cudaSuccess) return e;
        parallel_memcpy((char*)dst + c * StagingRing::kChunk, R.buf[k], len_of(c));
    }
    return cudaSuccess;
}

} // namespace rt
