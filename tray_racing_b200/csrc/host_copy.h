// host_copy.h — moving large HOST buffers of a caller (pageable memory, borrowed for the call) to the device at PCIe speed:
// persistent host copy threads stage the bytes into pinned slots, the copy engine takes them from there.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace tray {

// Host-side copy spread over a few persistent worker threads: one thread moves ~10 GB/s, the PCIe link of a B200 takes ~50.
class CopyPool {
public:
    static CopyPool& get() { static CopyPool p; return p; }
    void copy(void* dst, const void* src, size_t bytes) {
        const size_t nt = bytes < (2u << 20) ? 1 : workers_.size() + 1;
        if (nt <= 1) { memcpy(dst, src, bytes); return; }
        std::lock_guard<std::mutex> call(call_mu_);                  // one copy at a time (scenes on several host threads)
        const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> g(mu_);
            dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; per_ = per;
            pending_ = (int)workers_.size(); generation_++;
        }
        cv_.notify_all();
        memcpy(dst, src, per < bytes ? per : bytes);                 // part 0 on the calling thread
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [&] { return pending_ == 0; });
    }
private:
    CopyPool() {
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned n = hw >= 16 ? 7 : hw >= 4 ? hw / 2 - 1 : 0;
        if (const char* e = getenv("TRAY_CUDA_COPY_THREADS")) {          // total threads taking part in a copy (caller included)
            const int v = atoi(e);
            if (v >= 1 && v <= 64) n = (unsigned)v - 1;
        }
        for (unsigned i = 0; i < n; i++) workers_.emplace_back([this, i] { run(i + 1); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void run(size_t part) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> g(mu_);
            cv_.wait(g, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            char* d = dst_; const char* s = src_; const size_t bytes = bytes_, per = per_;
            g.unlock();
            const size_t a = part * per;
            if (a < bytes) memcpy(d + a, s + a, a + per <= bytes ? per : bytes - a);
            g.lock();
            if (--pending_ == 0) done_cv_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_cv_;
    char* dst_ = nullptr; const char* src_ = nullptr; size_t bytes_ = 0, per_ = 0;
    int pending_ = 0; unsigned long long generation_ = 0; bool stop_ = false;
};
inline void par_copy(void* dst, const void* src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }


// cudaMemcpyAsync(H2D) for big pageable sources: 2 pinned slots of 32 MiB, the host copy of chunk i+1 overlaps the DMA of
// chunk i.  Stream-ordered like cudaMemcpyAsync; returns once the last chunk has been ENQUEUED (the source may then be
// reused — its bytes are in the pinned slots or on the device).
inline cudaError_t upload_pipelined(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    constexpr size_t SLOT = 32u << 20;
    if (bytes < (8u << 20)) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
    static std::mutex mu;
    static void* slot[2] = { nullptr, nullptr };                       // pinned host memory is visible to every device (UVA)
    static cudaEvent_t evs[64][2] = {};                                // events are per device
    std::lock_guard<std::mutex> g(mu);
    cudaError_t e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
    cudaEvent_t* ev = evs[dev];
    for (int i = 0; i < 2; i++) {
        if (!slot[i] && cudaMallocHost(&slot[i], SLOT) != cudaSuccess) {
            slot[i] = nullptr; cudaGetLastError();
            return cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st);
        }
        if (!ev[i]) {
            if ((e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventRecord(ev[i], st)) != cudaSuccess) return e;
        }
    }
    // a slot may still be feeding a DMA of ANOTHER device: wait for every device's last use of it
    for (int d = 0; d < 64; d++)
        for (int i = 0; i < 2; i++)
            if (d != dev && evs[d][i] && (e = cudaEventSynchronize(evs[d][i])) != cudaSuccess) return e;
    size_t off = 0;
    for (int i = 0; off < bytes; i ^= 1) {
        const size_t len = bytes - off < SLOT ? bytes - off : SLOT;
        if ((e = cudaEventSynchronize(ev[i])) != cudaSuccess) return e;       // the slot's previous DMA is done
        par_copy(slot[i], (const char*)h_src + off, len);
        if ((e = cudaMemcpyAsync((char*)d_dst + off, slot[i], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(ev[i], st)) != cudaSuccess) return e;
        off += len;
    }
    return cudaSuccess;
}

}  // namespace tray
