// host_copy.h -- a small pool of threads that copy host memory to host memory.
//
// The host-pointer entry points (srt_render_xray, srt_render_shaded, ...) return a frame into whatever buffer the caller
// passed, like the reference's renderLoop fills its Image.  Into a PAGEABLE buffer the CUDA driver copies through its own
// pinned staging area with a single thread: 16 MB took 1.0 ms of a 1.2 ms call (profiles/r02_bench_n1_c3.json) against
// 0.3 ms into a page-locked buffer.  srt_api.cpp therefore lands the frame chunk by chunk in a pinned staging buffer of
// its own and hands every chunk that has arrived to this pool, so that several cores move the bytes to the caller's
// buffer while the DMA engine brings the next chunks.  No CUDA in this file: it is unit-tested without a GPU
// (srt_debug_host_copy, tests/test_host.py).
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace srt {

class HostCopyPool {
public:
    struct Job { void* dst; const void* src; size_t bytes; };

    // spin_us: how long an idle worker polls for the next chunk after activate() before it goes back to sleep
    explicit HostCopyPool(int n_workers, unsigned spin_us = 0) : spin_us_(spin_us)
    {
        if (n_workers < 0) n_workers = 0;
        workers_.reserve((size_t)n_workers);
        try {
            for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { work(); });
        } catch (...) {
            // fewer threads than asked for (resource limits): the pool works with what it got, down to the caller alone
        }
    }
    HostCopyPool(const HostCopyPool&) = delete;
    HostCopyPool& operator=(const HostCopyPool&) = delete;
    ~HostCopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
            active_.store(false, std::memory_order_release);
        }
        cv_work_.notify_all();
        for (auto& t : workers_) if (t.joinable()) t.join();
    }

    int workers() const { return (int)workers_.size(); }
    unsigned spin_us() const { return spin_us_; }

    // Chunks are about to arrive: wake the workers now and let them poll for work (for at most spin_us each) instead of
    // sleeping until the first submit -- a thread that has slept through a frame takes tens of microseconds to wake up,
    // a chunk of a frame lands every ~20.  drain() ends the activation.
    void activate()
    {
        if (!spin_us_ || workers_.empty()) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            ++epoch_;
            active_.store(true, std::memory_order_release);
        }
        cv_work_.notify_all();
    }

    // Queue one copy; returns at once.  The ranges of the jobs in flight must not overlap.
    void submit(void* dst, const void* src, size_t bytes)
    {
        if (!bytes) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(Job{ dst, src, bytes });
            ++pending_;
            queued_.store(q_.size(), std::memory_order_release);
        }
        cv_work_.notify_one();
    }

    // The calling thread copies queued jobs itself until the queue is empty, then waits for the jobs the workers still hold.
    // On return every job submitted so far has been carried out and its bytes are visible to the caller.
    void drain()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            if (!q_.empty()) {
                const Job j = q_.front();
                q_.pop_front();
                queued_.store(q_.size(), std::memory_order_release);
                lk.unlock();
                std::memcpy(j.dst, j.src, j.bytes);
                lk.lock();
                --pending_;
                continue;
            }
            if (pending_ == 0) break;
            cv_done_.wait(lk, [this] { return pending_ == 0 || !q_.empty(); });
        }
        active_.store(false, std::memory_order_release);
    }

private:
    static void cpu_relax()
    {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }

    // Poll (without the lock) until a job is queued or the activation ends; true = the budget ran out first.
    bool spin_for_work() const
    {
        const auto t0 = std::chrono::steady_clock::now();
        const auto budget = std::chrono::microseconds(spin_us_);
        for (unsigned i = 1;; ++i) {
            if (queued_.load(std::memory_order_acquire) != 0 || !active_.load(std::memory_order_acquire)) return false;
            cpu_relax();
            if ((i & 255u) == 0 && std::chrono::steady_clock::now() - t0 > budget) return true;
        }
    }

    void work()
    {
        std::unique_lock<std::mutex> lk(m_);
        uint64_t spent_epoch = 0;                       // the activation this worker has used its polling budget on
        for (;;) {
            if (!q_.empty()) {
                const Job j = q_.front();
                q_.pop_front();
                queued_.store(q_.size(), std::memory_order_release);
                lk.unlock();
                std::memcpy(j.dst, j.src, j.bytes);
                lk.lock();
                if (--pending_ == 0) cv_done_.notify_all();
                continue;
            }
            if (stop_) return;
            if (active_.load(std::memory_order_acquire) && spent_epoch != epoch_) {
                const uint64_t e = epoch_;
                lk.unlock();
                const bool timed_out = spin_for_work();
                lk.lock();
                if (timed_out) spent_epoch = e;
                continue;
            }
            cv_work_.wait(lk);                          // submit / activate / the destructor wake it; conditions re-checked above
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<Job> q_;
    size_t pending_ = 0;
    uint64_t epoch_ = 0;                                // activations so far (under m_)
    bool stop_ = false;
    std::atomic<bool> active_{ false };                 // written under m_, polled without it
    std::atomic<size_t> queued_{ 0 };                   // q_.size(), polled without the lock
    const unsigned spin_us_;
};

} // namespace srt
