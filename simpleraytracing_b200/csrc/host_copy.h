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

#include <condition_variable>
#include <cstddef>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace srt {

class HostCopyPool {
public:
    struct Job { void* dst; const void* src; size_t bytes; };

    explicit HostCopyPool(int n_workers)
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
        }
        cv_work_.notify_all();
        for (auto& t : workers_) if (t.joinable()) t.join();
    }

    int workers() const { return (int)workers_.size(); }

    // Queue one copy; returns at once.  The ranges of the jobs in flight must not overlap.
    void submit(void* dst, const void* src, size_t bytes)
    {
        if (!bytes) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            q_.push_back(Job{ dst, src, bytes });
            ++pending_;
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
                lk.unlock();
                std::memcpy(j.dst, j.src, j.bytes);
                lk.lock();
                --pending_;
                continue;
            }
            if (pending_ == 0) return;
            cv_done_.wait(lk, [this] { return pending_ == 0 || !q_.empty(); });
        }
    }

private:
    void work()
    {
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            cv_work_.wait(lk, [this] { return stop_ || !q_.empty(); });
            if (q_.empty()) {
                if (stop_) return;
                continue;
            }
            const Job j = q_.front();
            q_.pop_front();
            lk.unlock();
            std::memcpy(j.dst, j.src, j.bytes);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }

    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<Job> q_;
    size_t pending_ = 0;
    bool stop_ = false;
};

} // namespace srt
