// pfv_pool.h — a small fixed pool of host threads (entropy coding in pfv_codec.cpp, dense -> token compaction in
// pfv_ctx.cu).  Host only.
#pragma once
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace pfv {

class Pool {
public:
    explicit Pool(unsigned n)
    {
        if (n == 0) n = 1;
        for (unsigned i = 0; i < n; i++) th_.emplace_back([this] { run(); });
    }
    ~Pool()
    {
        { std::lock_guard<std::mutex> l(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : th_) t.join();
    }
    void post(std::function<void()> f)
    {
        { std::lock_guard<std::mutex> l(m_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
private:
    void run()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

// run f(0..n-1) on the pool and wait for all of them
inline void parallel_for(Pool &pool, unsigned n, const std::function<void(unsigned)> &f)
{
    if (n == 0) return;
    std::mutex m;
    std::condition_variable cv;
    unsigned left = n;
    for (unsigned i = 0; i < n; i++)
        pool.post([&, i] {
            f(i);
            std::lock_guard<std::mutex> l(m);
            if (--left == 0) cv.notify_all();
        });
    std::unique_lock<std::mutex> l(m);
    cv.wait(l, [&] { return left == 0; });
}

}  // namespace pfv
