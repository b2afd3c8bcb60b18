#include "Pool.h"

void Pool::push(std::function<void()> func)
{
    {
        std::lock_guard<std::mutex> g(lock_);
        jobs_.push_back(std::move(func));
    }
    wake_.notify_one();
}

void Pool::done()
{
    {
        std::lock_guard<std::mutex> g(lock_);
        open_ = false;
    }
    wake_.notify_all();
}

void Pool::infinite_loop_func()
{
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> g(lock_);
            wake_.wait(g, [this] { return !jobs_.empty() || !open_; });
            if (jobs_.empty()) return;   // closed and drained
            job = std::move(jobs_.front());
            jobs_.pop_front();
        }
        job();
    }
}
