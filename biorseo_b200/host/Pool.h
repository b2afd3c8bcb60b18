// Task pool with the reference's interface (cppsrc/Pool.h:9-25): push / done /
// infinite_loop_func. The GPU search does not use it; it is kept so that host code written
// against the reference keeps compiling, and for host-side jobs such as parsing library
// files in parallel.
#ifndef BIORSEO_B200_POOL_H_
#define BIORSEO_B200_POOL_H_

#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>

class Pool
{
  public:
    Pool() : open_(true) {}
    ~Pool() {}
    std::mutex& get_mutex_reference(void) { return lock_; }
    void        push(std::function<void()> func);   // enqueue one job
    void        done();                             // no more jobs: workers drain the queue and return
    void        infinite_loop_func();               // worker body

  private:
    std::deque<std::function<void()>> jobs_;
    std::mutex                        lock_;
    std::condition_variable           wake_;
    bool                              open_;
};

#endif   // BIORSEO_B200_POOL_H_
