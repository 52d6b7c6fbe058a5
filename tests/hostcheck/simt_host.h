// Host SIMT shim (TEST INFRASTRUCTURE): lets the one-warp __global__ wrappers of
// nova_b200/csrc/transcript*.cuh run on the CPU as 32 std::threads, one per lane, with `__shared__`
// variables shared between them and `__syncwarp()` as a barrier -- so the lane roles, the shared-memory
// hand-offs and the ordering of a kernel wrapper are exercised without a GPU (the arithmetic underneath is
// the bit-exact host emulation of field.cuh).  Include BEFORE the headers under test.
#pragma once
#define NOVA_SIMT_HOST 1
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct simt_dim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local simt_dim3 threadIdx, blockIdx;
inline simt_dim3 blockDim{32, 1, 1}, gridDim{1, 1, 1};

class simt_barrier {
 public:
  explicit simt_barrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      gen_++;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};
inline simt_barrier*& simt_current_barrier() { static simt_barrier* b = nullptr; return b; }
inline void __syncwarp() { simt_current_barrier()->wait(); }

#define __global__
#define __shared__ static
#define __launch_bounds__(n)

// run `kernel` once per lane of one 32-lane warp (grid <<<1, 32>>>)
inline void simt_launch_warp(const std::function<void()>& kernel) {
  simt_barrier bar(32);
  simt_current_barrier() = &bar;
  std::vector<std::thread> th;
  for (unsigned lane = 0; lane < 32; lane++)
    th.emplace_back([&, lane] {
      threadIdx.x = lane;
      kernel();
    });
  for (auto& t : th) t.join();
  simt_current_barrier() = nullptr;
}
