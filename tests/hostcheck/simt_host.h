// Host SIMT shim (TEST INFRASTRUCTURE): lets __global__ kernels of nova_b200/csrc run on the CPU with one
// std::thread per CUDA thread, `__shared__` variables shared between them, `__syncwarp()` / `__syncthreads()` as
// barriers and `__shfl_down_sync` as an exchange through per-warp slots -- so the lane roles, the shared-memory
// hand-offs, the warp / block reductions and the ordering of a kernel are exercised without a GPU (the
// arithmetic underneath is the bit-exact host emulation of field.cuh).  Blocks of a grid run one after the
// other.  Include BEFORE the headers under test.
#pragma once
#define NOVA_SIMT_HOST 1
#include <array>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct simt_dim3 { unsigned x = 0, y = 0, z = 0; };
inline thread_local simt_dim3 threadIdx, blockIdx;
inline simt_dim3 blockDim{32, 1, 1}, gridDim{1, 1, 1};

// sense-reversing barrier that spins with yield: the warp-parallel Keccak of the round kernels crosses thousands of
// barriers per call, and a mutex + condition variable costs ~50 us per crossing with 32 threads on a few cores
class simt_barrier {
 public:
  explicit simt_barrier(int n) : n_(n) {}
  void wait() {
    const int gen = gen_.load(std::memory_order_acquire);
    if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
      count_.store(0, std::memory_order_relaxed);
      gen_.store(gen + 1, std::memory_order_release);
    } else {
      while (gen_.load(std::memory_order_acquire) == gen) std::this_thread::yield();
    }
  }
 private:
  int n_;
  std::atomic<int> count_{0}, gen_{0};
};
inline simt_barrier*& simt_current_barrier() { static simt_barrier* b = nullptr; return b; }

// one running block: a block barrier, one barrier and one 32-slot exchange array per warp
struct simt_block {
  explicit simt_block(unsigned nthreads) : all((int)nthreads) {
    for (unsigned w = 0; w * 32 < nthreads; w++) {
      unsigned lanes = nthreads - w * 32 < 32 ? nthreads - w * 32 : 32;
      warps.emplace_back(new simt_barrier((int)lanes));
      slots.emplace_back();
    }
  }
  simt_barrier all;
  std::vector<std::unique_ptr<simt_barrier>> warps;
  std::vector<std::array<uint32_t, 32>> slots;
};
inline simt_block*& simt_current_block() { static simt_block* b = nullptr; return b; }

inline void __syncwarp() {
  if (simt_current_block()) simt_current_block()->warps[threadIdx.x >> 5]->wait();
  else simt_current_barrier()->wait();
}
inline void __syncthreads() { simt_current_block()->all.wait(); }
// every lane of the warp must call it (full mask), as in the kernels under test
inline uint32_t __shfl_down_sync(unsigned, uint32_t v, int delta) {
  simt_block* b = simt_current_block();
  unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  b->slots[w][l] = v;
  b->warps[w]->wait();
  uint32_t r = l + (unsigned)delta < 32 ? b->slots[w][l + delta] : v;
  b->warps[w]->wait();
  return r;
}
// arbitrary source lane (every lane of the warp must call it)
inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) {
  simt_block* b = simt_current_block();
  static std::array<uint32_t, 32> lone_slots;  // simt_launch_warp has no block context: one warp, one slot array
  std::array<uint32_t, 32>& slots = b ? b->slots[threadIdx.x >> 5] : lone_slots;
  unsigned l = threadIdx.x & 31;
  slots[l] = v;
  __syncwarp();
  uint32_t r = slots[(unsigned)src & 31u];
  __syncwarp();
  return r;
}
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicMin(uint32_t* p, uint32_t v) {
  uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return old;
}

#define __global__
#define __shared__ static
#define __launch_bounds__(...)

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

// run `kernel` as <<<grid, block>>> (1-D): the blocks one after the other, the threads of a block concurrently
inline void simt_launch_grid(unsigned grid, unsigned block, const std::function<void()>& kernel) {
  const simt_dim3 saved_block = blockDim, saved_grid = gridDim;
  blockDim = simt_dim3{block, 1, 1};
  gridDim = simt_dim3{grid, 1, 1};
  for (unsigned b = 0; b < grid; b++) {
    simt_block ctx(block);
    simt_current_block() = &ctx;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < block; t++)
      th.emplace_back([&, t, b] {
        threadIdx.x = t;
        blockIdx.x = b;
        kernel();
      });
    for (auto& x : th) x.join();
    simt_current_block() = nullptr;
  }
  blockDim = saved_block;
  gridDim = saved_grid;
}
