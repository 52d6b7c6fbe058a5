// Instantiates every kernel for one field and fills the launcher table (see ops.cuh).
#pragma once
#include "ops.cuh"
#include "msm_kernels.cuh"
#include "field_kernels.cuh"

namespace nova {

inline int stream_grid(size_t n, int block, int waves = 8) {
  // grid = multiple of the SM count (148), capped by the work available
  size_t need = (n + block - 1) / block;
  size_t cap = (size_t)148 * waves;
  size_t g = need < cap ? need : cap;
  return (int)(g == 0 ? 1 : g);
}

// sum of n points (optionally gathered through idx): used by the 0/1-scalar and sparse paths
// (msm.rs:432-454 accumulate_bases, msm.rs:689-708 batch_add).  Two-level: each thread sums a
// strided slice with mixed adds, then a single block tree-sums the partials.
template <class F>
__global__ void __launch_bounds__(128) k_sum_points1(const void* __restrict__ tables,
                                                     const uint32_t* __restrict__ idx, size_t n,
                                                     void* __restrict__ partial) {
  size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t nthreads = (size_t)gridDim.x * blockDim.x;
  xyzz_t acc = xyzz_identity<F>();
  for (size_t i = tid; i < n; i += nthreads) {
    affine_t p = affine_load(tables, idx ? idx[i] : i);
    if (!affine_is_identity(p)) xyzz_madd<F>(acc, p.x, p.y);
  }
  xyzz_store(partial, tid, acc);
}

template <class F>
struct ops_impl {
  static void digits(cudaStream_t s, const void* scalars, const msm_plan& p) {
    int block = 256;
    int grid = (int)((p.n + block - 1) / block);
    k_digits<F><<<grid, block, 0, s>>>(scalars, p.n, p.c, p.W, p.G, p.B, p.digits, p.counts);
  }
  static void expand_key(cudaStream_t s, void* tables, size_t n_ck, int ntables, int shift) {
    int block = 128;
    int grid = (int)((n_ck + block - 1) / block);
    k_expand_key<F><<<grid, block, 0, s>>>(tables, n_ck, ntables, shift);
  }
  static void accumulate(cudaStream_t s, const void* tables, const msm_plan& p) {
    uint32_t K = (uint32_t)p.G * p.B;
    size_t max_entries = p.n * (size_t)p.W;
    size_t nseg = (max_entries + p.L - 1) / p.L;
    int block = 128;
    int grid = (int)((nseg + block - 1) / block);
    k_accumulate<F><<<grid, block, 0, s>>>(p.entries, p.start, K, tables, p.L, p.buckets, p.parts,
                                           p.pkeys);
  }
  static void fixup(cudaStream_t s, const msm_plan& p) {
    uint32_t K = (uint32_t)p.G * p.B;
    k_fixup<F><<<(K + 127) / 128, 128, 0, s>>>(p.start, K, p.L, p.heavy_min, p.parts, p.pkeys,
                                               p.buckets);
    k_fixup_heavy<F><<<148, 256, 0, s>>>(p.start, p.L, p.heavy, p.parts, p.pkeys, p.buckets);
  }
  static void index_bases(cudaStream_t s, void* bases, size_t n, const void* gen, uint64_t k0) {
    k_index_bases<F><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(bases, n, gen, k0);
  }
  static void jacobian_sum(cudaStream_t s, const void* pts, int k, void* out_jac) {
    k_jacobian_sum<F><<<1, 32, 0, s>>>(pts, k, out_jac);
  }
  static void reduce(cudaStream_t s, const msm_plan& p, void* out_jac) {
    uint32_t T = p.B / p.m;
    int block = 128;
    int grid = (int)(((size_t)T * p.G + block - 1) / block);
    k_reduce1<F><<<grid, block, 0, s>>>(p.start, p.B, p.G, p.m, p.buckets, p.rparts);
    k_reduce2<F><<<1, 256, 0, s>>>(p.rparts, T, p.G, p.c, out_jac);
  }
  static void sum_points(cudaStream_t s, const void* tables, const uint32_t* idx, size_t n,
                         void* scratch, void* out_jac) {
    // scratch must hold SUM_THREADS xyzz
    int block = 128, grid = 148;
    k_sum_points1<F><<<grid, block, 0, s>>>(tables, idx, n, scratch);
    k_reduce2<F><<<1, 256, 0, s>>>(scratch, (uint32_t)(grid * block), 1, 0, out_jac);
  }
  static void cross_term(cudaStream_t s, const void* az, const void* bz, const void* cz,
                         const void* e1, const void* e2, const void* u, size_t n, void* t) {
    k_cross_term<F><<<stream_grid(n, 256), 256, 0, s>>>(az, bz, cz, e1, e2, u, n, t);
  }
  static void axpy(cudaStream_t s, const void* a, const void* b, const void* r, size_t n,
                   void* out) {
    k_axpy<F><<<stream_grid(n, 256), 256, 0, s>>>(a, b, r, n, out);
  }
  static void vec_add(cudaStream_t s, const void* a, const void* b, size_t n, void* out) {
    k_vec_add<F><<<stream_grid(n, 256), 256, 0, s>>>(a, b, n, out);
  }
  static void bind_top(cudaStream_t s, void* z, size_t n, const void* r) {
    k_bind_top<F><<<stream_grid(n / 2, 256), 256, 0, s>>>(z, n / 2, r);
  }
  static constexpr field_ops table() {
    return field_ops{F::ID,  digits,       expand_key, accumulate, fixup,   reduce,
                     sum_points, jacobian_sum, index_bases, cross_term, axpy,       vec_add, bind_top};
  }
};

}  // namespace nova
