// Streaming prime-field vector kernels (K4 of SURVEY.md §2).  All HBM-bound: one element per
// thread-iteration, 2 x 128-bit loads per operand, grid-stride over a grid sized in
// multiples of the SM count.
//
//   k_cross_term  T = Az o Bz - u*Cz - E1 (- E2)      src/r1cs/mod.rs:614-620, 650-657
//   k_axpy        out = a + r*b                        src/r1cs/mod.rs:1044-1073 (W, E folds)
//   k_vec_add     out = a + b                          src/r1cs/mod.rs:589-609 (Z = Z1 + Z2)
//   k_bind_top    Z[i] += r*(Z[i+n/2] - Z[i])          src/spartan/polys/multilinear.rs:65-84
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace nova {

template <class F>
__global__ void __launch_bounds__(256) k_cross_term(const void* __restrict__ az,
                                                    const void* __restrict__ bz,
                                                    const void* __restrict__ cz,
                                                    const void* __restrict__ e1,
                                                    const void* __restrict__ e2,
                                                    const void* __restrict__ u_ptr, size_t n,
                                                    void* __restrict__ t) {
  const fe_t u = fe_load(u_ptr, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fe_t a = fe_load(az, i), b = fe_load(bz, i), c = fe_load(cz, i), e = fe_load(e1, i);
    fe_t r = fe_sub<F>(fe_sub<F>(fe_mul<F>(a, b), fe_mul<F>(u, c)), e);
    if (e2 != nullptr) r = fe_sub<F>(r, fe_load(e2, i));
    fe_store(t, i, r);
  }
}

template <class F>
__global__ void __launch_bounds__(256) k_axpy(const void* __restrict__ a,
                                              const void* __restrict__ b,
                                              const void* __restrict__ r_ptr, size_t n,
                                              void* __restrict__ out) {
  const fe_t r = fe_load(r_ptr, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fe_t x = fe_load(a, i), y = fe_load(b, i);
    fe_store(out, i, fe_add<F>(x, fe_mul<F>(r, y)));
  }
}

template <class F>
__global__ void __launch_bounds__(256) k_vec_add(const void* __restrict__ a,
                                                 const void* __restrict__ b, size_t n,
                                                 void* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    fe_store(out, i, fe_add<F>(fe_load(a, i), fe_load(b, i)));
  }
}

// in place on the low half; the caller truncates to n/2 (multilinear.rs:82)
template <class F>
__global__ void __launch_bounds__(256) k_bind_top(void* z, size_t half,
                                                  const void* __restrict__ r_ptr) {
  const fe_t r = fe_load(r_ptr, 0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half;
       i += (size_t)gridDim.x * blockDim.x) {
    fe_t lo = fe_load_rw(z, i), hi = fe_load_rw(z, i + half);
    fe_store(z, i, fe_add<F>(lo, fe_mul<F>(r, fe_sub<F>(hi, lo))));
  }
}

}  // namespace nova
