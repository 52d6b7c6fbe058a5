// Poseidon random oracle on the device: the sponge `PoseidonRO::squeeze` runs
// (src/provider/poseidon.rs:41-127 over the vendored neptune sponge, src/frontend/gadgets/poseidon/) -- the hash
// NIFS::prove draws its folding challenge from (src/nova/nifs.rs:47-63).  With it the step
//     commit_T -> absorb(comm_T) -> r = squeeze -> fold W, E
// can be enqueued without a host round trip for r.
//
// One block, `t` warps (t = arity + 1 <= 25): the HADES permutation in its plain schedule -- per round: add the
// round constants, x^5 on all elements (full rounds) or on element 0 (partial rounds), then the dense MDS product.
//   warp 0            S-box layer (lane i = state element i), result to shared memory
//   warp j, lane i    state[i] * M[i][j], then a shuffle tree over the lanes: warp j owns output element j
// The reference runs the algebraically identical "optimized static" schedule (compressed constants, sparse
// partial-round matrices, poseidon_inner.rs:300-342); constants arrive precomputed from the host mirror
// (nova_b200/poseidon.py), in Montgomery form.
// Sponge (sponge/api.rs:205-243, Simplex mode): element 0 = the IO-pattern tag, elements are ADDED into the rate
// slots, a permutation whenever the rate is full, one more before the squeeze; output = rate element 0.
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace nova {

struct poseidon_desc {
  int t;    // width = arity + 1
  int r_f;  // full rounds
  int r_p;  // partial rounds
};
constexpr int POSEIDON_MAX_T = 32;

#if defined(__CUDACC__)
template <class F>
NOVA_D fe_t fe_pow5(const fe_t& x) {
  fe_t x2 = fe_sqr<F>(x);
  fe_t x4 = fe_sqr<F>(x2);
  return fe_mul<F>(x4, x);
}

template <class F>
NOVA_D void poseidon_permute(const poseidon_desc& d, const void* __restrict__ rc, const void* __restrict__ mds,
                             fe_t* state, fe_t* tmp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = d.r_f / 2;
  for (int r = 0; r < d.r_f + d.r_p; r++) {
    const bool full = r < half || r >= half + d.r_p;
    if (warp == 0 && lane < d.t) {
      fe_t x = fe_add<F>(state[lane], fe_load(rc, (size_t)r * d.t + lane));
      if (full || lane == 0) x = fe_pow5<F>(x);
      tmp[lane] = x;
    }
    __syncthreads();
    fe_t acc = fe_zero<F>();
    if (warp < d.t && lane < d.t) acc = fe_mul<F>(tmp[lane], fe_load(mds, (size_t)lane * d.t + warp));
    if (warp < d.t) {
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        fe_t o;
#pragma unroll
        for (int l = 0; l < 8; l++) o.l[l] = __shfl_xor_sync(0xffffffffu, acc.l[l], s);
        acc = fe_add<F>(acc, o);
      }
      if (lane == 0) state[warp] = acc;  // the S-box layer already read the old state (barrier above)
    }
    __syncthreads();
  }
}

// elems: n Montgomery field elements; tag: the IO-pattern tag as a CANONICAL 256-bit integer (host-computed, < 2^128);
// out[0] = hash (Montgomery), out[1] = challenge (Montgomery, same field), out[2] = challenge as a canonical integer
// (for a caller that needs it in ANOTHER field: base_as_scalar, traits.rs).  blockDim = 32 * t.
template <class F>
__global__ void __launch_bounds__(1024) k_poseidon_ro(const poseidon_desc d, const void* __restrict__ rc,
                                                      const void* __restrict__ mds, const void* __restrict__ elems,
                                                      uint32_t n, const void* __restrict__ tag_canonical, int num_bits,
                                                      int start_with_one, void* __restrict__ out) {
  __shared__ fe_t state[POSEIDON_MAX_T], tmp[POSEIDON_MAX_T];
  const int rate = d.t - 1;
  if (threadIdx.x < POSEIDON_MAX_T) state[threadIdx.x] = fe_zero<F>();
  __syncthreads();
  if (threadIdx.x == 0) state[0] = fe_to_mont<F>(fe_load_rw(tag_canonical, 0));
  __syncthreads();
  uint32_t done = 0;
  while (done < n) {
    const uint32_t take = n - done < (uint32_t)rate ? n - done : (uint32_t)rate;
    if (done) poseidon_permute<F>(d, rc, mds, state, tmp);  // the rate was full: permute before absorbing more
    if (threadIdx.x < take) state[1 + threadIdx.x] = fe_add<F>(state[1 + threadIdx.x], fe_load_rw(elems, done + threadIdx.x));
    __syncthreads();
    done += take;
  }
  poseidon_permute<F>(d, rc, mds, state, tmp);
  if (threadIdx.x == 0) {
    const fe_t h = state[1];
    fe_t c = fe_from_mont<F>(h);
    // keep the low num_bits bits (poseidon.rs:108-125); optionally force bit num_bits - 1
#pragma unroll
    for (int l = 0; l < 8; l++) {
      const int lo = 32 * l;
      if (num_bits <= lo) c.l[l] = 0;
      else if (num_bits < lo + 32) c.l[l] &= (1u << (num_bits - lo)) - 1u;
    }
    if (start_with_one && num_bits > 0) c.l[(num_bits - 1) >> 5] |= 1u << ((num_bits - 1) & 31);
    fe_store(out, 0, h);
    fe_store(out, 1, fe_to_mont<F>(c));
    fe_store(out, 2, c);
  }
}

// out[i] = in[i] * R mod p: canonical integers (each < p) -> Montgomery form of THIS field (the challenge of an RO over
// the other curve's field enters the folds this way: base_as_scalar, src/traits/mod.rs)
template <class F>
__global__ void __launch_bounds__(256) k_to_mont(const void* __restrict__ in, size_t n, void* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_store(out, i, fe_to_mont<F>(fe_load_rw(in, i)));
}
#endif

}  // namespace nova
