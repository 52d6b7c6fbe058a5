// The short rounds of a batched sum-check in ONE kernel (prove_helper ppsnark.rs:886-983).
//
// Once the tables are a few hundred entries long a round is pure latency: two reduction launches, the round kernel and
// the bind launch cost ~55 us while touching a few KiB.  k_scb_tail runs every remaining round inside one CTA:
//     all sums of the round (one warp per sum, all sums at once)  ->  warp 0: scb_round_warp (evaluation points, combination,
//     cubic, Keccak absorb / squeeze, claim and eq-bound updates)  ->  all threads: bind every table with r
// with block barriers in between -- no launches, no host.  The tables stay in global memory (L1 / L2 resident at this
// size) and are read through the coherent path (sc_form<.., RW = true>) because the same kernel rewrites them.
#pragma once
#include "poly_kernels.cuh"
#include "transcript_batched.cuh"

namespace nova {

constexpr int SCB_TAIL_MAX_TABLES = 24;
constexpr int SCB_TAIL_THREADS = 512;

// eq tables of one EqSumCheckInstance (sumcheck.rs:606-664) as the C API lays them out: left[k] = eq(taus[fh-k .. fh)) at
// element offset 2^k - 1 of `left`, right[k] = eq(taus[l-k .. l)) at element offset 2^k - 1 of `right`
struct scb_tail_eq {
  const void* left;
  const void* right;
  const void* taus;      // l Montgomery elements (device)
  const void* tau_inv;   // their inverses (0 -> 0)
  uint64_t tau_zero;     // bit j: taus[j] == 0
};

struct scb_tail_args {
  scb_desc d;                       // nclaims, neq, kind, eq_of (slot / slot_m1 / tau pointers are set per round here)
  int32_t form[SCB_MAX_CLAIMS];     // sc_form_id of claim i's sums
  int32_t form_m1[SCB_MAX_CLAIMS];  // third-sum form for a round whose tau is 0, or -1
  int32_t tab[SCB_MAX_CLAIMS][3];   // table indices of A, B, C (-1: none)
  int32_t ntables;
  int32_t num_rounds;               // l: rounds of the whole sum-check
  int32_t first_round;              // 0-based round this kernel starts with; the tables hold 2^(l - first_round) entries
  void* tables[SCB_TAIL_MAX_TABLES];
  scb_tail_eq eq[SCB_MAX_EQ];
};

template <class F>
__global__ void __launch_bounds__(SCB_TAIL_THREADS) k_scb_tail(const scb_tail_args a, scb_state* __restrict__ state,
                                                                void* __restrict__ sums /* 6 * SCB_MAX_CLAIMS elements */,
                                                                const uint8_t* __restrict__ pending, uint32_t pending_len,
                                                                uint8_t absorb_label, uint8_t squeeze_label,
                                                                void* __restrict__ polys, void* __restrict__ rs) {
  __shared__ scb_round_smem sh;
  __shared__ scb_desc d;
  const int l = a.num_rounds, fh = l / 2, shh = l - fh;
  if (threadIdx.x == 0) d = a.d;
  __syncthreads();
  size_t len = (size_t)1 << (l - a.first_round);
  for (int j = a.first_round; j < l; j++, len >>= 1) {
    const size_t half = len >> 1;
    const int round = j + 1;  // EqSumCheckInstance::round (sumcheck.rs:1233-1251)
    // one warp per sum (claim i, or its third sum in a tau = 0 round): lane-strided partial sums, shuffle tree
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int t = warp; t < 2 * a.d.nclaims; t += nwarps) {
      const int i = t % a.d.nclaims, pass = t / a.d.nclaims;
      const int g = a.d.eq_of[i];
      multi_sum m;
      m.A = a.tab[i][0] >= 0 ? a.tables[a.tab[i][0]] : nullptr;
      m.B = a.tab[i][1] >= 0 ? a.tables[a.tab[i][1]] : nullptr;
      m.C = a.tab[i][2] >= 0 ? a.tables[a.tab[i][2]] : nullptr;
      m.eq_left = m.eq_right = nullptr;
      m.shift = 0;
      bool tau_zero = false;
      if (g >= 0) {
        const scb_tail_eq& q = a.eq[g];
        if (round < fh) {
          m.eq_left = (const char*)q.left + 32 * (((size_t)1 << (fh - round)) - 1);
          m.eq_right = (const char*)q.right + 32 * (((size_t)1 << shh) - 1);
          m.shift = shh;
        } else {
          m.eq_right = (const char*)q.right + 32 * (((size_t)1 << (l - round)) - 1);
        }
        tau_zero = (q.tau_zero >> j) & 1;
      }
      const bool third = tau_zero && a.form_m1[i] >= 0;
      if (pass == 1 && !third) continue;  // warp-uniform
      m.form = pass ? a.form_m1[i] : a.form[i];
      fe_t acc[3] = {fe_zero<F>(), fe_zero<F>(), fe_zero<F>()};
      multi_dispatch<F, true>(m, half, 1, 0, (size_t)lane, 32, acc);
#pragma unroll
      for (int k = 0; k < 3; k++)
#pragma unroll
        for (int dlt = 16; dlt > 0; dlt >>= 1) acc[k] = fe_add<F>(acc[k], fe_shfl_down(acc[k], dlt));
      if (lane == 0) {
        const int slot = 3 * (pass ? SCB_MAX_CLAIMS + i : i);
        for (int k = 0; k < 3; k++) fe_store(sums, (size_t)slot + k, acc[k]);
        if (pass == 0) {
          d.slot[i] = slot;
          if (!third) d.slot_m1[i] = -1;
        } else {
          d.slot_m1[i] = slot;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < (unsigned)a.d.neq) {
      d.tau[threadIdx.x] = (const char*)a.eq[threadIdx.x].taus + 32 * (size_t)j;
      d.tau_inv[threadIdx.x] = (const char*)a.eq[threadIdx.x].tau_inv + 32 * (size_t)j;
    }
    __syncthreads();
    void* r_j = (char*)rs + 32 * (size_t)j;
    if (threadIdx.x < 32)
      scb_round_warp<F>(d, state, sums, j == 0 ? pending : nullptr, j == 0 ? pending_len : 0u, absorb_label,
                        squeeze_label, (char*)polys + 96 * (size_t)j, r_j, sh);
    __syncthreads();
    const fe_t r = fe_load_rw(r_j, 0);
    for (int t = 0; t < a.ntables; t++) {
      void* z = a.tables[t];
      for (size_t i = threadIdx.x; i < half; i += blockDim.x) {
        fe_t lo = fe_load_rw(z, i), hi = fe_load_rw(z, i + half);
        fe_store(z, i, fe_add<F>(lo, fe_mul<F>(r, fe_sub<F>(hi, lo))));
      }
    }
    __syncthreads();
  }
}

}  // namespace nova
