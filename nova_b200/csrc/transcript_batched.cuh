// Device-side round of a BATCHED sum-check: several claims, combined with fixed coefficients into one
// cubic per round -- RelaxedR1CSSNARK::prove_helper of the MicroSpartan prover
// (src/spartan/ppsnark.rs:886-983) with its three engines (MemorySumcheckInstance ppsnark.rs:497-670,
// InnerBatchedSumcheckInstance :677-786, WitnessBoundSumcheck :270-325).  Per round the reference
//   * asks every engine for its claims' evaluation points [s(0), lead, s(-1)]; the eq-weighted claims
//     derive them from (t(0), t(inf)) and their OWN running claim (sumcheck.rs:680-747),
//   * combines them with powers of one challenge, builds the cubic, absorbs, squeezes (ppsnark.rs:924-957),
//   * updates every running claim (update_claim, sumcheck.rs:68-75) and the eq instances' bound value.
// This header does that O(1) work on the device so that the nine reductions, this kernel and the
// sixteen binds of every round are enqueued back to back.  Built on transcript.cuh; all of it is
// `NOVA_HD` and runs on the CPU in tests/hostcheck.
#pragma once
#include "transcript.cuh"

namespace nova {

constexpr int SCB_MAX_CLAIMS = 16;
constexpr int SCB_MAX_EQ = 4;

enum scb_kind {
  SCB_RAW3 = 0,     // sums = [s(0), lead, s(-1)] used as they are           (compute_eval_points_cubic)
  SCB_LIN2 = 1,     // sums = [s(0), s(-1)], lead = 0                        (linear / quadratic forms)
  SCB_EQ_DEG2 = 2,  // sums = [t(0), t(inf)], eq-weighted, derived           (evaluation_points_cubic_*)
  SCB_EQ_DEG1 = 3,  // sums = [t(0)], eq-weighted, derived, lead = 0         (evaluation_points_quadratic_with_one_input)
};

// Filled by the host for every round (the slots change when a tau = 0 round needs third sums);
// mirrors b200_scb_desc in include/nova_b200.h.
struct scb_desc {
  int32_t nclaims, neq;
  int32_t kind[SCB_MAX_CLAIMS];
  int32_t slot[SCB_MAX_CLAIMS];     // element index of the claim's first sum in the sums buffer
  int32_t slot_m1[SCB_MAX_CLAIMS];  // element index of t(-1) when the claim's eq instance has tau = 0 this round, else -1
  int32_t eq_of[SCB_MAX_CLAIMS];    // eq instance of an EQ claim
  const void* tau[SCB_MAX_EQ];      // this round's tau of each eq instance (device, Montgomery)
  const void* tau_inv[SCB_MAX_EQ];  // and its inverse (unused when tau = 0)
};

// Device-resident state; mirrors b200_scb_state (1296 bytes).
struct scb_state {
  sc_state head;                   // head.claim = combined running claim e, transcript, rounds_done (head.q unused)
  fe_t coeff[SCB_MAX_CLAIMS];      // powers of the batching challenge (constant over the rounds)
  fe_t claim[SCB_MAX_CLAIMS];      // per-claim running claims (EQ kinds)
  fe_t q[SCB_MAX_EQ];              // eval_eq_left of each eq instance
};

// [s(0), lead, s(-1)] of claim i
template <class F>
NOVA_HD void scb_claim_evals(const scb_desc& d, int i, const scb_state& st, const fe_t* s, const fe_t* tm1,
                             const fe_t& tau, const fe_t& tau_inv, fe_t (&ev)[3]) {
  const int kind = d.kind[i];
  if (kind == SCB_RAW3) {
    ev[0] = s[0];
    ev[1] = s[1];
    ev[2] = s[2];
    return;
  }
  if (kind == SCB_LIN2) {
    ev[0] = s[0];
    ev[1] = fe_zero<F>();
    ev[2] = s[1];
    return;
  }
  const fe_t& q = st.q[d.eq_of[i]];
  const fe_t one = fe_one<F>();
  const fe_t two_tau = fe_dbl<F>(tau);
  const fe_t e0c = fe_sub<F>(one, tau);
  const fe_t em1c = fe_sub<F>(fe_dbl<F>(one), fe_add<F>(two_tau, tau));
  const fe_t qt0 = fe_mulx<F>(q, s[0]);
  const fe_t s0 = fe_mulx<F>(e0c, qt0);
  fe_t qtinf = fe_zero<F>();
  ev[0] = s0;
  ev[1] = fe_zero<F>();
  if (kind == SCB_EQ_DEG2) {
    qtinf = fe_mulx<F>(q, s[1]);
    ev[1] = fe_mulx<F>(fe_sub<F>(two_tau, one), qtinf);
  }
  if (tm1) {  // tau = 0: third sum (sumcheck.rs:1082-1213)
    ev[2] = fe_mulx<F>(em1c, fe_mulx<F>(q, *tm1));
  } else {    // q t(1) = (claim - s0) / tau ; t(-1) = 2 t(inf) + 2 t(0) - t(1)   (t(inf) = 0 for DEG1)
    fe_t qt1 = fe_mulx<F>(fe_sub<F>(st.claim[i], s0), tau_inv);
    ev[2] = fe_mulx<F>(em1c, fe_sub<F>(fe_dbl<F>(fe_add<F>(qtinf, qt0)), qt1));
  }
}

// component k (0: s(0), 1: lead, 2: s(-1)) of the combined evaluation points
template <class F>
NOVA_HD fe_t scb_combine(const scb_desc& d, const scb_state& st, const fe_t (*ev)[3], int k) {
  fe_t acc = fe_zero<F>();
  for (int i = 0; i < d.nclaims; i++) acc = fe_add<F>(acc, fe_mulx<F>(ev[i][k], st.coeff[i]));
  return acc;
}

// UniPoly::from_evals_deg3([c0, e - c0, cb, ci])  (univariate.rs:103-114)
template <class F>
NOVA_HD void scb_poly(const fe_t& e, const fe_t& c0, const fe_t& cb, const fe_t& ci, sc_round_poly& p) {
  fe_t e1 = fe_sub<F>(e, c0);
  fe_t b = fe_sub<F>(fe_half<F>(fe_add<F>(e1, ci)), c0);
  p.deg = 3;
  p.c[0] = c0;
  p.c[3] = cb;
  p.c[2] = b;
  p.c[1] = fe_sub<F>(fe_sub<F>(fe_sub<F>(e1, cb), c0), b);
  p.tau = fe_zero<F>();
}

// SumcheckProof::update_claim (sumcheck.rs:68-75) for evaluation points [e0, c3, em1]
template <class F>
NOVA_HD fe_t scb_update_claim(const fe_t& claim, const fe_t (&ev)[3], const fe_t& r) {
  fe_t e1 = fe_sub<F>(claim, ev[0]);
  fe_t a1 = fe_sub<F>(fe_half<F>(fe_sub<F>(e1, ev[2])), ev[1]);
  fe_t a2 = fe_sub<F>(fe_half<F>(fe_add<F>(e1, ev[2])), ev[0]);
  fe_t acc = fe_add<F>(fe_mulx<F>(ev[1], r), a2);
  acc = fe_add<F>(fe_mulx<F>(acc, r), a1);
  return fe_add<F>(fe_mulx<F>(acc, r), ev[0]);
}

// eval_eq_left *= 1 - tau - r + 2 r tau   (sumcheck.rs:1226-1231)
template <class F>
NOVA_HD fe_t scb_bound(const fe_t& q, const fe_t& tau, const fe_t& r) {
  fe_t f = fe_add<F>(fe_sub<F>(fe_sub<F>(fe_one<F>(), tau), r), fe_dbl<F>(fe_mulx<F>(r, tau)));
  return fe_mulx<F>(q, f);
}

#if defined(__CUDACC__) || defined(NOVA_SIMT_HOST)  // NOVA_SIMT_HOST: tests/hostcheck/simt_host.h
// One warp.  Lane i < nclaims: claim i's evaluation points and claim update; lanes 0..2: one component of
// the combination each; the whole warp: the two squeeze hashes; lanes g < neq: the eq bounds.
struct scb_round_smem {
  msg_buf msg;
  uint32_t flip_pos;
  fe_t ev[SCB_MAX_CLAIMS][3];
  fe_t comb[3];
  fe_t part[30];
  scb_state st;
};

template <class F>
NOVA_D void scb_round_warp(const scb_desc& d, scb_state* __restrict__ state, const void* __restrict__ sums,
                           const uint8_t* __restrict__ pending, uint32_t pending_len, uint8_t absorb_label,
                           uint8_t squeeze_label, void* __restrict__ out_poly, void* __restrict__ out_r,
                           scb_round_smem& sh) {
  const int lane = (int)(threadIdx.x & 31u);
  SCB_STAMP(0);
  static_assert(sizeof(scb_state) % 4 == 0, "word copy");
  for (unsigned w = lane; w < sizeof(scb_state) / 4; w += 32)
    reinterpret_cast<uint32_t*>(&sh.st)[w] = reinterpret_cast<const uint32_t*>(state)[w];
  __syncwarp();
  fe_t tau = fe_zero<F>(), tau_inv = fe_zero<F>();
  fe_t ev[3] = {fe_zero<F>(), fe_zero<F>(), fe_zero<F>()};
  const bool is_claim = lane < d.nclaims;
  const bool is_eq_claim = is_claim && d.kind[lane] >= SCB_EQ_DEG2;
  if (is_claim) {
    fe_t s[3];
    for (int k = 0; k < 3; k++) s[k] = fe_load_rw(sums, (size_t)d.slot[lane] + k);
    fe_t tm1 = fe_zero<F>();
    const bool has_m1 = is_eq_claim && d.slot_m1[lane] >= 0;
    if (is_eq_claim) {
      tau = fe_load_rw(d.tau[d.eq_of[lane]], 0);
      if (has_m1) tm1 = fe_load_rw(sums, (size_t)d.slot_m1[lane]);
      else tau_inv = fe_load_rw(d.tau_inv[d.eq_of[lane]], 0);
    }
    scb_claim_evals<F>(d, lane, sh.st, s, has_m1 ? &tm1 : nullptr, tau, tau_inv, ev);
    for (int k = 0; k < 3; k++) sh.ev[lane][k] = ev[k];
  }
  __syncwarp();
  SCB_STAMP(1);
  {  // combination sum_i coeff_i ev_i[k]: lane 3 j + k takes claims j and j + 10, lanes 0..2 add the ten partial sums
    static_assert(SCB_MAX_CLAIMS <= 20, "two claims per lane");
    const int k = lane % 3, j = lane / 3;
    fe_t part = fe_zero<F>();
    if (lane < 30) {
      if (j < d.nclaims) part = fe_mulx<F>(sh.ev[j][k], sh.st.coeff[j]);
      if (j + 10 < d.nclaims) part = fe_add<F>(part, fe_mulx<F>(sh.ev[j + 10][k], sh.st.coeff[j + 10]));
      sh.part[lane] = part;
    }
    __syncwarp();
    if (lane < 3) {
      fe_t acc = sh.part[lane];
      for (int q = 1; q < 10; q++) acc = fe_add<F>(acc, sh.part[3 * q + lane]);
      sh.comb[lane] = acc;
    }
  }
  __syncwarp();
  SCB_STAMP(2);
  sc_round_poly poly;
  scb_poly<F>(sh.st.head.claim, sh.comb[0], sh.comb[1], sh.comb[2], poly);
  fe_t canon[3];
  sc_round_compressed<F>(poly, canon);
  SCB_STAMP(3);
  if (lane == 0) {
    sh.flip_pos = sc_round_message(sh.msg, pending, pending_len, absorb_label, canon, 3, sh.st.head, squeeze_label);
    for (int k = 0; k < 3; k++) fe_store(out_poly, k, canon[k]);
  }
  __syncwarp();
  SCB_STAMP(4);
  uint64_t digest[8];
  {
    uint64_t d0[4], d1[4];
    keccak256_msg_warp_pair(sh.msg, sh.flip_pos, d0, d1);
    for (int i = 0; i < 4; i++) {
      digest[i] = d0[i];
      digest[4 + i] = d1[i];
    }
  }
  SCB_STAMP(5);
  sc_state head = sh.st.head;  // every lane computes the same r and e; lane 0 stores them
  fe_t r = sc_round_finish<F>(SC_ROUND_QUAD_PROD, head, poly, digest);
  __syncwarp();                // all reads of sh.st are done
  SCB_STAMP(6);
  if (is_eq_claim) state->claim[lane] = scb_update_claim<F>(sh.st.claim[lane], ev, r);
  if (lane < d.neq) {
    fe_t t = fe_load_rw(d.tau[lane], 0);
    state->q[lane] = scb_bound<F>(sh.st.q[lane], t, r);
  }
  if (lane == 0) {
    state->head = head;
    fe_store(out_r, 0, r);
  }
  SCB_STAMP(7);
}

// The round with the last stage of the reductions folded in: <<<1, 32 * nsums>>>.  Warp w adds the `nblocks` partial
// triples k_form_reduce_multi left for sum w (partials[(w * nblocks + b) * 3 + k]) -- all sums at once -- then warp 0
// runs the round on the totals.  One launch (and one dependent-launch gap) less per round than k_form_final_multi +
// k_sc_round_batched.
constexpr int SCB_FUSED_MAX_SUMS = 32;
template <class F>
__global__ void __launch_bounds__(32 * SCB_FUSED_MAX_SUMS) k_sc_round_batched_fused(
    scb_desc d, scb_state* __restrict__ state, const void* __restrict__ partials, int nblocks,
    const uint8_t* __restrict__ pending, uint32_t pending_len, uint8_t absorb_label, uint8_t squeeze_label,
    void* __restrict__ out_poly, void* __restrict__ out_r) {
  __shared__ scb_round_smem sh;
  __shared__ fe_t sums[3 * SCB_FUSED_MAX_SUMS];
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  fe_t acc[3] = {fe_zero<F>(), fe_zero<F>(), fe_zero<F>()};
  for (int b = lane; b < nblocks; b += 32)
#pragma unroll
    for (int k = 0; k < 3; k++) acc[k] = fe_add<F>(acc[k], fe_load_rw(partials, ((size_t)warp * nblocks + b) * 3 + k));
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int dlt = 16; dlt > 0; dlt >>= 1) {
      fe_t o;
#pragma unroll
      for (int i = 0; i < 8; i++) o.l[i] = __shfl_down_sync(0xffffffffu, acc[k].l[i], dlt);
      acc[k] = fe_add<F>(acc[k], o);
    }
    if (lane == 0) sums[3 * warp + k] = acc[k];
  }
  __syncthreads();
  if (warp == 0)
    scb_round_warp<F>(d, state, sums, pending, pending_len, absorb_label, squeeze_label, out_poly, out_r, sh);
}

// <<<1, 32>>>
template <class F>
__global__ void __launch_bounds__(32) k_sc_round_batched(scb_desc d, scb_state* __restrict__ state,
                                                         const void* __restrict__ sums,
                                                         const uint8_t* __restrict__ pending, uint32_t pending_len,
                                                         uint8_t absorb_label, uint8_t squeeze_label,
                                                         void* __restrict__ out_poly, void* __restrict__ out_r) {
  __shared__ scb_round_smem sh;
  scb_round_warp<F>(d, state, sums, pending, pending_len, absorb_label, squeeze_label, out_poly, out_r, sh);
}
#endif

}  // namespace nova
