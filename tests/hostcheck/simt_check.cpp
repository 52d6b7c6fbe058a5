// The REAL kernel wrappers k_sc_round / k_sc_round_batched (nova_b200/csrc/transcript*.cuh) run on the CPU
// through the SIMT shim: 32 threads = 32 lanes.  Same C signatures as the sequential hooks in hostcheck.cpp
// (hc_sc_round, hc_sc_round_batched), so the tests can drive complete proofs through either.
#include <cstring>
#include "simt_host.h"
#include "../../nova_b200/csrc/transcript_batched.cuh"
using namespace nova;

extern "C" int hc_simt_sc_round(int fid, int kind, void* state144, const void* res, const void* tau, const void* tau_inv,
                                const void* pending, uint32_t pending_len, int absorb_label, int squeeze_label,
                                void* out_poly, void* out_r) {
  alignas(16) sc_state st;
  memcpy(&st, state144, 144);
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_warp([&] {
      k_sc_round<F>(kind, &st, res, tau, tau_inv, (const uint8_t*)pending, pending_len, (uint8_t)absorb_label,
                    (uint8_t)squeeze_label, out_poly, out_r);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 1: run(BN254_FQ{}); break;
    case 2: run(PALLAS_FP{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  memcpy(state144, &st, 144);
  return 0;
}

extern "C" int hc_simt_sc_round_batched(int fid, const void* desc, void* state, const void* sums, const void* pending,
                                        uint32_t pending_len, int absorb_label, int squeeze_label, void* out_poly,
                                        void* out_r) {
  const scb_desc d = *(const scb_desc*)desc;
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_warp([&] {
      k_sc_round_batched<F>(d, (scb_state*)state, sums, (const uint8_t*)pending, pending_len, (uint8_t)absorb_label,
                            (uint8_t)squeeze_label, out_poly, out_r);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 1: run(BN254_FQ{}); break;
    case 2: run(PALLAS_FP{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

// ---- the two-stage reduction kernels of the sum-check forms (poly_kernels.cuh: k_form_reduce with every
// sc_form, then k_form_final) as <<<grid, 256>>> blocks of host threads: grid-stride loops, warp shuffles,
// the shared-memory stage and the eq factor with its shard mapping run as written.  Mirrors
// ops_impl.cuh sc_launch / capi_poly.inc b200_sc_eval_sharded_dev. ----
#include "../../nova_b200/csrc/poly_kernels.cuh"

template <class F, int FORM>
static void simt_sc_form(const void* A, const void* B, const void* C, size_t len, const void* eq_left,
                         const void* eq_right, int shift, size_t id_mul, size_t id_add, unsigned grid, void* out) {
  constexpr int NOUT = sc_form_nout(FORM);
  sc_form<F, FORM> f;
  f.A = A;
  f.B = B;
  f.C = C;
  f.h = len / 2;
  f.eq.left = eq_left;
  f.eq.right = eq_right;
  f.eq.shift = shift;
  f.eq.mask = ((size_t)1 << shift) - 1;
  f.eq.id_mul = id_mul;
  f.eq.id_add = id_add;
  const size_t count = (FORM == SC_DOT_EQ || FORM == SC_DOT) ? len : len / 2;
  std::vector<fe_t> partials((size_t)grid * NOUT);
  simt_launch_grid(grid, 256, [&] { k_form_reduce<F, NOUT, sc_form<F, FORM>>(f, count, partials.data()); });
  simt_launch_grid(1, 256, [&] { k_form_final<F, NOUT>(partials.data(), (int)grid, out); });
}

template <class F>
static int simt_sc_eval_t(int form, const void* A, const void* B, const void* C, size_t len, const void* l,
                          const void* r, int shift, size_t id_mul, size_t id_add, unsigned grid, void* out) {
#define SC_CASE(X) \
  case X: simt_sc_form<F, X>(A, B, C, len, l, r, shift, id_mul, id_add, grid, out); return 0
  switch (form) {
    SC_CASE(SC_QUAD_PROD);
    SC_CASE(SC_LINEAR);
    SC_CASE(SC_QUADRATIC);
    SC_CASE(SC_CUBIC);
    SC_CASE(SC_EQ_CUBIC3);
    SC_CASE(SC_EQ_CUBIC2);
    SC_CASE(SC_EQ_QUAD1);
    SC_CASE(SC_EQ_CUBIC3_M1);
    SC_CASE(SC_EQ_CUBIC2_M1);
    SC_CASE(SC_EQ_QUAD1_M1);
    SC_CASE(SC_DOT_EQ);
    SC_CASE(SC_DOT);
    default: return 1;
  }
#undef SC_CASE
}

extern "C" int hc_simt_sc_eval(int fid, int form, const void* A, const void* B, const void* C, size_t len,
                               const void* eq_left, const void* eq_right, int shift, size_t id_mul, size_t id_add,
                               unsigned grid, void* out) {
  switch (fid) {
    case 0: return simt_sc_eval_t<BN254_FR>(form, A, B, C, len, eq_left, eq_right, shift, id_mul, id_add, grid, out);
    case 3: return simt_sc_eval_t<PALLAS_FQ>(form, A, B, C, len, eq_left, eq_right, shift, id_mul, id_add, grid, out);
    default: return 1;
  }
}

// the segmented reduction of the eq-weighted forms (k_form_reduce_eqseg), launched as ops_impl.cuh sc_launch does
// when NOVA_B200_SC_SEG=1: one block per segment of 2^shift indices (at most `grid` blocks)
template <class F, int FORM>
static void simt_sc_form_seg(const void* A, const void* B, const void* C, size_t len, const void* eq_left,
                             const void* eq_right, int shift, unsigned grid, void* out) {
  constexpr int NOUT = sc_form_nout(FORM);
  sc_form<F, FORM> f;
  f.A = A;
  f.B = B;
  f.C = C;
  f.h = len / 2;
  f.eq.left = eq_left;
  f.eq.right = eq_right;
  f.eq.shift = shift;
  f.eq.mask = ((size_t)1 << shift) - 1;
  f.eq.id_mul = 1;
  f.eq.id_add = 0;
  const size_t count = FORM == SC_DOT_EQ ? len : len / 2;
  std::vector<fe_t> partials((size_t)grid * NOUT);
  simt_launch_grid(grid, 256, [&] { k_form_reduce_eqseg<F, NOUT, sc_form<F, FORM>>(f, count, partials.data()); });
  simt_launch_grid(1, 256, [&] { k_form_final<F, NOUT>(partials.data(), (int)grid, out); });
}

extern "C" int hc_simt_sc_eval_seg(int fid, int form, const void* A, const void* B, const void* C, size_t len,
                                   const void* eq_left, const void* eq_right, int shift, unsigned grid, void* out) {
  if (fid != 0 && fid != 3) return 1;
#define SEG_CASE(X)                                                                                   \
  case X:                                                                                             \
    if (fid == 0) simt_sc_form_seg<BN254_FR, X>(A, B, C, len, eq_left, eq_right, shift, grid, out);   \
    else simt_sc_form_seg<PALLAS_FQ, X>(A, B, C, len, eq_left, eq_right, shift, grid, out);          \
    return 0
  switch (form) {
    SEG_CASE(SC_EQ_CUBIC3);
    SEG_CASE(SC_EQ_CUBIC2);
    SEG_CASE(SC_EQ_QUAD1);
    SEG_CASE(SC_EQ_CUBIC3_M1);
    SEG_CASE(SC_EQ_CUBIC2_M1);
    SEG_CASE(SC_EQ_QUAD1_M1);
    SEG_CASE(SC_DOT_EQ);
    default: return 1;
  }
#undef SEG_CASE
}

// ---- all sums of a batched round in one launch (k_form_reduce_multi + k_form_final_multi) and the short rounds of a
// batched sum-check inside one CTA (k_scb_tail, sumcheck_tail.cuh): the kernels as written, grid.y emulated by a loop ----
#include "../../nova_b200/csrc/sumcheck_tail.cuh"

extern "C" int hc_simt_sc_reduce_multi(int fid, const void* args, unsigned grid, void* out) {
  const multi_args a = *(const multi_args*)args;
  auto run = [&](auto tag) {
    using F = decltype(tag);
    std::vector<fe_t> partials((size_t)a.n * grid * 3);
    for (int y = 0; y < a.n; y++)
      simt_launch_grid(grid, 256, [&] {
        blockIdx.y = (unsigned)y;
        k_form_reduce_multi<F>(a, partials.data());
      });
    for (int y = 0; y < a.n; y++)
      simt_launch_grid(1, 256, [&] {
        blockIdx.y = (unsigned)y;
        k_form_final_multi<F>(partials.data(), (int)grid, out);
      });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

extern "C" int hc_simt_scb_tail(int fid, const void* args, void* state, void* sums, const void* pending,
                                uint32_t pending_len, int absorb_label, int squeeze_label, void* polys, void* rs) {
  const scb_tail_args a = *(const scb_tail_args*)args;
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_grid(1, SCB_TAIL_THREADS, [&] {
      k_scb_tail<F>(a, (scb_state*)state, sums, (const uint8_t*)pending, pending_len, (uint8_t)absorb_label,
                    (uint8_t)squeeze_label, polys, rs);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

// k_form_reduce_multi followed by k_sc_round_batched_fused (the last reduction stage inside the round kernel: one warp
// per sum, then warp 0 runs the round) -- what b200_sumcheck_batched launches per round
extern "C" int hc_simt_round_fused(int fid, const void* args, unsigned grid, const void* desc, void* state,
                                   const void* pending, uint32_t pending_len, int absorb_label, int squeeze_label,
                                   void* out_poly, void* out_r) {
  const multi_args a = *(const multi_args*)args;
  const scb_desc d = *(const scb_desc*)desc;
  auto run = [&](auto tag) {
    using F = decltype(tag);
    std::vector<fe_t> partials((size_t)a.n * grid * 3);
    for (int y = 0; y < a.n; y++)
      simt_launch_grid(grid, 256, [&] {
        blockIdx.y = (unsigned)y;
        k_form_reduce_multi<F>(a, partials.data());
      });
    simt_launch_grid(1, 32 * (unsigned)a.n, [&] {
      k_sc_round_batched_fused<F>(d, (scb_state*)state, partials.data(), (int)grid, (const uint8_t*)pending, pending_len,
                                  (uint8_t)absorb_label, (uint8_t)squeeze_label, out_poly, out_r);
    });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

// nested eq tables in one block of 1024 host threads (k_eq_prefix_tables)
extern "C" int hc_simt_eq_prefix(int fid, const void* taus, int hi, int K, void* out) {
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_grid(1, 1024, [&] { k_eq_prefix_tables<F>(taus, hi, K, out); });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

// several short polynomials at three points in one launch (k_poly_eval_small_multi): one block of 256 host threads each
extern "C" int hc_simt_poly_eval_small_multi(int fid, const void* const* polys, const size_t* lens, int k, const void* us,
                                             void* evals) {
  poly_multi_args a;
  a.k = k;
  for (int i = 0; i < k; i++) {
    a.p[i] = polys[i];
    a.len[i] = lens[i];
    a.out_index[i] = i;
  }
  auto run = [&](auto tag) {
    using F = decltype(tag);
    simt_launch_grid((unsigned)k, 256, [&] { k_poly_eval_small_multi<F, 3>(a, us, evals); });
  };
  switch (fid) {
    case 0: run(BN254_FR{}); break;
    case 3: run(PALLAS_FQ{}); break;
    default: return 1;
  }
  return 0;
}

extern "C" int hc_simt_sizes(int which) {
  switch (which) {
    case 0: return (int)sizeof(multi_args);
    case 1: return (int)sizeof(multi_sum);
    case 2: return (int)sizeof(scb_tail_args);
    case 3: return (int)sizeof(scb_desc);
    case 4: return (int)sizeof(scb_state);
    default: return -1;
  }
}
