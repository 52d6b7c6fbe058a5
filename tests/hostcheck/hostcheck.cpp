// Host-side build of the device field/curve templates (the non-__CUDA_ARCH__ branches of
// nova_b200/csrc/field.cuh emulate every PTX carry chain bit-exactly).  Test-only: lets the
// CPU test-suite validate the Montgomery/XYZZ algorithm structure without a GPU.
#include <cstring>
#include "../../nova_b200/csrc/curve.cuh"
using namespace nova;

template <class F>
static void fe_op(int op, const fe_t& a, const fe_t& b, fe_t& r) {
  switch (op) {
    case 0: r = fe_add<F>(a, b); break;
    case 1: r = fe_sub<F>(a, b); break;
    case 2: r = fe_mul_chain<F>(a, b); break;
    case 3: r = fe_inv<F>(a); break;
    case 4: r = fe_to_mont<F>(a); break;
    case 5: r = fe_from_mont<F>(a); break;
    case 6: r = fe_neg<F>(a); break;
    case 7: r = fe_mul_cs<F>(a, b); break;
    case 8: r = fe_sqr_dedicated<F>(a); break;
  }
}

// sum of n affine points via madd (mode 0), or via pairwise xyzz_add of singletons (mode 1),
// or [k]P via xyzz_mul_small of point 0 with k = n (mode 2) -> Jacobian 96 B
template <class F>
static void pt_sum(int mode, const affine_t* pts, size_t n, fe_t* out) {
  xyzz_t acc = xyzz_identity<F>();
  if (mode == 0) {
    for (size_t i = 0; i < n; i++)
      if (!affine_is_identity(pts[i])) xyzz_madd<F>(acc, pts[i].x, pts[i].y);
  } else if (mode == 1) {
    for (size_t i = 0; i < n; i++) {
      xyzz_t q = xyzz_identity<F>();
      if (!affine_is_identity(pts[i])) xyzz_madd<F>(q, pts[i].x, pts[i].y);
      xyzz_add<F>(acc, q);
    }
  } else {
    xyzz_t q = xyzz_identity<F>();
    if (!affine_is_identity(pts[0])) xyzz_madd<F>(q, pts[0].x, pts[0].y);
    acc = xyzz_mul_small<F>(q, (uint32_t)n);
  }
  xyzz_to_jacobian<F>(acc, out[0], out[1], out[2]);
}

extern "C" {
int hc_fe_op(int fid, int op, const void* a, const void* b, void* out, size_t n) {
  const fe_t* A = (const fe_t*)a;
  const fe_t* B = (const fe_t*)b;
  fe_t* R = (fe_t*)out;
  for (size_t i = 0; i < n; i++) {
    switch (fid) {
      case 0: fe_op<BN254_FR>(op, A[i], B[i], R[i]); break;
      case 1: fe_op<BN254_FQ>(op, A[i], B[i], R[i]); break;
      case 2: fe_op<PALLAS_FP>(op, A[i], B[i], R[i]); break;
      case 3: fe_op<PALLAS_FQ>(op, A[i], B[i], R[i]); break;
      default: return 1;
    }
  }
  return 0;
}
int hc_pt_sum(int fid, int mode, const void* pts, size_t n, void* out) {
  switch (fid) {
    case 0: pt_sum<BN254_FR>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 1: pt_sum<BN254_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 2: pt_sum<PALLAS_FP>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 3: pt_sum<PALLAS_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    default: return 1;
  }
  return 0;
}
}

// ---- 9 x 29-bit carry-free backend (field29.cuh / curve29.cuh), same sources as the device ----
#include "../../nova_b200/csrc/curve29.cuh"

template <class F>
static void fe29_op(int op, const fe_t& a, const fe_t& b, fe_t& r) {
  fe29_t x = f29_from_std<F>(a), y = f29_from_std<F>(b), z;
  switch (op) {
    case 0: z = f29_add(x, y); f29_carry(z); break;
    case 1: z = f29_sub<F, 2>(x, y); break;
    case 2: z = f29_mul<F>(x, y); break;
    case 3: z = pa29<F>::inv(x); break;
    case 6: z = f29_neg<F, 2>(x); break;
    case 7: z = f29_sqr<F>(x); break;
    case 8: {  // lazy chain: ((x + y) * (x - y + 4p)) with un-normalised sum, then squared
      fe29_t s = f29_add(x, y), d = f29_sub<F, 4>(x, y);
      z = f29_sqr<F>(f29_mul<F>(s, d));
      break;
    }
    case 9: {  // zero test: returns 1 iff x == y mod p
      fe29_t d = f29_sub<F, 8>(x, y);
      r = fe_zero<F>();
      r.l[0] = f29_is_zero_modp<F>(d) ? 1 : 0;
      return;
    }
    default: z = x;
  }
  r = f29_to_std<F>(z);
}

template <class F>
static void pt29_sum(int mode, const affine_t* pts, size_t n, fe_t* out) {
  using PA = pa29<F>;
  typename PA::pt acc = PA::identity();
  auto conv = [](const affine_t& a) { return PA::from_std_affine(a.x, a.y); };
  if (mode == 0) {
    for (size_t i = 0; i < n; i++)
      if (!affine_is_identity(pts[i])) PA::madd(acc, conv(pts[i]));
  } else if (mode == 1) {
    for (size_t i = 0; i < n; i++) {
      typename PA::pt q = PA::identity();
      if (!affine_is_identity(pts[i])) PA::madd(q, conv(pts[i]));
      PA::add(acc, q);
    }
  } else if (mode == 2) {
    typename PA::pt q = PA::identity();
    if (!affine_is_identity(pts[0])) PA::madd(q, conv(pts[0]));
    acc = PA::mul_small(q, (uint32_t)n);
  } else if (mode == 3) {  // through the table format and negation, then to_affine -> store/load raw
    for (size_t i = 0; i < n; i++) {
      if (affine_is_identity(pts[i])) continue;
      typename PA::aff a = conv(pts[i]);
      fe_t rx = f29_store_raw<F>(a.x), ry = f29_store_raw<F>(a.y);
      typename PA::aff b;
      b.x = f29_load_raw(rx);
      b.y = f29_load_raw(ry);
      if (i & 1) { PA::neg_aff(b); PA::neg_aff(b); }  // double negation (y -> 2p - y twice)
      PA::madd(acc, b);
    }
    if (!PA::is_identity(acc)) {
      typename PA::aff t = PA::to_affine(acc);
      typename PA::pt q = PA::identity();
      PA::madd(q, t);
      acc = q;
    }
  }
  PA::to_jacobian_std(acc, out[0], out[1], out[2]);
}

extern "C" {
int hc_fe29_op(int fid, int op, const void* a, const void* b, void* out, size_t n) {
  const fe_t* A = (const fe_t*)a;
  const fe_t* B = (const fe_t*)b;
  fe_t* R = (fe_t*)out;
  for (size_t i = 0; i < n; i++) {
    switch (fid) {
      case 0: fe29_op<BN254_FR>(op, A[i], B[i], R[i]); break;
      case 1: fe29_op<BN254_FQ>(op, A[i], B[i], R[i]); break;
      case 2: fe29_op<PALLAS_FP>(op, A[i], B[i], R[i]); break;
      case 3: fe29_op<PALLAS_FQ>(op, A[i], B[i], R[i]); break;
      default: return 1;
    }
  }
  return 0;
}
int hc_pt29_sum(int fid, int mode, const void* pts, size_t n, void* out) {
  switch (fid) {
    case 0: pt29_sum<BN254_FR>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 1: pt29_sum<BN254_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 2: pt29_sum<PALLAS_FP>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 3: pt29_sum<PALLAS_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    default: return 1;
  }
  return 0;
}
}

// ---- quad-cooperative XYZZ ops (coop.cuh): 4 std::threads per quad, barrier-based exchange ----
#include <condition_variable>
#include <mutex>
#include <thread>
#include "../../nova_b200/csrc/coop.cuh"

namespace {
struct host_barrier {
  std::mutex m;
  std::condition_variable cv;
  int count = 0, gen = 0;
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    int g = gen;
    if (++count == 4) { count = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g; });
  }
};
struct quad_shared { fe_t slot[4]; host_barrier bar; };
struct quad_comm_host {
  quad_shared* sh;
  int q;
  int lane() const { return q; }
  fe_t get(const fe_t& v, int src) const {
    sh->slot[q] = v;
    sh->bar.wait();
    fe_t r = sh->slot[src];
    sh->bar.wait();
    return r;
  }
};
}  // namespace

// mode 0: sum of points with coop_add (pairs, incl. P+P and P-P); mode 1: [2^k]P with coop_dbl
template <class F>
static void coop_run(int mode, const affine_t* pts, size_t n, fe_t* out) {
  quad_shared sh;
  xyzz_t results[4];
  std::thread th[4];
  for (int q = 0; q < 4; q++)
    th[q] = std::thread([&, q] {
      quad_comm_host cm{&sh, q};
      xyzz_t acc = xyzz_identity<F>();
      if (mode == 0) {
        for (size_t i = 0; i < n; i++) {
          xyzz_t p = xyzz_identity<F>();
          if (!affine_is_identity(pts[i])) xyzz_madd<F>(p, pts[i].x, pts[i].y);
          coop_add<F>(acc, p, cm);
        }
      } else {
        xyzz_madd<F>(acc, pts[0].x, pts[0].y);
        for (size_t i = 0; i < n; i++) coop_dbl<F>(acc, cm);
      }
      results[q] = acc;
    });
  for (auto& t : th) t.join();
  for (int q = 1; q < 4; q++)  // all lanes must agree bit for bit
    if (memcmp(&results[q], &results[0], sizeof(xyzz_t)) != 0) { memset(out, 0xff, 96); return; }
  xyzz_to_jacobian<F>(results[0], out[0], out[1], out[2]);
}
extern "C" int hc_coop(int fid, int mode, const void* pts, size_t n, void* out) {
  switch (fid) {
    case 0: coop_run<BN254_FR>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 1: coop_run<BN254_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 2: coop_run<PALLAS_FP>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    case 3: coop_run<PALLAS_FQ>(mode, (const affine_t*)pts, n, (fe_t*)out); break;
    default: return 1;
  }
  return 0;
}

// ---- device-side Fiat-Shamir (transcript.cuh): the kernel body of k_sc_round, run sequentially ----
#include "../../nova_b200/csrc/transcript.cuh"

extern "C" int hc_keccak256(const void* data, size_t len, void* out32) {
  if (len > MSG_MAX_BYTES) return 1;
  static msg_buf m;
  msg_reset(m);
  msg_put_bytes(m, (const uint8_t*)data, (uint32_t)len);
  uint64_t d[4];
  keccak256_msg(m, 0, 0, d);
  memcpy(out32, d, 32);
  return 0;
}

template <class F>
static void from_uniform_t(const void* in64, void* out) {
  uint64_t w[8];
  memcpy(w, in64, 64);
  fe_t r = fe_from_uniform<F>(w);
  memcpy(out, &r, 32);
}
extern "C" int hc_from_uniform(int fid, const void* in64, void* out) {
  switch (fid) {
    case 0: from_uniform_t<BN254_FR>(in64, out); break;
    case 1: from_uniform_t<BN254_FQ>(in64, out); break;
    case 2: from_uniform_t<PALLAS_FP>(in64, out); break;
    case 3: from_uniform_t<PALLAS_FQ>(in64, out); break;
    default: return 1;
  }
  return 0;
}

template <class F>
static void sc_round_t(int kind, sc_state* st, const fe_t* res, const fe_t* tau, const fe_t* tau_inv,
                       const uint8_t* pending, uint32_t pending_len, uint8_t la, uint8_t ls, fe_t* out_poly,
                       fe_t* out_r) {
  static msg_buf msg;
  fe_t r3[3] = {res[0], res[1], kind == SC_ROUND_CUBIC3_EQ_M1 ? res[2] : fe_zero<F>()};
  fe_t t = fe_zero<F>(), ti = fe_zero<F>();
  if (kind != SC_ROUND_QUAD_PROD) {
    t = *tau;
    if (kind == SC_ROUND_CUBIC3_EQ) ti = *tau_inv;
  }
  sc_round_poly poly;
  sc_round_build<F>(kind, *st, r3, t, ti, poly);
  fe_t canon[3];
  int ncoef = sc_round_compressed<F>(poly, canon);
  uint32_t flip = sc_round_message(msg, pending, pending_len, la, canon, ncoef, *st, ls);
  for (int k = 0; k < ncoef; k++) out_poly[k] = canon[k];
  uint64_t digest[8], d[4];
  for (int lane = 0; lane < 2; lane++) {
    keccak256_msg(msg, flip, (uint8_t)lane, d);
    for (int i = 0; i < 4; i++) digest[4 * lane + i] = d[i];
  }
  *out_r = sc_round_finish<F>(kind, *st, poly, digest);
}
extern "C" int hc_sc_round(int fid, int kind, void* state144, const void* res, const void* tau, const void* tau_inv,
                           const void* pending, uint32_t pending_len, int absorb_label, int squeeze_label,
                           void* out_poly, void* out_r) {
  static_assert(sizeof(sc_state) == 144, "b200_sc_state layout");
  sc_state st;
  memcpy(&st, state144, 144);
#define RUN(FT) sc_round_t<FT>(kind, &st, (const fe_t*)res, (const fe_t*)tau, (const fe_t*)tau_inv, \
                               (const uint8_t*)pending, pending_len, (uint8_t)absorb_label, (uint8_t)squeeze_label, \
                               (fe_t*)out_poly, (fe_t*)out_r)
  switch (fid) {
    case 0: RUN(BN254_FR); break;
    case 1: RUN(BN254_FQ); break;
    case 2: RUN(PALLAS_FP); break;
    case 3: RUN(PALLAS_FQ); break;
    default: return 1;
  }
#undef RUN
  memcpy(state144, &st, 144);
  return 0;
}

// ---- key validation (curve.cuh affine_valid_raw: canonical coordinates and on the curve; what k_on_curve runs) ----
template <class F>
static void on_curve_t(int b_small, const affine_t* pts, size_t n, unsigned char* ok) {
  fe_t b = fe_from_small_int<F>(b_small);
  for (size_t i = 0; i < n; i++) ok[i] = affine_valid_raw<F>(pts[i], b) ? 1 : 0;
}
extern "C" int hc_on_curve(int fid, int b_small, const void* pts, size_t n, void* ok) {
  switch (fid) {
    case 0: on_curve_t<BN254_FR>(b_small, (const affine_t*)pts, n, (unsigned char*)ok); break;
    case 1: on_curve_t<BN254_FQ>(b_small, (const affine_t*)pts, n, (unsigned char*)ok); break;
    case 2: on_curve_t<PALLAS_FP>(b_small, (const affine_t*)pts, n, (unsigned char*)ok); break;
    case 3: on_curve_t<PALLAS_FQ>(b_small, (const affine_t*)pts, n, (unsigned char*)ok); break;
    default: return 1;
  }
  return 0;
}

// ---- sum of two products with one reduction (field.cuh fe_mul2_add) ----
template <class F>
static void mul2_t(const fe_t* a, const fe_t* b, const fe_t* c, const fe_t* d, fe_t* r, size_t n) {
  for (size_t i = 0; i < n; i++) r[i] = fe_mul2_add<F>(a[i], b[i], c[i], d[i]);
}
extern "C" int hc_mul2_add(int fid, const void* a, const void* b, const void* c, const void* d, void* r, size_t n) {
  switch (fid) {
    case 0: mul2_t<BN254_FR>((const fe_t*)a, (const fe_t*)b, (const fe_t*)c, (const fe_t*)d, (fe_t*)r, n); break;
    case 1: mul2_t<BN254_FQ>((const fe_t*)a, (const fe_t*)b, (const fe_t*)c, (const fe_t*)d, (fe_t*)r, n); break;
    case 2: mul2_t<PALLAS_FP>((const fe_t*)a, (const fe_t*)b, (const fe_t*)c, (const fe_t*)d, (fe_t*)r, n); break;
    case 3: mul2_t<PALLAS_FQ>((const fe_t*)a, (const fe_t*)b, (const fe_t*)c, (const fe_t*)d, (fe_t*)r, n); break;
    default: return 1;
  }
  return 0;
}

// ---- batched sum-check round (transcript_batched.cuh): the body of k_sc_round_batched, sequentially ----
#include "../../nova_b200/csrc/transcript_batched.cuh"

template <class F>
static void sc_round_batched_t(const scb_desc& d, scb_state* state, const fe_t* sums, const uint8_t* pending,
                               uint32_t pending_len, uint8_t la, uint8_t ls, fe_t* out_poly, fe_t* out_r) {
  static msg_buf msg;
  const scb_state st = *state;  // every "lane" reads the state of the round's start
  fe_t evs[SCB_MAX_CLAIMS][3];
  for (int i = 0; i < d.nclaims; i++) {
    fe_t s[3] = {sums[d.slot[i]], sums[d.slot[i] + 1], sums[d.slot[i] + 2]};
    fe_t tau = fe_zero<F>(), tau_inv = fe_zero<F>(), tm1 = fe_zero<F>();
    const bool eqc = d.kind[i] >= SCB_EQ_DEG2;
    const bool has_m1 = eqc && d.slot_m1[i] >= 0;
    if (eqc) {
      tau = *(const fe_t*)d.tau[d.eq_of[i]];
      if (has_m1) tm1 = sums[d.slot_m1[i]];
      else tau_inv = *(const fe_t*)d.tau_inv[d.eq_of[i]];
    }
    scb_claim_evals<F>(d, i, st, s, has_m1 ? &tm1 : nullptr, tau, tau_inv, evs[i]);
  }
  fe_t comb[3];
  for (int k = 0; k < 3; k++) comb[k] = scb_combine<F>(d, st, evs, k);
  sc_round_poly poly;
  scb_poly<F>(st.head.claim, comb[0], comb[1], comb[2], poly);
  fe_t canon[3];
  sc_round_compressed<F>(poly, canon);
  uint32_t flip = sc_round_message(msg, pending, pending_len, la, canon, 3, st.head, ls);
  for (int k = 0; k < 3; k++) out_poly[k] = canon[k];
  uint64_t digest[8], dg[4];
  for (int lane = 0; lane < 2; lane++) {
    keccak256_msg(msg, flip, (uint8_t)lane, dg);
    for (int i = 0; i < 4; i++) digest[4 * lane + i] = dg[i];
  }
  sc_state head = st.head;
  fe_t r = sc_round_finish<F>(SC_ROUND_QUAD_PROD, head, poly, digest);
  for (int i = 0; i < d.nclaims; i++)
    if (d.kind[i] >= SCB_EQ_DEG2) state->claim[i] = scb_update_claim<F>(st.claim[i], evs[i], r);
  for (int g = 0; g < d.neq; g++) state->q[g] = scb_bound<F>(st.q[g], *(const fe_t*)d.tau[g], r);
  state->head = head;
  *out_r = r;
}
extern "C" int hc_sc_round_batched(int fid, const void* desc, void* state, const void* sums, const void* pending,
                                   uint32_t pending_len, int absorb_label, int squeeze_label, void* out_poly,
                                   void* out_r) {
  static_assert(sizeof(scb_state) == 1296, "b200_scb_state layout");
  static_assert(sizeof(scb_desc) == 8 + 4 * 4 * SCB_MAX_CLAIMS + 2 * 8 * SCB_MAX_EQ, "b200_scb_desc layout");
  const scb_desc& d = *(const scb_desc*)desc;
  if (d.nclaims < 1 || d.nclaims > SCB_MAX_CLAIMS || d.neq < 0 || d.neq > SCB_MAX_EQ) return 1;
#define RUNB(FT) sc_round_batched_t<FT>(d, (scb_state*)state, (const fe_t*)sums, (const uint8_t*)pending, pending_len, \
                                        (uint8_t)absorb_label, (uint8_t)squeeze_label, (fe_t*)out_poly, (fe_t*)out_r)
  switch (fid) {
    case 0: RUNB(BN254_FR); break;
    case 1: RUNB(BN254_FQ); break;
    case 2: RUNB(PALLAS_FP); break;
    case 3: RUNB(PALLAS_FQ); break;
    default: return 1;
  }
#undef RUNB
  return 0;
}
