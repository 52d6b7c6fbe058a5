// A CPU stand-in for libnova_b200.so, exporting the C-ABI symbols the C++ host-mirror test uses
// (include/nova_b200.h), answered by the C oracle (oracle/liboracle.so) on host memory.
// TEST INFRASTRUCTURE ONLY: it lets tests/cpp/host_mirror_test -- the compiled-language host layer of
// include/nova_b200.hpp, including the device-resident folding step and commits issued from 8 threads --
// run on the CPU box, so that its glue (argument order, sizes, ownership, error mapping) is checked
// without a GPU.  It says nothing about the CUDA kernels; the `-m gpu` tests do that.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nova_b200.h"

extern "C" {
int orc_msm(int curve, const void* scalars, const void* bases, size_t n, int nthreads, void* out_affine);
int orc_msm_small(int curve, const uint64_t* scalars, const void* bases, size_t n, int max_bits, int nthreads, void* out_affine);
int orc_batch_add(int curve, const void* bases, const uint64_t* idx, size_t m, int nthreads, void* out_affine);
int orc_cross_term(int fid, const void* az, const void* bz, const void* cz, const void* e1, const void* e2, const void* u,
                   size_t n, void* t);
int orc_axpy(int fid, const void* a, const void* b, const void* r, size_t n, void* out);
int orc_vec_add(int fid, const void* a, const void* b, size_t n, void* out);
int orc_bind_top(int fid, void* z, size_t n, const void* r);
int orc_field_from_u64(int fid, const uint64_t* v, size_t n, void* out);
int orc_fe_op(int fid, int op, const void* a, const void* b, void* out, size_t n);
int orc_on_curve(int curve, const void* pt, const void* b_mont);
int orc_gen_bases(int curve, const void* gen, const uint64_t* k0, size_t n, void* out);
int orc_sc_eval(int fid, int form, const void* a, const void* b, const void* c, size_t len, const void* eql, const void* eqr,
                int shift, void* out);
int orc_spmv(int fid, const void* data, const uint64_t* indices, const uint64_t* indptr, size_t rows, const void* z, void* out);
int orc_eq_table(int fid, const void* r, int ell, void* out);
// the HOST BUILD of the device round kernel (tests/hostcheck/hostcheck.cpp, from nova_b200/csrc/transcript.cuh)
int hc_sc_round(int fid, int kind, void* state144, const void* res, const void* tau, const void* tau_inv, const void* pending,
                uint32_t pending_len, int absorb_label, int squeeze_label, void* out_poly, void* out_r);
}

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& m) { g_err = m; return code; }
const int BASE_FIELD[4] = {B200_FIELD_BN254_FQ, B200_FIELD_BN254_FR, B200_FIELD_PALLAS_FP, B200_FIELD_PALLAS_FQ};

struct ck_rec { int curve; std::vector<unsigned char> bases, h; size_t n; };
struct mat_rec { int fid; std::vector<unsigned char> data; std::vector<uint64_t> idx, ptr; size_t rows, cols; };
struct ws_rec { std::shared_ptr<ck_rec> key; size_t n; std::vector<unsigned char> w; size_t filled = 0; bool done = false; };
std::mutex g_mu;
std::map<uint64_t, std::shared_ptr<ck_rec>> g_keys;
std::map<uint64_t, std::shared_ptr<mat_rec>> g_mats;
std::map<uint64_t, std::shared_ptr<ws_rec>> g_ws;
uint64_t g_next = 1;

template <class M>
typename M::mapped_type lookup(M& m, uint64_t h) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = m.find(h);
  return it == m.end() ? nullptr : it->second;
}

// affine (64 B) -> Jacobian (x, y, 1) or the identity (z = 0)
void to_jacobian(int curve, const unsigned char* aff, void* out) {
  unsigned char* o = (unsigned char*)out;
  memcpy(o, aff, 64);
  bool ident = true;
  for (int i = 0; i < 64; i++) ident &= aff[i] == 0;
  uint64_t one = 1;
  if (ident) memset(o + 64, 0, 32);
  else orc_field_from_u64(BASE_FIELD[curve], &one, 1, o + 64);
}

int msm_key(const ck_rec& k, const void* scalars, size_t off, size_t n, const void* blind, void* out) {
  std::vector<unsigned char> sc((n + 1) * 32), bs((n + 1) * 64);
  memcpy(sc.data(), scalars, n * 32);
  memcpy(bs.data(), k.bases.data() + 64 * off, n * 64);
  size_t m = n;
  if (blind) {
    memcpy(sc.data() + 32 * n, blind, 32);
    memcpy(bs.data() + 64 * n, k.h.data(), 64);
    m++;
  }
  unsigned char aff[64];
  if (orc_msm(k.curve, sc.data(), bs.data(), m, 1, aff)) return fail(B200_E_ARG, "orc_msm failed");
  to_jacobian(k.curve, aff, out);
  return B200_OK;
}
}  // namespace

extern "C" {
const char* b200_last_error(void) { return g_err.c_str(); }
const char* b200_version(void) { return "nova_b200 EMULATED on the CPU oracle (tests only)"; }
int b200_init(int) { return B200_OK; }
int b200_sync(void) { return B200_OK; }
int b200_dev_alloc(size_t bytes, void** p) { *p = malloc(bytes ? bytes : 1); return *p ? B200_OK : B200_E_NOMEM; }
int b200_dev_free(void* p) { free(p); return B200_OK; }
int b200_memcpy_h2d(void* d, const void* h, size_t n) { memcpy(d, h, n); return B200_OK; }
int b200_memcpy_d2h(void* h, const void* d, size_t n) { memcpy(h, d, n); return B200_OK; }
int b200_memcpy_d2d(void* d, const void* s, size_t n, void*) { memcpy(d, s, n); return B200_OK; }
int b200_memset_dev(void* d, int byte, size_t n, void*) { memset(d, byte, n); return B200_OK; }

int b200_ck_register(int curve, const void* bases, size_t n, const void* h, int, uint64_t* handle) {
  if (curve < 0 || curve > 3 || !bases || !n || !handle) return fail(B200_E_ARG, "bad key");
  auto k = std::make_shared<ck_rec>();
  k->curve = curve;
  k->n = n;
  k->bases.assign((const unsigned char*)bases, (const unsigned char*)bases + 64 * n);
  if (h) k->h.assign((const unsigned char*)h, (const unsigned char*)h + 64);
  std::lock_guard<std::mutex> lk(g_mu);
  *handle = g_next++;
  g_keys[*handle] = k;
  return B200_OK;
}
int b200_ck_release(uint64_t h) { std::lock_guard<std::mutex> lk(g_mu); return g_keys.erase(h) ? B200_OK : B200_E_HANDLE; }
// one process, N GPUs: on the CPU the "devices" are one oracle
int b200_device_count(int* n) { *n = 1; return B200_OK; }
int b200_mgpu_init(int ndev, const int*) { return ndev >= 1 && ndev <= 8 ? B200_OK : B200_E_ARG; }
int b200_mgpu_ck_register(int curve, const void* bases, size_t n, const void* h, int wb, uint64_t* key) {
  return b200_ck_register(curve, bases, n, h, wb, key);
}
int b200_mgpu_ck_release(uint64_t key) { return b200_ck_release(key); }
int b200_commit(uint64_t h, const void* scalars, size_t n, const void* r, void* out);
int b200_mgpu_commit(uint64_t key, const void* scalars, size_t n, const void* r, void* out) {
  return b200_commit(key, scalars, n, r, out);
}
int b200_host_alloc(size_t bytes, void** p) { *p = malloc(bytes ? bytes : 1); return *p ? B200_OK : B200_E_NOMEM; }
int b200_host_free(void* p) { free(p); return B200_OK; }
int b200_ck_setup_synthetic(int curve, const void* gen, uint64_t k0, size_t n, int with_h, int, uint64_t* handle) {
  if (curve < 0 || curve > 3 || !gen || !n || !handle) return fail(B200_E_ARG, "bad key");
  std::vector<unsigned char> pts(64 * (n + 1));
  uint64_t k[4] = {k0, 0, 0, 0};
  if (orc_gen_bases(curve, gen, k, n + 1, pts.data())) return fail(B200_E_ARG, "orc_gen_bases failed");  // P_i = (k0 + i) G
  return b200_ck_register(curve, pts.data(), n, with_h ? pts.data() + 64 * n : nullptr, 0, handle);
}

int b200_msm(uint64_t h, size_t off, const void* scalars, size_t n, void* out) {
  auto k = lookup(g_keys, h);
  if (!k) return fail(B200_E_HANDLE, "unknown key");
  if (off + n > k->n) return fail(B200_E_RANGE, "msm slice exceeds key length");
  return msm_key(*k, scalars, off, n, nullptr, out);
}
int b200_commit(uint64_t h, const void* scalars, size_t n, const void* r, void* out) {
  auto k = lookup(g_keys, h);
  if (!k) return fail(B200_E_HANDLE, "unknown key");
  if (n > k->n) return fail(B200_E_RANGE, "commit exceeds key length");
  if (r && k->h.empty()) return fail(B200_E_ARG, "key was registered without a blinding generator");
  return msm_key(*k, scalars, 0, n, r, out);
}
int b200_commit_dev(uint64_t h, const void* scalars, size_t n, const void* blind, void* out, void*) {
  return b200_commit(h, scalars, n, blind, out);
}
int b200_msm_adhoc(int curve, const void* bases, const void* scalars, size_t n, void* out) {
  unsigned char aff[64];
  if (orc_msm(curve, scalars, bases, n, 1, aff)) return fail(B200_E_ARG, "orc_msm failed");
  to_jacobian(curve, aff, out);
  return B200_OK;
}
int b200_msm_batch(uint64_t h, const void* const* scalars, const size_t* lens, size_t k, void* out) {
  for (size_t j = 0; j < k; j++) {
    int rc = b200_msm(h, 0, scalars[j], lens[j], (char*)out + 96 * j);
    if (rc) return rc;
  }
  return B200_OK;
}
int b200_msm_small(uint64_t h, size_t off, const void* scalars, int elem_bytes, size_t n, int max_bits, void* out) {
  auto k = lookup(g_keys, h);
  if (!k) return fail(B200_E_HANDLE, "unknown key");
  if (off + n > k->n) return fail(B200_E_RANGE, "msm slice exceeds key length");
  std::vector<uint64_t> v(n);
  for (size_t i = 0; i < n; i++) {
    uint64_t x = 0;
    memcpy(&x, (const char*)scalars + (size_t)elem_bytes * i, elem_bytes);
    v[i] = x;
  }
  unsigned char aff[64];
  if (orc_msm_small(k->curve, v.data(), k->bases.data() + 64 * off, n, max_bits > 0 ? max_bits : -1, 1, aff))
    return fail(B200_E_ARG, "orc_msm_small failed");
  to_jacobian(k->curve, aff, out);
  return B200_OK;
}
int b200_msm_indices(uint64_t h, const uint64_t* idx, size_t m, void* out) {
  auto k = lookup(g_keys, h);
  if (!k) return fail(B200_E_HANDLE, "unknown key");
  unsigned char aff[64];
  if (orc_batch_add(k->curve, k->bases.data(), idx, m, 1, aff)) return fail(B200_E_ARG, "orc_batch_add failed");
  to_jacobian(k->curve, aff, out);
  return B200_OK;
}
// smallest index of a point with a non-canonical coordinate or off the curve (k_on_curve's contract)
static size_t first_invalid(int curve, const unsigned char* pts, size_t n) {
  static const int64_t B[4] = {3, -17, 5, 5};
  uint64_t mag = (uint64_t)(B[curve] < 0 ? -B[curve] : B[curve]);
  unsigned char b[32], canon[64];
  orc_field_from_u64(BASE_FIELD[curve], &mag, 1, b);
  if (B[curve] < 0) orc_fe_op(BASE_FIELD[curve], 6, b, b, b, 1);
  for (size_t i = 0; i < n; i++) {
    // canonical <=> converting out of and back into Montgomery form reproduces the bytes
    unsigned char back[64];
    orc_fe_op(BASE_FIELD[curve], 5, pts + 64 * i, pts + 64 * i, canon, 2);
    orc_fe_op(BASE_FIELD[curve], 4, canon, canon, back, 2);
    if (memcmp(back, pts + 64 * i, 64) != 0) return i;
    if (orc_on_curve(curve, pts + 64 * i, b) != 1) return i;
  }
  return SIZE_MAX;
}
int b200_ck_validate(int curve, const void* bases, size_t n, size_t* first_bad) {
  *first_bad = first_invalid(curve, (const unsigned char*)bases, n);
  return B200_OK;
}
int b200_ck_register_checked(int curve, const void* bases, size_t n, const void* h, int wb, uint64_t* handle,
                             size_t* first_bad) {
  if (curve < 0 || curve > 3 || !bases || !n || !handle || !first_bad) return fail(B200_E_ARG, "bad key");
  *handle = 0;
  *first_bad = first_invalid(curve, (const unsigned char*)bases, n);
  if (*first_bad == SIZE_MAX && h && first_invalid(curve, (const unsigned char*)h, 1) == 0) *first_bad = n;
  if (*first_bad != SIZE_MAX) return fail(B200_E_POINT, "key point is non-canonical or not on the curve");
  return b200_ck_register(curve, bases, n, h, wb, handle);
}

// field vectors (host and "device" pointers are the same thing here)
int b200_axpy(int f, const void* a, const void* b, const void* r, size_t n, void* out) { return orc_axpy(f, a, b, r, n, out) ? B200_E_ARG : B200_OK; }
int b200_axpy_dev(int f, const void* a, const void* b, const void* r, size_t n, void* out, void*) { return b200_axpy(f, a, b, r, n, out); }
int b200_vec_add_dev(int f, const void* a, const void* b, size_t n, void* out, void*) { return orc_vec_add(f, a, b, n, out) ? B200_E_ARG : B200_OK; }
int b200_cross_term(int f, const void* az, const void* bz, const void* cz, const void* e1, const void* e2, const void* u, size_t n, void* t) {
  return orc_cross_term(f, az, bz, cz, e1, e2, u, n, t) ? B200_E_ARG : B200_OK;
}
int b200_cross_term_dev(int f, const void* az, const void* bz, const void* cz, const void* e1, const void* e2, const void* u, size_t n, void* t, void*) {
  return b200_cross_term(f, az, bz, cz, e1, e2, u, n, t);
}
int b200_bind_top(int f, void* z, size_t n, const void* r) {
  if (n & 1) return fail(B200_E_ARG, "bind_top needs an even length");
  return orc_bind_top(f, z, n, r) ? B200_E_ARG : B200_OK;
}
int b200_sc_eval(int f, int form, const void* A, const void* B, const void* C, size_t len, const void* el, size_t, const void* er, size_t,
                 int shift, void* out) {
  return orc_sc_eval(f, form, A, B, C, len, el, er, shift, out) ? B200_E_ARG : B200_OK;
}

// sparse matrices
int b200_spmv_register(int f, const void* data, const uint64_t* indices, const uint64_t* indptr, size_t rows, size_t cols, uint64_t* handle) {
  auto m = std::make_shared<mat_rec>();
  m->fid = f;
  m->rows = rows;
  m->cols = cols;
  m->ptr.assign(indptr, indptr + rows + 1);
  size_t nnz = m->ptr[rows];
  m->idx.assign(indices, indices + nnz);
  m->data.assign((const unsigned char*)data, (const unsigned char*)data + 32 * nnz);
  std::lock_guard<std::mutex> lk(g_mu);
  *handle = g_next++;
  g_mats[*handle] = m;
  return B200_OK;
}
int b200_spmv_release(uint64_t h) { std::lock_guard<std::mutex> lk(g_mu); return g_mats.erase(h) ? B200_OK : B200_E_HANDLE; }
int b200_spmv_dev(uint64_t h, const void* z1, const void* z2, void* o1, void* o2, void*) {
  auto m = lookup(g_mats, h);
  if (!m) return fail(B200_E_HANDLE, "unknown matrix");
  orc_spmv(m->fid, m->data.data(), m->idx.data(), m->ptr.data(), m->rows, z1, o1);
  if (z2) orc_spmv(m->fid, m->data.data(), m->idx.data(), m->ptr.data(), m->rows, z2, o2);
  return B200_OK;
}
int b200_spmv_multi(const uint64_t* hs, size_t k, const void* z1, const void* z2, size_t z_len, void* const* o1, void* const* o2) {
  for (size_t j = 0; j < k; j++) {
    auto m = lookup(g_mats, hs[j]);
    if (!m) return fail(B200_E_HANDLE, "unknown matrix");
    if (z_len != m->cols) return fail(B200_E_ARG, "InvalidWitnessLength");
    int rc = b200_spmv_dev(hs[j], z1, z2, o1[j], z2 ? o2[j] : nullptr, nullptr);
    if (rc) return rc;
  }
  return B200_OK;
}

// sum-check loops with the transcript "on the device": reductions and binds from the oracle, the round step from
// the host build of k_sc_round's body, strung together as csrc/capi_sumcheck.inc does
static int sumcheck_loop(int f, const void* claim, int l, void* const* polys, int npoly, int ncoef, b200_transcript* tr,
                         const void* pending, size_t plen, void* polys_out, void* r_out, void* finals_out,
                         const unsigned char* taus) {
  unsigned char state[144];
  uint64_t one = 1;
  memcpy(state, claim, 32);
  orc_field_from_u64(f, &one, 1, state + 32);
  memcpy(state + 64, &tr->round, 8);
  memcpy(state + 72, tr->state, 64);
  memset(state + 136, 0, 8);
  const int fh = l / 2, sh = l - fh;
  std::vector<std::vector<unsigned char>> left, right;
  std::vector<unsigned char> tinv;
  if (taus) {
    for (int k = 0; k < (fh > 0 ? fh : 1); k++) { left.emplace_back((size_t)32 << k); orc_eq_table(f, taus + 32 * (fh - k), k, left[k].data()); }
    for (int k = 0; k <= sh; k++) { right.emplace_back((size_t)32 << k); orc_eq_table(f, taus + 32 * (l - k), k, right[k].data()); }
    tinv.resize(32 * l);
    orc_fe_op(f, 3, taus, taus, tinv.data(), l);  // inverses (0 -> 0)
  }
  size_t len = (size_t)1 << l;
  for (int j = 0; j < l; j++, len >>= 1) {
    unsigned char res[96] = {0};
    int kind = 0;
    if (!taus) {
      orc_sc_eval(f, 0, polys[0], polys[1], nullptr, len, nullptr, nullptr, 0, res);
    } else {
      const int rnd = j + 1;
      const void *L = nullptr, *R;
      int shift = 0;
      if (rnd < fh) { L = left[fh - rnd].data(); R = right[sh].data(); shift = sh; } else R = right[l - rnd].data();
      bool zero = true;
      for (int b = 0; b < 32; b++) zero &= taus[32 * j + b] == 0;
      orc_sc_eval(f, 4, polys[0], polys[1], polys[2], len, L, R, shift, res);
      if (zero) orc_sc_eval(f, 7, polys[0], polys[1], polys[2], len, L, R, shift, res + 64);
      kind = zero ? 2 : 1;
    }
    if (hc_sc_round(f, kind, state, res, taus ? taus + 32 * j : nullptr, taus ? tinv.data() + 32 * j : nullptr,
                    j == 0 ? pending : nullptr, j == 0 ? (uint32_t)plen : 0u, 'p', 'c', (char*)polys_out + 32 * ncoef * j,
                    (char*)r_out + 32 * j))
      return fail(B200_E_ARG, "hc_sc_round failed");
    for (int k = 0; k < npoly; k++) orc_bind_top(f, polys[k], len, (char*)r_out + 32 * j);
  }
  for (int k = 0; k < npoly; k++) memcpy((char*)finals_out + 32 * k, polys[k], 32);
  memcpy(&tr->round, state + 64, 8);
  memcpy(tr->state, state + 72, 64);
  return B200_OK;
}
int b200_sumcheck_quad_prod(int f, const void* claim, int num_rounds, void* A, void* B, b200_transcript* tr, const void* pending,
                            size_t plen, void* polys_out, void* r_out, void* finals_out) {
  if (plen > B200_SC_MAX_PENDING) return fail(B200_E_ARG, "pending transcript bytes exceed the limit");
  void* polys[2] = {A, B};
  return sumcheck_loop(f, claim, num_rounds, polys, 2, 2, tr, pending, plen, polys_out, r_out, finals_out, nullptr);
}
int b200_sumcheck_cubic3(int f, const void* claim, const void* taus, int num_rounds, void* A, void* B, void* C, b200_transcript* tr,
                         const void* pending, size_t plen, void* polys_out, void* r_out, void* finals_out) {
  if (plen > B200_SC_MAX_PENDING) return fail(B200_E_ARG, "pending transcript bytes exceed the limit");
  void* polys[3] = {A, B, C};
  return sumcheck_loop(f, claim, num_rounds, polys, 3, 3, tr, pending, plen, polys_out, r_out, finals_out,
                       (const unsigned char*)taus);
}

// streamed witness hand-off
int b200_witness_begin(uint64_t ck, size_t n, uint64_t* h) {
  auto k = lookup(g_keys, ck);
  if (!k) return fail(B200_E_HANDLE, "unknown key");
  if (n > k->n) return fail(B200_E_RANGE, "witness exceeds key length");
  auto w = std::make_shared<ws_rec>();
  w->key = k;
  w->n = n;
  w->w.assign(32 * n, 0);
  std::lock_guard<std::mutex> lk(g_mu);
  *h = g_next++;
  g_ws[*h] = w;
  return B200_OK;
}
int b200_witness_append(uint64_t h, const void* scalars, size_t count) {
  auto w = lookup(g_ws, h);
  if (!w) return fail(B200_E_HANDLE, "unknown witness stream");
  if (w->done) return fail(B200_E_ARG, "witness stream already finished");
  if (w->filled + count > w->n) return fail(B200_E_RANGE, "append overflows the witness");
  memcpy(w->w.data() + 32 * w->filled, scalars, 32 * count);
  w->filled += count;
  return B200_OK;
}
int b200_witness_finish(uint64_t h, const void* r, void* out, void** d_w) {
  auto w = lookup(g_ws, h);
  if (!w) return fail(B200_E_HANDLE, "unknown witness stream");
  if (w->done) return fail(B200_E_ARG, "witness stream already finished");
  if (r && w->key->h.empty()) return fail(B200_E_ARG, "key was registered without a blinding generator");
  w->done = true;
  if (d_w) *d_w = w->w.data();
  return msm_key(*w->key, w->w.data(), 0, w->n, r, out);
}
int b200_witness_reset(uint64_t h) {
  auto w = lookup(g_ws, h);
  if (!w) return fail(B200_E_HANDLE, "unknown witness stream");
  w->w.assign(32 * w->n, 0);
  w->filled = 0;
  w->done = false;
  return B200_OK;
}
int b200_witness_release(uint64_t h) { std::lock_guard<std::mutex> lk(g_mu); return g_ws.erase(h) ? B200_OK : B200_E_HANDLE; }
}  // extern "C"
