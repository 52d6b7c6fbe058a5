/* CPU oracle for the Nova prover hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  nova_b200/ never links or calls it; the product path fails loudly when
 * the CUDA library is missing.
 *
 * What it restates (plain C, unsigned __int128 Montgomery arithmetic, pthreads):
 *   - prime fields of halo2curves 0.9.0 (Cargo.toml:38; NOT vendored in /root/reference):
 *     4 x u64 little-endian limbs, Montgomery R = 2^256, moduli from
 *     src/provider/bn256_grumpkin.rs:39-40,84-85 and src/provider/pasta.rs:37-38,45-46
 *   - XYZZ bucket arithmetic                         src/provider/msm.rs:38-183
 *   - msm() classification + 11-group dispatch       src/provider/msm.rs:225-419
 *   - msm_simple / accumulate_bases                  src/provider/msm.rs:422-454
 *   - msm_small_with_max_num_bits, msm_binary, msm_10, msm_small_rest, compute_ln
 *                                                    src/provider/msm.rs:469-686
 *   - batch_add                                      src/provider/msm.rs:689-708
 *   - halo2curves::msm::msm_best (called at msm.rs:411,500; source absent): restated from its
 *     published algorithm -- windowed Pippenger over canonical scalar bytes with signed (Booth)
 *     digits, window from ln(n), per-thread slices reduced at the end.
 *   - R1CS field arithmetic: cross-term, folds, bind  src/r1cs/mod.rs:614-620,650-657,1044-1073;
 *                                                    src/spartan/polys/multilinear.rs:65-84
 *
 * Parity status: field encodings PINNED by the keccak transcript golden vectors
 * (src/provider/keccak.rs:241-258) through oracle/pyref.py, against which this file is checked
 * element-by-element (tests/test_oracle_golden.py).  MSM outputs: the reference holds no literal
 * commitment bytes; they are pinned semantically (every algorithm here == naive
 * sum == Python big-int group law) exactly as the reference's own tests do (msm.rs:722-821,
 * curve_property_tests.rs:172-218).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
static double now_s(void){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+t.tv_nsec*1e-9;}

#include "field_constants.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct { fe x, y; } aff;            /* identity: x = y = 0 */
typedef struct { fe x, y, zz, zzz; } xyzz;  /* identity: zz = 0 (msm.rs:53-62) */

/* ------------------------------------------------------------------ field ---------------- */
static inline int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe* a, const fe* b) {
  return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline int geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (uint64_t)t;
    br = (t >> 64) & 1;
  }
  return (uint64_t)br;
}
static inline uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (uint64_t)t;
    c = t >> 64;
  }
  return (uint64_t)c;
}
static inline void fe_add(const orc_field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  add4(t, a->l, b->l); /* p < 2^255: no carry out */
  if (geq(t, F->p)) sub4(r->l, t, F->p); else memcpy(r->l, t, 32);
}
static inline void fe_sub(const orc_field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  if (sub4(t, a->l, b->l)) add4(r->l, t, F->p); else memcpy(r->l, t, 32);
}
static inline void fe_neg(const orc_field_t* F, fe* r, const fe* a) {
  if (fe_is_zero(a)) { memset(r, 0, 32); return; }
  sub4(r->l, F->p, a->l);
}
static inline void fe_dbl(const orc_field_t* F, fe* r, const fe* a) { fe_add(F, r, a, a); }
/* CIOS Montgomery product, portable form (unsigned __int128) */
static inline void fe_mul_portable(const orc_field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      u128 s = (u128)a->l[j] * b->l[i] + t[j] + c;
      t[j] = (uint64_t)s;
      c = s >> 64;
    }
    u128 s = (u128)t[4] + c;
    t[4] = (uint64_t)s;
    t[5] = (uint64_t)(s >> 64);
    uint64_t m = t[0] * F->inv;
    c = ((u128)m * F->p[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) {
      u128 s2 = (u128)m * F->p[j] + t[j] + c;
      t[j - 1] = (uint64_t)s2;
      c = s2 >> 64;
    }
    s = (u128)t[4] + c;
    t[3] = (uint64_t)s;
    t[4] = t[5] + (uint64_t)(s >> 64);
  }
  if (t[4] || geq(t, F->p)) sub4(r->l, t, F->p); else memcpy(r->l, t, 32);
}
/* The same product on the mulx / adcx-style path (BMI2 + ADX): what halo2curves' `asm` feature (README.md:54, "up to
 * 50 %") gives the reference on x86-64 -- so that the CPU baseline is not flattered by a slow multiplier.  Per row:
 * four mulx, the low halves on the adcx carry chain and the high halves on the adox chain (inline assembly: gcc's
 * _addcarry_u64 intrinsics serialise the two chains and come out no faster than the portable form).  Selected at run time
 * (cpu_has_adx; ORACLE_NO_ADX=1 forces the portable form); bit-identical to it (tests/test_oracle_golden.py). */
#if defined(__x86_64__)
/* CIOS rows fully unrolled; accumulators rotate through r11-r15, rbx (generated, see the comment above):
 *   row i:    (t0..t5) += a * b[i]           mulx + adcx (low halves) / adox (high halves), two carry chains
 *             m = t0 * inv ; (t0..t5) += m * p   -> t0 == 0, the register is reused as the next row's t5 */
__attribute__((target("bmi2,adx"))) static void fe_mul_adx(const orc_field_t* F, fe* r, const fe* a, const fe* b) {
  uint64_t t[5];
  const uint64_t inv = F->inv;
  __asm__ volatile(
      "xorl %%r11d, %%r11d\n\t"
      "xorl %%r12d, %%r12d\n\t"
      "xorl %%r13d, %%r13d\n\t"
      "xorl %%r14d, %%r14d\n\t"
      "xorl %%r15d, %%r15d\n\t"
      "xorl %%ebx, %%ebx\n\t"
      "movq 0(%[b]), %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r11\n\t"
      "adoxq %%r9, %%r12\n\t"
      "mulxq 8(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r12\n\t"
      "adoxq %%r9, %%r13\n\t"
      "mulxq 16(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 24(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "adcxq %%r10, %%r15\n\t"
      "adoxq %%r10, %%rbx\n\t"
      "adcxq %%r10, %%rbx\n\t"
      "movq %%r11, %%rdx\n\t"
      "imulq %[inv], %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r11\n\t"
      "adoxq %%r9, %%r12\n\t"
      "mulxq 8(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r12\n\t"
      "adoxq %%r9, %%r13\n\t"
      "mulxq 16(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 24(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "adcxq %%r10, %%r15\n\t"
      "adoxq %%r10, %%rbx\n\t"
      "adcxq %%r10, %%rbx\n\t"
      "movq 8(%[b]), %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r12\n\t"
      "adoxq %%r9, %%r13\n\t"
      "mulxq 8(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 16(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 24(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "adcxq %%r10, %%rbx\n\t"
      "adoxq %%r10, %%r11\n\t"
      "adcxq %%r10, %%r11\n\t"
      "movq %%r12, %%rdx\n\t"
      "imulq %[inv], %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r12\n\t"
      "adoxq %%r9, %%r13\n\t"
      "mulxq 8(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 16(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 24(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "adcxq %%r10, %%rbx\n\t"
      "adoxq %%r10, %%r11\n\t"
      "adcxq %%r10, %%r11\n\t"
      "movq 16(%[b]), %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 8(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 16(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "mulxq 24(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%rbx\n\t"
      "adoxq %%r9, %%r11\n\t"
      "adcxq %%r10, %%r11\n\t"
      "adoxq %%r10, %%r12\n\t"
      "adcxq %%r10, %%r12\n\t"
      "movq %%r13, %%rdx\n\t"
      "imulq %[inv], %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r13\n\t"
      "adoxq %%r9, %%r14\n\t"
      "mulxq 8(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 16(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "mulxq 24(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%rbx\n\t"
      "adoxq %%r9, %%r11\n\t"
      "adcxq %%r10, %%r11\n\t"
      "adoxq %%r10, %%r12\n\t"
      "adcxq %%r10, %%r12\n\t"
      "movq 24(%[b]), %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 8(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "mulxq 16(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%rbx\n\t"
      "adoxq %%r9, %%r11\n\t"
      "mulxq 24(%[a]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r11\n\t"
      "adoxq %%r9, %%r12\n\t"
      "adcxq %%r10, %%r12\n\t"
      "adoxq %%r10, %%r13\n\t"
      "adcxq %%r10, %%r13\n\t"
      "movq %%r14, %%rdx\n\t"
      "imulq %[inv], %%rdx\n\t"
      "xorl %%r10d, %%r10d\n\t"
      "mulxq 0(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r14\n\t"
      "adoxq %%r9, %%r15\n\t"
      "mulxq 8(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r15\n\t"
      "adoxq %%r9, %%rbx\n\t"
      "mulxq 16(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%rbx\n\t"
      "adoxq %%r9, %%r11\n\t"
      "mulxq 24(%[p]), %%r8, %%r9\n\t"
      "adcxq %%r8, %%r11\n\t"
      "adoxq %%r9, %%r12\n\t"
      "adcxq %%r10, %%r12\n\t"
      "adoxq %%r10, %%r13\n\t"
      "adcxq %%r10, %%r13\n\t"
      "movq %%r15, 0(%[out])\n\t"
      "movq %%rbx, 8(%[out])\n\t"
      "movq %%r11, 16(%[out])\n\t"
      "movq %%r12, 24(%[out])\n\t"
      "movq %%r13, 32(%[out])\n\t"
      :
      : [a] "r"(a->l), [b] "r"(b->l), [p] "r"(F->p), [inv] "m"(inv), [out] "r"(t)
      : "rdx", "r8", "r9", "r10", "r11", "r12", "r13", "r14", "r15", "rbx", "cc", "memory");
  if (t[4] || geq(t, F->p)) sub4(r->l, t, F->p); else memcpy(r->l, t, 32);
}
static int g_have_adx = -1;
static inline int cpu_has_adx(void) {
  if (g_have_adx < 0) {
    const char* e = getenv("ORACLE_NO_ADX");
    g_have_adx = (e == NULL || e[0] != '1') && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx");
  }
  return g_have_adx;
}
static inline void fe_mul(const orc_field_t* F, fe* r, const fe* a, const fe* b) {
  if (cpu_has_adx()) fe_mul_adx(F, r, a, b);
  else fe_mul_portable(F, r, a, b);
}
#else
static inline void fe_mul(const orc_field_t* F, fe* r, const fe* a, const fe* b) { fe_mul_portable(F, r, a, b); }
#endif
static inline void fe_sqr(const orc_field_t* F, fe* r, const fe* a) { fe_mul(F, r, a, a); }
static void fe_one(const orc_field_t* F, fe* r) { memcpy(r->l, F->r, 32); }
static void fe_from_mont(const orc_field_t* F, fe* r, const fe* a) {
  fe one = {{1, 0, 0, 0}};
  fe_mul(F, r, a, &one);
}
static void fe_to_mont(const orc_field_t* F, fe* r, const fe* a) {
  fe r2;
  memcpy(r2.l, F->r2, 32);
  fe_mul(F, r, a, &r2);
}
static void fe_inv(const orc_field_t* F, fe* r, const fe* a) { /* a^(p-2); inv(0) = 0 */
  uint64_t e[4], two[4] = {2, 0, 0, 0};
  sub4(e, F->p, two);
  fe acc;
  fe_one(F, &acc);
  for (int i = 255; i >= 0; i--) {
    fe_sqr(F, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(F, &acc, &acc, a);
  }
  *r = acc;
}
static void fe_from_u64(const orc_field_t* F, fe* r, uint64_t v) {
  fe t = {{v, 0, 0, 0}};
  fe_to_mont(F, r, &t);
}

/* ------------------------------------------------------------------ XYZZ (msm.rs:38-183) -- */
static void xyzz_zero(const orc_field_t* F, xyzz* b) { /* msm.rs:53-60 */
  fe_one(F, &b->x);
  fe_one(F, &b->y);
  memset(&b->zz, 0, 32);
  memset(&b->zzz, 0, 32);
}
static int xyzz_is_zero(const xyzz* b) { return fe_is_zero(&b->zz); }
static int aff_is_identity(const aff* p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }

static void xyzz_double(const orc_field_t* F, xyzz* s) { /* msm.rs:65-88 */
  if (xyzz_is_zero(s)) return;
  fe u, v, w, S, xsq, m, t, x3, y3;
  fe_dbl(F, &u, &s->y);
  fe_sqr(F, &v, &u);
  fe_mul(F, &w, &u, &v);
  fe_mul(F, &S, &s->x, &v);
  fe_sqr(F, &xsq, &s->x);
  fe_dbl(F, &m, &xsq);
  fe_add(F, &m, &m, &xsq);
  fe_sqr(F, &x3, &m);
  fe_dbl(F, &t, &S);
  fe_sub(F, &x3, &x3, &t);
  fe_sub(F, &t, &S, &x3);
  fe_mul(F, &y3, &m, &t);
  fe_mul(F, &t, &w, &s->y);
  fe_sub(F, &y3, &y3, &t);
  s->x = x3;
  s->y = y3;
  fe_mul(F, &s->zz, &s->zz, &v);
  fe_mul(F, &s->zzz, &s->zzz, &w);
}
static void xyzz_add(const orc_field_t* F, xyzz* s, const xyzz* o) { /* msm.rs:91-123 */
  if (xyzz_is_zero(o)) return;
  if (xyzz_is_zero(s)) { *s = *o; return; }
  fe u1, u2, s1, s2;
  fe_mul(F, &u1, &s->x, &o->zz);
  fe_mul(F, &u2, &o->x, &s->zz);
  fe_mul(F, &s1, &s->y, &o->zzz);
  fe_mul(F, &s2, &o->y, &s->zzz);
  if (fe_eq(&u1, &u2)) {
    if (fe_eq(&s1, &s2)) xyzz_double(F, s); else xyzz_zero(F, s);
    return;
  }
  fe p, r, pp, ppp, q, t, x3, y3;
  fe_sub(F, &p, &u2, &u1);
  fe_sub(F, &r, &s2, &s1);
  fe_sqr(F, &pp, &p);
  fe_mul(F, &ppp, &p, &pp);
  fe_mul(F, &q, &u1, &pp);
  fe_sqr(F, &x3, &r);
  fe_sub(F, &x3, &x3, &ppp);
  fe_dbl(F, &t, &q);
  fe_sub(F, &x3, &x3, &t);
  fe_sub(F, &t, &q, &x3);
  fe_mul(F, &y3, &r, &t);
  fe_mul(F, &t, &s1, &ppp);
  fe_sub(F, &y3, &y3, &t);
  s->x = x3;
  s->y = y3;
  fe_mul(F, &s->zz, &s->zz, &o->zz);
  fe_mul(F, &s->zz, &s->zz, &pp);
  fe_mul(F, &s->zzz, &s->zzz, &o->zzz);
  fe_mul(F, &s->zzz, &s->zzz, &ppp);
}
static void xyzz_add_affine(const orc_field_t* F, xyzz* b, const aff* p) { /* msm.rs:126-165 */
  if (aff_is_identity(p)) return;
  if (xyzz_is_zero(b)) {
    b->x = p->x;
    b->y = p->y;
    fe_one(F, &b->zz);
    fe_one(F, &b->zzz);
    return;
  }
  fe u2, s2;
  fe_mul(F, &u2, &p->x, &b->zz);
  fe_mul(F, &s2, &p->y, &b->zzz);
  if (fe_eq(&b->x, &u2)) {
    if (fe_eq(&b->y, &s2)) xyzz_double(F, b); else xyzz_zero(F, b);
    return;
  }
  fe pv, r, pp, ppp, q, t, x3, y3;
  fe_sub(F, &pv, &u2, &b->x);
  fe_sub(F, &r, &s2, &b->y);
  fe_sqr(F, &pp, &pv);
  fe_mul(F, &ppp, &pv, &pp);
  fe_mul(F, &q, &b->x, &pp);
  fe_sqr(F, &x3, &r);
  fe_sub(F, &x3, &x3, &ppp);
  fe_dbl(F, &t, &q);
  fe_sub(F, &x3, &x3, &t);
  fe_sub(F, &t, &q, &x3);
  fe_mul(F, &y3, &r, &t);
  fe_mul(F, &t, &b->y, &ppp);
  fe_sub(F, &y3, &y3, &t);
  b->x = x3;
  b->y = y3;
  fe_mul(F, &b->zz, &b->zz, &pp);
  fe_mul(F, &b->zzz, &b->zzz, &ppp);
}
static void xyzz_neg(const orc_field_t* F, xyzz* b) { fe_neg(F, &b->y, &b->y); }
/* msm.rs:172-183 bucket_to_curve: affine via two inversions; identity -> (0,0) */
static void xyzz_to_affine(const orc_field_t* F, aff* out, const xyzz* b) {
  if (xyzz_is_zero(b)) { memset(out, 0, sizeof(aff)); return; }
  fe zi, zzi;
  fe_inv(F, &zi, &b->zz);
  fe_inv(F, &zzi, &b->zzz);
  fe_mul(F, &out->x, &b->x, &zi);
  fe_mul(F, &out->y, &b->y, &zzi);
}
static void aff_neg(const orc_field_t* F, aff* r, const aff* p) {
  r->x = p->x;
  fe_neg(F, &r->y, &p->y);
}

/* ------------------------------------------------------------------ curves ---------------- */
typedef struct { const orc_field_t* base; const orc_field_t* scalar; } curve_t;
static int get_curve(int id, curve_t* c) {
  static const int base_of[4] = {1, 0, 2, 3}, scalar_of[4] = {0, 1, 3, 2};
  if (id < 0 || id > 3) return 1;
  c->base = &ORC_FIELDS[base_of[id]];
  c->scalar = &ORC_FIELDS[scalar_of[id]];
  return 0;
}

/* scalar helpers (msm.rs:191-211): operate on canonical little-endian limbs */
static uint32_t num_bits4(const uint64_t* c) {
  for (int i = 3; i >= 0; i--)
    if (c[i]) return (uint32_t)(i * 64 + 64 - __builtin_clzll(c[i]));
  return 0;
}

/* [k]P by double-and-add over canonical bits: the definition (msm.rs:422-429 `base * coeff`) */
static void scalar_mul(const orc_field_t* F, xyzz* out, const aff* p, const uint64_t* k) {
  xyzz acc;
  xyzz_zero(F, &acc);
  int nb = (int)num_bits4(k);
  for (int i = nb - 1; i >= 0; i--) {
    xyzz_double(F, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) xyzz_add_affine(F, &acc, p);
  }
  *out = acc;
}

/* ------------------------------------------------------------------ threading ------------- */
typedef void (*chunk_fn)(void* ctx, size_t lo, size_t hi, int tid);
typedef struct { chunk_fn fn; void* ctx; size_t lo, hi; int tid; } job_t;
static void* job_main(void* a) {
  job_t* j = (job_t*)a;
  j->fn(j->ctx, j->lo, j->hi, j->tid);
  return NULL;
}
/* split [0,n) into nthreads contiguous chunks (rayon par_chunks shape, msm.rs:564-575) */
static void par_chunks(size_t n, int nthreads, chunk_fn fn, void* ctx) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  if (nthreads == 1) { fn(ctx, 0, n, 0); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  size_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    size_t lo = (size_t)t * chunk, hi = lo + chunk;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    jobs[t] = (job_t){fn, ctx, lo, hi, t};
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* dynamic variant: threads pull single job indices from a shared counter (rayon-style work
 * stealing is approximated by this; jobs are coarse so contention is nil) */
typedef struct { chunk_fn fn; void* ctx; size_t njobs; size_t* next; int tid; } djob_t;
static void* djob_main(void* a) {
  djob_t* j = (djob_t*)a;
  for (;;) {
    size_t k = __atomic_fetch_add(j->next, 1, __ATOMIC_RELAXED);
    if (k >= j->njobs) break;
    j->fn(j->ctx, k, k + 1, j->tid);
  }
  return NULL;
}
static void par_jobs(size_t njobs, int nthreads, chunk_fn fn, void* ctx) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > njobs) nthreads = njobs ? (int)njobs : 1;
  size_t next = 0;
  if (nthreads == 1) { fn(ctx, 0, njobs, 0); return; }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  djob_t* jobs = (djob_t*)malloc(sizeof(djob_t) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (djob_t){fn, ctx, njobs, &next, t};
    pthread_create(&th[t], NULL, djob_main, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* ------------------------------------------------------------------ MSM pieces ------------ */
typedef struct {
  const orc_field_t* F;
  const aff* bases;
  const uint64_t* u64s;   /* integer scalars, or NULL */
  const size_t* idx;      /* gather indices, or NULL */
  int max_bits;
  xyzz* partial;          /* one per thread */
} small_ctx;

/* msm.rs:432-454 accumulate_bases / msm.rs:505-530 msm_binary / msm.rs:689-708 batch_add */
static void binary_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  small_ctx* c = (small_ctx*)vctx;
  xyzz acc;
  xyzz_zero(c->F, &acc);
  for (size_t i = lo; i < hi; i++) {
    if (c->u64s && c->u64s[i] == 0) continue;
    xyzz_add_affine(c->F, &acc, &c->bases[c->idx ? c->idx[i] : i]);
  }
  c->partial[tid] = acc;
}
/* msm.rs:533-575 msm_10: one window, 2^bits buckets, suffix running sum */
static void msm10_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  small_ctx* c = (small_ctx*)vctx;
  size_t nb = (size_t)1 << c->max_bits;
  xyzz* buckets = (xyzz*)malloc(sizeof(xyzz) * nb);
  for (size_t b = 0; b < nb; b++) xyzz_zero(c->F, &buckets[b]);
  for (size_t i = lo; i < hi; i++) {
    uint64_t s = c->u64s[i];
    if (s == 0) continue;
    xyzz_add_affine(c->F, &buckets[s], &c->bases[i]);
  }
  xyzz result, running;
  xyzz_zero(c->F, &result);
  xyzz_zero(c->F, &running);
  for (size_t b = nb - 1; b >= 1; b--) {
    xyzz_add(c->F, &running, &buckets[b]);
    xyzz_add(c->F, &result, &running);
  }
  free(buckets);
  c->partial[tid] = result;
}
static size_t compute_ln(size_t a) { /* msm.rs:679-686 */
  if (a == 0) return 0;
  size_t lg = 63 - (size_t)__builtin_clzll((unsigned long long)a);
  return lg * 69 / 100;
}
/* msm.rs:577-677 msm_small_rest (serial body) */
static void small_rest_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  small_ctx* cx = (small_ctx*)vctx;
  const orc_field_t* F = cx->F;
  size_t len = hi - lo;
  size_t c = len < 32 ? 3 : compute_ln(len) + 2;
  if (cx->max_bits == 32 || cx->max_bits == 64) c = 8;
  size_t nwin = ((size_t)cx->max_bits + c - 1) / c;
  size_t nb = ((size_t)1 << c) - 1;
  xyzz* buckets = (xyzz*)malloc(sizeof(xyzz) * nb);
  xyzz* wsum = (xyzz*)malloc(sizeof(xyzz) * (nwin ? nwin : 1));
  for (size_t w = 0; w < nwin; w++) {
    size_t w_start = w * c;
    xyzz res, running;
    xyzz_zero(F, &res);
    for (size_t b = 0; b < nb; b++) xyzz_zero(F, &buckets[b]);
    for (size_t i = lo; i < hi; i++) {
      uint64_t s = cx->u64s[i];
      if (s == 0) continue;
      if (s == 1) {
        if (w_start == 0) xyzz_add_affine(F, &res, &cx->bases[i]); /* msm.rs:613-617 */
      } else {
        s >>= w_start;
        s &= ((uint64_t)1 << c) - 1;
        if (s != 0) xyzz_add_affine(F, &buckets[s - 1], &cx->bases[i]);
      }
    }
    xyzz_zero(F, &running);
    for (size_t b = nb; b-- > 0;) {
      xyzz_add(F, &running, &buckets[b]);
      xyzz_add(F, &res, &running);
    }
    wsum[w] = res;
  }
  /* msm.rs:647-661: lowest + fold(high -> low, c doublings each) */
  xyzz total;
  xyzz_zero(F, &total);
  for (size_t w = nwin; w-- > 1;) {
    xyzz_add(F, &total, &wsum[w]);
    for (size_t d = 0; d < c; d++) xyzz_double(F, &total);
  }
  if (nwin) xyzz_add(F, &total, &wsum[0]);
  free(buckets);
  free(wsum);
  cx->partial[tid] = total;
}

/* ---- halo2curves::msm::msm_best stand-in: full-width signed-digit Pippenger -------------- */
typedef struct {
  const orc_field_t* F;
  const aff* bases;
  const uint64_t* canon; /* n x 4 canonical limbs */
  int c, nwin;
  xyzz* partial;
} best_ctx;
static inline uint32_t get_window(const uint64_t* k, int bit, int c) {
  int limb = bit >> 6, sh = bit & 63;
  if (limb >= 4) return 0;
  u128 two = k[limb];
  if (limb + 1 < 4) two |= (u128)k[limb + 1] << 64;
  return (uint32_t)((two >> sh) & (((u128)1 << c) - 1));
}
static void best_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  best_ctx* cx = (best_ctx*)vctx;
  const orc_field_t* F = cx->F;
  int c = cx->c;
  size_t nb = (size_t)1 << (c - 1);
  xyzz* buckets = (xyzz*)malloc(sizeof(xyzz) * nb);
  uint8_t* carries = (uint8_t*)calloc(hi - lo ? hi - lo : 1, 1);
  xyzz total;
  xyzz_zero(F, &total);
  xyzz* wsum = (xyzz*)malloc(sizeof(xyzz) * cx->nwin);
  for (int w = 0; w < cx->nwin; w++) {
    for (size_t b = 0; b < nb; b++) xyzz_zero(F, &buckets[b]);
    for (size_t i = lo; i < hi; i++) {
      uint32_t v = get_window(&cx->canon[4 * i], w * c, c) + carries[i - lo];
      if (v > nb) { /* signed digit v - 2^c, carry to the next window */
        uint32_t mag = ((uint32_t)1 << c) - v;
        carries[i - lo] = 1;
        if (mag) {
          aff np;
          aff_neg(F, &np, &cx->bases[i]);
          xyzz_add_affine(F, &buckets[mag - 1], &np);
        }
      } else {
        carries[i - lo] = 0;
        if (v) xyzz_add_affine(F, &buckets[v - 1], &cx->bases[i]);
      }
    }
    xyzz res, running;
    xyzz_zero(F, &res);
    xyzz_zero(F, &running);
    for (size_t b = nb; b-- > 0;) {
      xyzz_add(F, &running, &buckets[b]);
      xyzz_add(F, &res, &running);
    }
    wsum[w] = res;
  }
  for (int w = cx->nwin - 1; w >= 0; w--) {
    for (int d = 0; d < c; d++) xyzz_double(F, &total);
    xyzz_add(F, &total, &wsum[w]);
  }
  free(buckets);
  free(carries);
  free(wsum);
  cx->partial[tid] = total;
}
/* halo2curves::msm::msm_best stand-in, parallel shape: one global window size c = ln(n) + 2 (the
 * published Pippenger choice), signed digits precomputed per scalar, and the work split into
 * (window, point-slice) jobs so that all host threads stay busy for any n -- this mirrors how
 * msm_best parallelises (over windows, then within a window), and is what the CPU baseline times. */
typedef struct {
  const orc_field_t* F;
  const aff* bases;
  const int32_t* digits; /* [n][nwin] */
  size_t n;
  int c, nwin, nslices;
  xyzz* job_out; /* [nwin][nslices] */
  xyzz* scratch; /* [nthreads][2^(c-1)] per-thread bucket arrays (allocated once) */
} best2_ctx;
static void best2_job(void* vctx, size_t lo, size_t hi, int tid) {
  best2_ctx* cx = (best2_ctx*)vctx;
  const orc_field_t* F = cx->F;
  size_t nb = (size_t)1 << (cx->c - 1);
  xyzz* buckets = cx->scratch + (size_t)tid * nb;
  for (size_t job = lo; job < hi; job++) {
    int w = (int)(job / cx->nslices), sl = (int)(job % cx->nslices);
    size_t per = (cx->n + cx->nslices - 1) / cx->nslices;
    size_t plo = (size_t)sl * per, phi = plo + per < cx->n ? plo + per : cx->n;
    for (size_t b = 0; b < nb; b++) xyzz_zero(F, &buckets[b]);
    for (size_t i = plo; i < phi; i++) {
      int32_t d = cx->digits[i * cx->nwin + w];
      if (d > 0) xyzz_add_affine(F, &buckets[d - 1], &cx->bases[i]);
      else if (d < 0) {
        aff np;
        aff_neg(F, &np, &cx->bases[i]);
        xyzz_add_affine(F, &buckets[-d - 1], &np);
      }
    }
    xyzz res, running;
    xyzz_zero(F, &res);
    xyzz_zero(F, &running);
    for (size_t b = nb; b-- > 0;) {
      xyzz_add(F, &running, &buckets[b]);
      xyzz_add(F, &res, &running);
    }
    cx->job_out[job] = res;
  }
}
typedef struct { const uint64_t* canon; int32_t* digits; int c, nwin; } dig_ctx;
static void digits_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  (void)tid;
  dig_ctx* cx = (dig_ctx*)vctx;
  uint32_t half = (uint32_t)1 << (cx->c - 1);
  for (size_t i = lo; i < hi; i++) {
    uint32_t carry = 0;
    for (int w = 0; w < cx->nwin; w++) {
      uint32_t v = get_window(&cx->canon[4 * i], w * cx->c, cx->c) + carry;
      if (v > half) { cx->digits[i * cx->nwin + w] = (int32_t)v - (int32_t)((uint32_t)1 << cx->c); carry = 1; }
      else { cx->digits[i * cx->nwin + w] = (int32_t)v; carry = 0; }
    }
  }
}
static void msm_best_canon(const curve_t* cv, const uint64_t* canon, const aff* bases, size_t n,
                           int nthreads, xyzz* out) {
  const orc_field_t* F = cv->base;
  xyzz_zero(F, out);
  if (n == 0) return;
  if (nthreads < 1) nthreads = 1;
  int c = n < 32 ? 3 : (int)compute_ln(n) + 2;
  if (c > 16) c = 16;
  int nwin = (cv->scalar->bits + 1 + c - 1) / c;
  int nslices = (2 * nthreads + nwin - 1) / nwin; /* ~2 jobs per thread, pulled dynamically */
  if ((size_t)nslices > n) nslices = (int)n;
  if (nslices < 1) nslices = 1;
  double td = now_s();
  int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * n * nwin);
  dig_ctx dc = {canon, digits, c, nwin};
  par_chunks(n, nthreads, digits_chunk, &dc);
  if (getenv("ORC_DEBUG")) fprintf(stderr, "msm_best: digits %.3f s\n", now_s() - td);
  double t0 = now_s();
  best2_ctx cx = {F, bases, digits, n, c, nwin, nslices, NULL, NULL};
  size_t njobs = (size_t)nwin * nslices;
  cx.job_out = (xyzz*)malloc(sizeof(xyzz) * njobs);
  cx.scratch = (xyzz*)malloc(sizeof(xyzz) * ((size_t)1 << (c - 1)) * nthreads);
  par_jobs(njobs, nthreads, best2_job, &cx);
  double t1 = now_s();
  for (int w = nwin - 1; w >= 0; w--) {
    for (int d = 0; d < c; d++) xyzz_double(F, out);
    for (int sl = 0; sl < nslices; sl++) xyzz_add(F, out, &cx.job_out[(size_t)w * nslices + sl]);
  }
  if (getenv("ORC_DEBUG"))
    fprintf(stderr, "msm_best: c=%d nwin=%d nslices=%d threads=%d  jobs %.3f s  combine %.3f s\n", c, nwin,
            nslices, nthreads, t1 - t0, now_s() - t1);
  free(cx.scratch);
  free(cx.job_out);
  free(digits);
}

/* msm.rs:478-503 msm_small_with_max_num_bits on u64 scalars */
static void msm_small_u64(const curve_t* cv, const uint64_t* s, const aff* bases, size_t n,
                          int max_bits, int nthreads, xyzz* out) {
  const orc_field_t* F = cv->base;
  xyzz_zero(F, out);
  if (n == 0 || max_bits == 0) return; /* msm.rs:487 */
  if (max_bits > 32) {                 /* msm.rs:491-501: convert to field, msm_best */
    uint64_t* canon = (uint64_t*)calloc(n * 4, 8);
    for (size_t i = 0; i < n; i++) canon[4 * i] = s[i];
    msm_best_canon(cv, canon, bases, n, nthreads, out);
    free(canon);
    return;
  }
  /* msm.rs:505-530 / 564-575 / 664-676: "if len > num_threads: chunks of len/num_threads" */
  int nt = nthreads < 1 ? 1 : nthreads;
  if (n <= (size_t)nt) nt = 1;
  small_ctx cx = {F, bases, s, NULL, max_bits, NULL};
  cx.partial = (xyzz*)malloc(sizeof(xyzz) * (nt + 1));
  for (int t = 0; t <= nt; t++) xyzz_zero(F, &cx.partial[t]);
  chunk_fn fn = max_bits == 1 ? binary_chunk : (max_bits <= 10 ? msm10_chunk : small_rest_chunk);
  par_chunks(n, nt, fn, &cx);
  for (int t = 0; t < nt; t++) xyzz_add(F, out, &cx.partial[t]);
  free(cx.partial);
}

/* ------------------------------------------------------------------ exported API ---------- */
#define EXPORT __attribute__((visibility("default")))

EXPORT int orc_fe_op(int fid, int op, const void* a, const void* b, void* out, size_t n) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe *A = (const fe*)a, *B = (const fe*)b;
  fe* R = (fe*)out;
  for (size_t i = 0; i < n; i++) {
    switch (op) {
      case 0: fe_add(F, &R[i], &A[i], &B[i]); break;
      case 1: fe_sub(F, &R[i], &A[i], &B[i]); break;
      case 2: fe_mul(F, &R[i], &A[i], &B[i]); break;
      case 3: fe_inv(F, &R[i], &A[i]); break;
      case 4: fe_to_mont(F, &R[i], &A[i]); break;
      case 5: fe_from_mont(F, &R[i], &A[i]); break;
      case 6: fe_neg(F, &R[i], &A[i]); break;
      default: return 1;
    }
  }
  return 0;
}

/* Jacobian {x,y,z} (96 B, Montgomery) -> affine {x,y} (64 B), identity -> zeros.  Host-side
 * normalisation helper for tests (group(p)/affine(), traits.rs:285-294). */
EXPORT int orc_jacobian_to_affine(int curve, const void* jac, size_t n, void* out) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  const orc_field_t* F = cv.base;
  const fe* J = (const fe*)jac;
  aff* O = (aff*)out;
  for (size_t i = 0; i < n; i++) {
    const fe *X = &J[3 * i], *Y = &J[3 * i + 1], *Z = &J[3 * i + 2];
    if (fe_is_zero(Z)) { memset(&O[i], 0, sizeof(aff)); continue; }
    fe zi, zi2, zi3;
    fe_inv(F, &zi, Z);
    fe_sqr(F, &zi2, &zi);
    fe_mul(F, &zi3, &zi2, &zi);
    fe_mul(F, &O[i].x, X, &zi2);
    fe_mul(F, &O[i].y, Y, &zi3);
  }
  return 0;
}

/* is (x,y) on y^2 = x^3 + b ?  b given as a Montgomery field element */
EXPORT int orc_on_curve(int curve, const void* pt, const void* b_mont) {
  curve_t cv;
  if (get_curve(curve, &cv)) return -1;
  const orc_field_t* F = cv.base;
  const aff* p = (const aff*)pt;
  if (aff_is_identity(p)) return 1;
  fe l, r;
  fe_sqr(F, &l, &p->y);
  fe_sqr(F, &r, &p->x);
  fe_mul(F, &r, &r, &p->x);
  fe_add(F, &r, &r, (const fe*)b_mont);
  return fe_eq(&l, &r);
}

/* naive definition: sum_i [s_i] P_i  (msm.rs:422-429) */
typedef struct { const curve_t* cv; const fe* scalars; const aff* bases; xyzz* partial; } naive_ctx;
static void naive_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  naive_ctx* c = (naive_ctx*)vctx;
  xyzz acc, t;
  xyzz_zero(c->cv->base, &acc);
  for (size_t i = lo; i < hi; i++) {
    fe k;
    fe_from_mont(c->cv->scalar, &k, &c->scalars[i]);
    scalar_mul(c->cv->base, &t, &c->bases[i], k.l);
    xyzz_add(c->cv->base, &acc, &t);
  }
  c->partial[tid] = acc;
}
EXPORT int orc_msm_naive(int curve, const void* scalars, const void* bases, size_t n, int nthreads,
                         void* out_affine) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  if (nthreads < 1) nthreads = 1;
  naive_ctx cx = {&cv, (const fe*)scalars, (const aff*)bases, NULL};
  cx.partial = (xyzz*)malloc(sizeof(xyzz) * nthreads);
  for (int t = 0; t < nthreads; t++) xyzz_zero(cv.base, &cx.partial[t]);
  par_chunks(n, nthreads, naive_chunk, &cx);
  xyzz tot;
  xyzz_zero(cv.base, &tot);
  for (int t = 0; t < nthreads; t++) xyzz_add(cv.base, &tot, &cx.partial[t]);
  free(cx.partial);
  xyzz_to_affine(cv.base, (aff*)out_affine, &tot);
  return 0;
}

/* the msm_best stand-in on Montgomery scalars */
EXPORT int orc_msm_best(int curve, const void* scalars, const void* bases, size_t n, int nthreads,
                        void* out_affine) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  uint64_t* canon = (uint64_t*)malloc((n ? n : 1) * 32);
  for (size_t i = 0; i < n; i++) {
    fe k;
    fe_from_mont(cv.scalar, &k, &((const fe*)scalars)[i]);
    memcpy(&canon[4 * i], k.l, 32);
  }
  xyzz tot;
  msm_best_canon(&cv, canon, (const aff*)bases, n, nthreads, &tot);
  free(canon);
  xyzz_to_affine(cv.base, (aff*)out_affine, &tot);
  return 0;
}

EXPORT int orc_msm_small(int curve, const uint64_t* scalars, const void* bases, size_t n,
                         int max_bits, int nthreads, void* out_affine) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  if (max_bits < 0) { /* msm.rs:469-476 msm_small: bits of the maximum */
    uint64_t mx = 0;
    for (size_t i = 0; i < n; i++)
      if (scalars[i] > mx) mx = scalars[i];
    max_bits = mx ? 64 - __builtin_clzll(mx) : 0;
  }
  xyzz tot;
  msm_small_u64(&cv, scalars, (const aff*)bases, n, max_bits, nthreads, &tot);
  xyzz_to_affine(cv.base, (aff*)out_affine, &tot);
  return 0;
}

EXPORT int orc_batch_add(int curve, const void* bases, const uint64_t* idx, size_t m, int nthreads,
                         void* out_affine) { /* msm.rs:689-708 */
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  if (nthreads < 1) nthreads = 1;
  size_t* ix = (size_t*)malloc((m ? m : 1) * sizeof(size_t));
  for (size_t i = 0; i < m; i++) ix[i] = (size_t)idx[i];
  small_ctx cx = {cv.base, (const aff*)bases, NULL, ix, 1, NULL};
  cx.partial = (xyzz*)malloc(sizeof(xyzz) * nthreads);
  for (int t = 0; t < nthreads; t++) xyzz_zero(cv.base, &cx.partial[t]);
  par_chunks(m, nthreads, binary_chunk, &cx);
  xyzz tot;
  xyzz_zero(cv.base, &tot);
  for (int t = 0; t < nthreads; t++) xyzz_add(cv.base, &tot, &cx.partial[t]);
  free(cx.partial);
  free(ix);
  xyzz_to_affine(cv.base, (aff*)out_affine, &tot);
  return 0;
}

typedef struct { uint64_t key; } cls_t;
typedef struct { const cls_t* cls; const aff* bases; const uint64_t* canon; aff* gb; uint64_t* lc; } gather_ctx;
static void gather_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  (void)tid;
  gather_ctx* g = (gather_ctx*)vctx;
  for (size_t k = lo; k < hi; k++) {
    size_t idx = (size_t)(g->cls[k].key & 0x0FFFFFFFFFFFFFFFull);
    g->gb[k] = g->bases[idx];
    memcpy(&g->lc[4 * k], &g->canon[4 * idx], 32);
  }
}
typedef struct {
  const orc_field_t* S;
  const fe* scalars;
  const aff* bases;
  uint64_t *canon, *ncanon;
  uint8_t* grp;
} classify_ctx;
static void classify_chunk(void* vctx, size_t lo, size_t hi, int tid) {
  (void)tid;
  classify_ctx* c = (classify_ctx*)vctx;
  for (size_t i = lo; i < hi; i++) {
    if (fe_is_zero(&c->scalars[i]) || aff_is_identity(&c->bases[i])) { c->grp[i] = 0xff; continue; } /* :247 */
    fe s, ns, t;
    fe_from_mont(c->S, &s, &c->scalars[i]);
    fe_neg(c->S, &t, &c->scalars[i]);
    fe_from_mont(c->S, &ns, &t);
    memcpy(&c->canon[4 * i], s.l, 32);
    memcpy(&c->ncanon[4 * i], ns.l, 32);
    uint32_t bs = num_bits4(s.l), bn = num_bits4(ns.l);
    c->grp[i] = bs <= 1 ? 0 : bn <= 1 ? 1 : bs <= 8 ? 2 : bn <= 8 ? 3 : bs <= 16 ? 4 : bn <= 16 ? 5
                : bs <= 32 ? 6 : bn <= 32 ? 7 : bs <= 64 ? 8 : bn <= 64 ? 9 : 10;
  }
}
/* msm.rs:225-419: the full dispatcher */
EXPORT int orc_msm(int curve, const void* scalars_v, const void* bases_v, size_t n, int nthreads,
                   void* out_affine) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  const orc_field_t *F = cv.base, *S = cv.scalar;
  const fe* scalars = (const fe*)scalars_v;
  const aff* bases = (const aff*)bases_v;
  xyzz total;
  xyzz_zero(F, &total);
  if (n == 0) { memset(out_affine, 0, 64); return 0; }          /* msm.rs:228-230 */
  if (n <= 16) return orc_msm_naive(curve, scalars_v, bases_v, n, 1, out_affine); /* :233 */
  double T0 = now_s();
  /* Phase 1: classify in parallel (msm.rs:243-279 par_iter().filter_map) */
  cls_t* cls = (cls_t*)malloc(sizeof(cls_t) * n);
  uint64_t* canon = (uint64_t*)malloc(n * 32);   /* canonical s        */
  uint64_t* ncanon = (uint64_t*)malloc(n * 32);  /* canonical of (-s)  */
  uint8_t* grp = (uint8_t*)malloc(n);
  classify_ctx cc = {S, scalars, bases, canon, ncanon, grp};
  par_chunks(n, nthreads, classify_chunk, &cc);
  size_t m = 0;
  for (size_t i = 0; i < n; i++)
    if (grp[i] != 0xff) cls[m++].key = ((uint64_t)i & 0x0FFFFFFFFFFFFFFFull) | ((uint64_t)grp[i] << 60);
  free(grp);
  if (m == 0) { free(cls); free(canon); free(ncanon); memset(out_affine, 0, 64); return 0; }
  /* Phase 2: sort by group, boundaries (msm.rs:287-301) */
  double T1 = now_s();
  size_t bnd[12];
  { /* 11 groups: a counting sort does what par_sort_unstable_by_key + partition_point do */
    size_t cnt[12];
    memset(cnt, 0, sizeof(cnt));
    for (size_t k = 0; k < m; k++) cnt[(cls[k].key >> 60) + 1]++;
    for (int g = 0; g < 11; g++) cnt[g + 1] += cnt[g];
    memcpy(bnd, cnt, sizeof(bnd));
    cls_t* sorted = (cls_t*)malloc(sizeof(cls_t) * m);
    for (size_t k = 0; k < m; k++) sorted[cnt[cls[k].key >> 60]++] = cls[k];
    free(cls);
    cls = sorted;
  }
  double T2 = now_s();
  /* Phase 3 (msm.rs:343-416) */
  aff* gb = (aff*)malloc(sizeof(aff) * m);
  uint64_t* gs = (uint64_t*)malloc(sizeof(uint64_t) * m);
  xyzz part;
  for (int g = 0; g < 10; g++) {
    size_t lo = bnd[g], hi = bnd[g + 1];
    if (lo >= hi) continue;
    int negate = g & 1;
    for (size_t k = lo; k < hi; k++) {
      size_t idx = (size_t)(cls[k].key & 0x0FFFFFFFFFFFFFFFull);
      gb[k - lo] = bases[idx];
      gs[k - lo] = negate ? ncanon[4 * idx] : canon[4 * idx]; /* repr_low_u64, msm.rs:202-211 */
    }
    if (g < 2) { /* unit groups: accumulate_bases (msm.rs:346-356) */
      small_ctx cx = {F, gb, NULL, NULL, 1, NULL};
      int nt = nthreads < 1 ? 1 : nthreads;
      cx.partial = (xyzz*)malloc(sizeof(xyzz) * nt);
      for (int t = 0; t < nt; t++) xyzz_zero(F, &cx.partial[t]);
      par_chunks(hi - lo, nt, binary_chunk, &cx);
      xyzz_zero(F, &part);
      for (int t = 0; t < nt; t++) xyzz_add(F, &part, &cx.partial[t]);
      free(cx.partial);
    } else {
      static const int widths[5] = {0, 8, 16, 32, 64};
      msm_small_u64(&cv, gs, gb, hi - lo, widths[g / 2], nthreads, &part); /* :362-396 */
    }
    if (negate) xyzz_neg(F, &part); /* pos - neg */
    xyzz_add(F, &total, &part);
  }
  if (bnd[10] < bnd[11]) { /* large group -> msm_best (msm.rs:399-411) */
    size_t lo = bnd[10], hi = bnd[11];
    uint64_t* lc = (uint64_t*)malloc((hi - lo) * 32);
    gather_ctx gc = {cls + lo, bases, canon, gb, lc};
    par_chunks(hi - lo, nthreads, gather_chunk, &gc); /* msm.rs:403-410 unzip, in parallel */
    double T3 = now_s();
    if (getenv("ORC_DEBUG")) fprintf(stderr, "classify %.3f sort %.3f gather %.3f\n", T1 - T0, T2 - T1, T3 - T2);
    msm_best_canon(&cv, lc, gb, hi - lo, nthreads, &part);
    xyzz_add(F, &total, &part);
    free(lc);
  }
  free(cls); free(canon); free(ncanon); free(gb); free(gs);
  xyzz_to_affine(F, (aff*)out_affine, &total);
  return 0;
}

/* ------------------------------------------------------------------ R1CS field vectors ---- */
/* T = AZ o BZ - u*CZ - E1 (- E2)   r1cs/mod.rs:614-620, 650-657 */
EXPORT int orc_cross_term(int fid, const void* az, const void* bz, const void* cz, const void* e1,
                          const void* e2, const void* u, size_t n, void* t) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe *A = az, *B = bz, *C = cz, *E1 = e1, *E2 = e2, *U = u;
  fe* T = (fe*)t;
  for (size_t i = 0; i < n; i++) {
    fe ab, uc, r;
    fe_mul(F, &ab, &A[i], &B[i]);
    fe_mul(F, &uc, U, &C[i]);
    fe_sub(F, &r, &ab, &uc);
    fe_sub(F, &r, &r, &E1[i]);
    if (E2) fe_sub(F, &r, &r, &E2[i]);
    T[i] = r;
  }
  return 0;
}
/* out = a + r*b   r1cs/mod.rs:1058-1069 (W fold, E fold) */
EXPORT int orc_axpy(int fid, const void* a, const void* b, const void* r, size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe *A = a, *B = b, *Rr = r;
  fe* O = (fe*)out;
  for (size_t i = 0; i < n; i++) {
    fe t;
    fe_mul(F, &t, Rr, &B[i]);
    fe_add(F, &O[i], &A[i], &t);
  }
  return 0;
}
EXPORT int orc_vec_add(int fid, const void* a, const void* b, size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  for (size_t i = 0; i < n; i++) fe_add(F, &((fe*)out)[i], &((const fe*)a)[i], &((const fe*)b)[i]);
  return 0;
}
/* multilinear.rs:65-84 bind_poly_var_top: in place on the low half */
EXPORT int orc_bind_top(int fid, void* z, size_t n, const void* r) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe* Z = (fe*)z;
  size_t h = n / 2;
  for (size_t i = 0; i < h; i++) {
    fe d, t;
    fe_sub(F, &d, &Z[i + h], &Z[i]);
    fe_mul(F, &t, (const fe*)r, &d);
    fe_add(F, &Z[i], &Z[i], &t);
  }
  return 0;
}

EXPORT int orc_field_from_u64(int fid, const uint64_t* v, size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  for (size_t i = 0; i < n; i++) fe_from_u64(&ORC_FIELDS[fid], &((fe*)out)[i], v[i]);
  return 0;
}

/* ------------------------------------------------------------------ synthetic inputs ------ */
/* SplitMix64 -- OUR generator (identical to oracle/pyref.py SplitMix64), not halo2curves' */
static inline uint64_t splitmix(uint64_t* s) {
  *s += 0x9E3779B97F4A7C15ull;
  uint64_t z = *s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
/* n uniform field elements by the from_uniform rule (64 LE bytes mod p, traits.rs:315-319),
 * Montgomery form out.  Element i consumes PRNG words 8i..8i+7, so the stream matches
 * pyref.SplitMix64(seed).field(p) called n times. */
EXPORT int orc_gen_scalars(int fid, uint64_t seed, size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe r2, r3;
  memcpy(r2.l, F->r2, 32);
  fe_mul(F, &r3, &r2, &r2);
  uint64_t s = seed;
  for (size_t i = 0; i < n; i++) {
    fe lo, hi, a, b;
    for (int k = 0; k < 4; k++) lo.l[k] = splitmix(&s);
    for (int k = 0; k < 4; k++) hi.l[k] = splitmix(&s);
    fe_mul(F, &a, &lo, &r2); /* lo * R */
    fe_mul(F, &b, &hi, &r3); /* hi * 2^256 * R */
    fe_add(F, &((fe*)out)[i], &a, &b);
  }
  return 0;
}
/* bases P_i = [k0]G + i*G, i < n, batch-normalised affine (the construction of
 * curve_property_tests.rs:186-194).  gen = affine generator (Montgomery), k0 canonical limbs. */
EXPORT int orc_gen_bases(int curve, const void* gen, const uint64_t* k0, size_t n, void* out) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  const orc_field_t* F = cv.base;
  const aff* G = (const aff*)gen;
  aff* O = (aff*)out;
  xyzz acc;
  scalar_mul(F, &acc, G, k0);
  const size_t BLK = 4096;
  xyzz* blk = (xyzz*)malloc(sizeof(xyzz) * BLK);
  fe* pref = (fe*)malloc(sizeof(fe) * (BLK + 1));
  for (size_t base = 0; base < n; base += BLK) {
    size_t m = n - base < BLK ? n - base : BLK;
    for (size_t j = 0; j < m; j++) {
      blk[j] = acc;
      xyzz_add_affine(F, &acc, G);
    }
    /* batch-invert zzz (identity entries, zzz = 0, are skipped) */
    fe_one(F, &pref[0]);
    for (size_t j = 0; j < m; j++) {
      if (xyzz_is_zero(&blk[j])) pref[j + 1] = pref[j];
      else fe_mul(F, &pref[j + 1], &pref[j], &blk[j].zzz);
    }
    fe inv;
    fe_inv(F, &inv, &pref[m]);
    for (size_t j = m; j-- > 0;) {
      if (xyzz_is_zero(&blk[j])) { memset(&O[base + j], 0, sizeof(aff)); continue; }
      fe zi3, zi2, t;
      fe_mul(F, &zi3, &inv, &pref[j]);        /* 1/zzz_j */
      fe_mul(F, &inv, &inv, &blk[j].zzz);
      fe_sqr(F, &t, &blk[j].zz);
      fe_sqr(F, &zi2, &zi3);
      fe_mul(F, &zi2, &zi2, &t);              /* 1/zz = zz^2 / zzz^2 */
      fe_mul(F, &O[base + j].x, &blk[j].x, &zi2);
      fe_mul(F, &O[base + j].y, &blk[j].y, &zi3);
    }
  }
  free(blk);
  free(pref);
  return 0;
}
/* sum_i s_i * (k0 + i) mod q, canonical limbs out: the scalar that multiplies G in
 * MSM(s, gen_bases(k0)) -- a size-independent closed form for full-size checks */
EXPORT int orc_dot_index(int fid, const void* scalars, size_t n, const uint64_t* k0, void* out_canon) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe acc, k, one, t;
  memset(&acc, 0, 32);
  fe kc;
  memcpy(kc.l, k0, 32);
  fe_to_mont(F, &k, &kc);
  fe_one(F, &one);
  for (size_t i = 0; i < n; i++) {
    fe_mul(F, &t, &((const fe*)scalars)[i], &k);
    fe_add(F, &acc, &acc, &t);
    fe_add(F, &k, &k, &one);
  }
  fe_from_mont(F, (fe*)out_canon, &acc);
  return 0;
}
/* [k]P for one point, canonical k -> affine */
EXPORT int orc_scalar_mul(int curve, const void* pt, const uint64_t* k, void* out_affine) {
  curve_t cv;
  if (get_curve(curve, &cv)) return 1;
  xyzz r;
  scalar_mul(cv.base, &r, (const aff*)pt, k);
  xyzz_to_affine(cv.base, (aff*)out_affine, &r);
  return 0;
}

/* ------------------------------------------------------------------ sum-check / MLE -------- */
/* eq factor of index id (sumcheck.rs:1233-1251): left[id >> shift] * right[id & mask], or right[id] */
static size_t g_id_mul = 1, g_id_add = 0; /* cyclic sharding of the index (single-threaded use) */
static void eq_factor(const orc_field_t* F, fe* f, const fe* left, const fe* right, int shift, size_t id) {
  id = id * g_id_mul + g_id_add;
  if (!left) { *f = right[id]; return; }
  fe_mul(F, f, &left[id >> shift], &right[id & (((size_t)1 << shift) - 1)]);
}
/* forms 0..10 of include/nova_b200.h b200_sc_eval; restates sumcheck.rs:165-186, 352-443,
 * 900-1213 (the O(N) sums only; claim derivation lives in oracle/pyref.py) */
EXPORT int orc_sc_eval(int fid, int form, const void* a, const void* b, const void* c, size_t len,
                       const void* eql, const void* eqr, int shift, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe *A = a, *B = b, *C = c, *L = eql, *R = eqr;
  fe acc[3], one;
  memset(acc, 0, sizeof(acc));
  fe_one(F, &one);
  size_t h = len / 2;
  size_t count = form == 10 ? len : h;
  for (size_t i = 0; i < count; i++) {
    fe t, u, v, f, da, db, dc, am, bm, cm;
    switch (form) {
      case 0: /* quad_prod */
        fe_mul(F, &t, &A[i], &B[i]); fe_add(F, &acc[0], &acc[0], &t);
        fe_sub(F, &da, &A[h + i], &A[i]); fe_sub(F, &db, &B[h + i], &B[i]);
        fe_mul(F, &t, &da, &db); fe_add(F, &acc[1], &acc[1], &t);
        break;
      case 1: /* linear */
        fe_sub(F, &t, &A[i], &B[i]); fe_add(F, &acc[0], &acc[0], &t);
        fe_add(F, &am, &A[i], &A[i]); fe_sub(F, &am, &am, &A[h + i]);
        fe_add(F, &bm, &B[i], &B[i]); fe_sub(F, &bm, &bm, &B[h + i]);
        fe_sub(F, &t, &am, &bm); fe_add(F, &acc[1], &acc[1], &t);
        break;
      case 2: /* quadratic */
        fe_mul(F, &t, &A[i], &B[i]); fe_add(F, &acc[0], &acc[0], &t);
        fe_add(F, &am, &A[i], &A[i]); fe_sub(F, &am, &am, &A[h + i]);
        fe_add(F, &bm, &B[i], &B[i]); fe_sub(F, &bm, &bm, &B[h + i]);
        fe_mul(F, &t, &am, &bm); fe_add(F, &acc[1], &acc[1], &t);
        break;
      case 3: /* cubic */
        fe_mul(F, &t, &A[i], &B[i]); fe_mul(F, &t, &t, &C[i]); fe_add(F, &acc[0], &acc[0], &t);
        fe_sub(F, &da, &A[h + i], &A[i]); fe_sub(F, &db, &B[h + i], &B[i]); fe_sub(F, &dc, &C[h + i], &C[i]);
        fe_mul(F, &t, &da, &db); fe_mul(F, &t, &t, &dc); fe_add(F, &acc[1], &acc[1], &t);
        fe_sub(F, &am, &A[i], &da); fe_sub(F, &bm, &B[i], &db); fe_sub(F, &cm, &C[i], &dc);
        fe_mul(F, &t, &am, &bm); fe_mul(F, &t, &t, &cm); fe_add(F, &acc[2], &acc[2], &t);
        break;
      case 4: case 5: /* eq * (A*B - C) / eq * (A*B - 1): t(0), t(inf) */
        fe_mul(F, &t, &A[i], &B[i]);
        if (form == 4) fe_sub(F, &t, &t, &C[i]); else fe_sub(F, &t, &t, &one);
        fe_sub(F, &da, &A[h + i], &A[i]); fe_sub(F, &db, &B[h + i], &B[i]);
        fe_mul(F, &u, &da, &db);
        eq_factor(F, &f, L, R, shift, i);
        fe_mul(F, &t, &t, &f); fe_add(F, &acc[0], &acc[0], &t);
        fe_mul(F, &u, &u, &f); fe_add(F, &acc[1], &acc[1], &u);
        break;
      case 6: /* eq * A: t(0) */
        eq_factor(F, &f, L, R, shift, i);
        fe_mul(F, &t, &A[i], &f); fe_add(F, &acc[0], &acc[0], &t);
        break;
      case 7: case 8: /* t(-1) fall-backs */
        fe_add(F, &am, &A[i], &A[i]); fe_sub(F, &am, &am, &A[h + i]);
        fe_add(F, &bm, &B[i], &B[i]); fe_sub(F, &bm, &bm, &B[h + i]);
        fe_mul(F, &t, &am, &bm);
        if (form == 7) { fe_add(F, &cm, &C[i], &C[i]); fe_sub(F, &cm, &cm, &C[h + i]); fe_sub(F, &t, &t, &cm); }
        else fe_sub(F, &t, &t, &one);
        eq_factor(F, &f, L, R, shift, i);
        fe_mul(F, &t, &t, &f); fe_add(F, &acc[0], &acc[0], &t);
        break;
      case 9:
        fe_add(F, &am, &A[i], &A[i]); fe_sub(F, &am, &am, &A[h + i]);
        eq_factor(F, &f, L, R, shift, i);
        fe_mul(F, &t, &am, &f); fe_add(F, &acc[0], &acc[0], &t);
        break;
      case 10: /* dot with the eq factor */
        eq_factor(F, &f, L, R, shift, i);
        fe_mul(F, &t, &A[i], &f); fe_add(F, &acc[0], &acc[0], &t);
        break;
      default: return 1;
    }
    (void)v;
  }
  int nout = form == 3 ? 3 : (form == 6 || form >= 7) ? 1 : 2;
  memcpy(out, acc, 32 * nout);
  return 0;
}
/* eq.rs:54-73 evals_from_points, the doubling algorithm verbatim */
EXPORT int orc_eq_table(int fid, const void* r, int ell, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe* ev = (fe*)out;
  size_t n = (size_t)1 << ell, size = 1;
  memset(ev, 0, n * 32);
  fe_one(F, &ev[0]);
  for (int k = ell - 1; k >= 0; k--) {
    const fe* rk = &((const fe*)r)[k];
    for (size_t i = 0; i < size; i++) {
      fe_mul(F, &ev[size + i], &ev[i], rk);
      fe_sub(F, &ev[i], &ev[i], &ev[size + i]);
    }
    size *= 2;
  }
  return 0;
}
/* multilinear.rs:98-127 evaluate_with: two sqrt-sized tables, row dots, then the column dot */
EXPORT int orc_mle_eval(int fid, const void* z, int ell, const void* r, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  int s_right = ell / 2, s_left = ell - s_right;
  size_t nl = (size_t)1 << s_left, nr = (size_t)1 << s_right;
  fe* el = (fe*)malloc(nl * 32);
  fe* er = (fe*)malloc(nr * 32);
  orc_eq_table(fid, r, s_left, el);
  orc_eq_table(fid, (const fe*)r + s_left, s_right, er);
  fe acc, t;
  memset(&acc, 0, 32);
  const fe* Z = (const fe*)z;
  for (size_t i = 0; i < nl; i++) {
    fe row;
    memset(&row, 0, 32);
    for (size_t j = 0; j < nr; j++) {
      fe_mul(F, &t, &Z[i * nr + j], &er[j]);
      fe_add(F, &row, &row, &t);
    }
    fe_mul(F, &t, &el[i], &row);
    fe_add(F, &acc, &acc, &t);
  }
  free(el);
  free(er);
  memcpy(out, &acc, 32);
  return 0;
}
/* spartan/mod.rs:119-145 batch_invert_serial; returns 2 when the product is zero (Err) */
EXPORT int orc_batch_invert(int fid, const void* in, size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe* V = (const fe*)in;
  fe* O = (fe*)out;
  fe* prod = (fe*)malloc((n ? n : 1) * 32);
  fe acc;
  fe_one(F, &acc);
  for (size_t i = 0; i < n; i++) { prod[i] = acc; fe_mul(F, &acc, &acc, &V[i]); }
  if (fe_is_zero(&acc)) { free(prod); return 2; }
  fe_inv(F, &acc, &acc);
  for (size_t i = n; i-- > 0;) {
    fe tmp;
    fe_mul(F, &tmp, &acc, &V[i]);
    fe_mul(F, &O[i], &prod[i], &acc);
    acc = tmp;
  }
  free(prod);
  return 0;
}
/* spartan/mod.rs:232-277 / hyperkzg.rs:1028-1040: out = sum_k coeff_k * P_k (zero-extended) */
EXPORT int orc_rlc(int fid, const void* const* polys, const size_t* lens, size_t k, const void* coeffs,
                   size_t n, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe* O = (fe*)out;
  memset(O, 0, n * 32);
  for (size_t j = 0; j < k; j++)
    for (size_t i = 0; i < lens[j]; i++) {
      fe t;
      fe_mul(F, &t, &((const fe*)coeffs)[j], &((const fe*)polys[j])[i]);
      fe_add(F, &O[i], &O[i], &t);
    }
  return 0;
}
/* hyperkzg.rs:1085-1095 */
EXPORT int orc_kzg_fold(int fid, const void* p, size_t n, const void* x, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe* P = (const fe*)p;
  for (size_t j = 0; j < n / 2; j++) {
    fe d, t;
    fe_sub(F, &d, &P[2 * j + 1], &P[2 * j]);
    fe_mul(F, &t, (const fe*)x, &d);
    fe_add(F, &((fe*)out)[j], &t, &P[2 * j]);
  }
  return 0;
}
/* hyperkzg.rs:1011-1019 Horner */
EXPORT int orc_poly_eval(int fid, const void* f, size_t n, const void* us, size_t nu, void* evals) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  for (size_t q = 0; q < nu; q++) {
    fe acc;
    memset(&acc, 0, 32);
    for (size_t i = n; i-- > 0;) {
      fe_mul(F, &acc, &acc, &((const fe*)us)[q]);
      fe_add(F, &acc, &acc, &((const fe*)f)[i]);
    }
    ((fe*)evals)[q] = acc;
  }
  return 0;
}
/* hyperkzg.rs:950-960 (the serial definition): h[n-2] = f[n-1]; h[i-1] = f[i] + u*h[i] */
EXPORT int orc_poly_div(int fid, const void* f, size_t n, const void* u, void* out) {
  if (fid < 0 || fid > 3 || n == 0) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe* Fc = (const fe*)f;
  fe* H = (fe*)out;
  if (n == 1) return 0;
  H[n - 2] = Fc[n - 1];
  for (size_t i = n - 2; i >= 1; i--) {
    fe t;
    fe_mul(F, &t, (const fe*)u, &H[i]);
    fe_add(F, &H[i - 1], &Fc[i], &t);
  }
  return 0;
}
/* r1cs/sparse.rs:19-230: classification into +1 / -1 / small +-2..7 / general and the row loop */
EXPORT int orc_spmv(int fid, const void* data, const uint64_t* indices, const uint64_t* indptr,
                    size_t rows, const void* z, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  const fe* D = (const fe*)data;
  const fe* Z = (const fe*)z;
  fe small[8], nsmall[8];
  for (uint64_t k = 1; k <= 7; k++) { fe_from_u64(F, &small[k], k); fe_neg(F, &nsmall[k], &small[k]); }
  for (size_t r = 0; r < rows; r++) {
    fe sum;
    memset(&sum, 0, 32);
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; e++) {
      const fe* x = &Z[indices[e]];
      int code = 0;
      for (int k = 1; k <= 7; k++) {
        if (fe_eq(&D[e], &small[k])) code = k;
        if (fe_eq(&D[e], &nsmall[k])) code = -k;
      }
      fe t;
      if (code == 0) {
        fe_mul(F, &t, &D[e], x);                 /* sparse.rs:155-158 general */
      } else {                                   /* sparse.rs:110-133 small_mul (1 = plain add) */
        int a = code < 0 ? -code : code;
        fe d;
        fe_dbl(F, &d, x);
        switch (a) {
          case 1: t = *x; break;
          case 2: t = d; break;
          case 3: fe_add(F, &t, &d, x); break;
          case 4: fe_dbl(F, &t, &d); break;
          case 5: fe_dbl(F, &t, &d); fe_add(F, &t, &t, x); break;
          case 6: fe_dbl(F, &t, &d); fe_add(F, &t, &t, &d); break;
          default: fe_dbl(F, &t, &d); fe_add(F, &t, &t, &d); fe_add(F, &t, &t, x); break;
        }
        if (code < 0) fe_neg(F, &t, &t);
      }
      fe_add(F, &sum, &sum, &t);
    }
    ((fe*)out)[r] = sum;
  }
  return 0;
}

/* ---- threaded forms of the R1CS field kernels for the CPU baseline: the reference runs these
 * under rayon (sparse.rs:203-212 into_par_iter over rows; r1cs/mod.rs:614-620 par_iter) ---------- */
typedef struct { int fid; const void* data; const uint64_t* indices; const uint64_t* indptr; const void* z; void* out; } spmv_par_ctx;
static void spmv_par_chunk(void* v, size_t lo, size_t hi, int tid) {
  (void)tid;
  spmv_par_ctx* c = (spmv_par_ctx*)v;
  /* rows [lo,hi): reuse the serial row loop on a shifted view */
  orc_spmv(c->fid, c->data, c->indices, c->indptr + lo, hi - lo, c->z, (fe*)c->out + lo);
}
EXPORT int orc_spmv_par(int fid, const void* data, const uint64_t* indices, const uint64_t* indptr,
                        size_t rows, const void* z, void* out, int nthreads) {
  spmv_par_ctx c = {fid, data, indices, indptr, z, out};
  par_chunks(rows, nthreads, spmv_par_chunk, &c);
  return 0;
}
typedef struct { int fid, op; const void *a, *b, *c, *e1, *e2, *u; void* out; } vec_par_ctx;
static void vec_par_chunk(void* v, size_t lo, size_t hi, int tid) {
  (void)tid;
  vec_par_ctx* c = (vec_par_ctx*)v;
  size_t n = hi - lo;
#define OFF(p) ((p) ? (const void*)((const fe*)(p) + lo) : NULL)
  if (c->op == 0) orc_cross_term(c->fid, OFF(c->a), OFF(c->b), OFF(c->c), OFF(c->e1), OFF(c->e2), c->u, n, (fe*)c->out + lo);
  else if (c->op == 1) orc_axpy(c->fid, OFF(c->a), OFF(c->b), c->u, n, (fe*)c->out + lo);
  else orc_vec_add(c->fid, OFF(c->a), OFF(c->b), n, (fe*)c->out + lo);
#undef OFF
}
EXPORT int orc_vec_par(int fid, int op, const void* a, const void* b, const void* c, const void* e1,
                       const void* e2, const void* u, size_t n, void* out, int nthreads) {
  vec_par_ctx x = {fid, op, a, b, c, e1, e2, u, out};
  par_chunks(n, nthreads, vec_par_chunk, &x);
  return 0;
}

/* spartan/mod.rs:497-534 compute_eval_table_sparse, one matrix: M_evals[col] += rx[row] * val */
EXPORT int orc_spmv_t(int fid, const void* data, const uint64_t* indices, const uint64_t* indptr,
                      size_t rows, const void* rx, size_t out_len, void* out) {
  if (fid < 0 || fid > 3) return 1;
  const orc_field_t* F = &ORC_FIELDS[fid];
  fe* O = (fe*)out;
  memset(O, 0, out_len * 32);
  for (size_t r = 0; r < rows; r++)
    for (uint64_t e = indptr[r]; e < indptr[r + 1]; e++) {
      fe t;
      fe_mul(F, &t, &((const fe*)rx)[r], &((const fe*)data)[e]);
      fe_add(F, &O[indices[e]], &O[indices[e]], &t);
    }
  return 0;
}

/* sharded form: local index j stands for global index j*id_mul + id_add in the eq weight */
EXPORT int orc_sc_eval_sharded(int fid, int form, const void* a, const void* b, const void* c, size_t len,
                               const void* eql, const void* eqr, int shift, size_t id_mul, size_t id_add,
                               void* out) {
  g_id_mul = id_mul;
  g_id_add = id_add;
  int rc = orc_sc_eval(fid, form, a, b, c, len, eql, eqr, shift, out);
  g_id_mul = 1;
  g_id_add = 0;
  return rc;
}
