// 256-bit Montgomery prime-field arithmetic for sm_100a (K1 of SURVEY.md §2).
//
// Replaces, on the device, what the reference gets from the third-party crate
// halo2curves 0.9.0 (bn256::{Fr,Fq}, pasta::{Fp,Fq}; re-exported at
// src/provider/bn256_grumpkin.rs:26-33 and src/provider/pasta.rs:24-31).
// In-memory layout at the boundary is halo2curves': 4 x u64 little-endian limbs in
// Montgomery form, R = 2^256  ==  8 x u32 little-endian limbs here.
//
// Design: 8 x 32-bit limbs per element, one element per thread, everything in registers.
// The multiplier is a CIOS Montgomery product whose partial products are split over two
// accumulators by the parity of their limb position, so that every (lo,hi) pair of one
// 32x32 product lands in adjacent limbs of ONE carry chain -- ptxas turns each
// mad.lo.cc/madc.hi.cc pair into a single IMAD.WIDE.U32(.X).  Each chain is one asm
// statement (the carry flag never crosses a statement boundary).  Every chain has a
// bit-exact host emulation so the algorithm is unit-tested on CPU (tests/test_host_field.py).
#pragma once
#include <cstdint>
#include "field_constants.cuh"

#if defined(__CUDACC__)
#define NOVA_HD __host__ __device__ __forceinline__
#define NOVA_D __device__ __forceinline__
#else
#define NOVA_HD inline
#define NOVA_D inline
#endif

namespace nova {

struct alignas(16) fe_t {
  uint32_t l[8];
};

// ---------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------

// X[OFF .. OFF+7] += (x0,x1,x2,x3) * y  with product k occupying limbs (OFF+2k, OFF+2k+1);
// X[OFF+8] += carry-out.  If CIN, the chain starts with carry-in = carry32(ca + cb).
template <int OFF, bool CIN, int N>
NOVA_HD void chain_mad(uint32_t (&X)[N], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                       uint32_t y, uint32_t ca = 0, uint32_t cb = 0) {
  static_assert(OFF + 8 < N, "accumulator too short");
#ifdef __CUDA_ARCH__
  if constexpr (CIN) {
    asm("{\n\t"
        ".reg .u32 t;\n\t"
        "add.cc.u32 t, %14, %15;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        "}"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2]), "+r"(X[OFF + 3]),
          "+r"(X[OFF + 4]), "+r"(X[OFF + 5]), "+r"(X[OFF + 6]), "+r"(X[OFF + 7]),
          "+r"(X[OFF + 8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y), "r"(ca), "r"(cb));
  } else {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2]), "+r"(X[OFF + 3]),
          "+r"(X[OFF + 4]), "+r"(X[OFF + 5]), "+r"(X[OFF + 6]), "+r"(X[OFF + 7]),
          "+r"(X[OFF + 8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
  }
#else
  uint64_t cf = CIN ? (((uint64_t)ca + cb) >> 32) : 0;
  const uint32_t xs[4] = {x0, x1, x2, x3};
  for (int k = 0; k < 4; k++) {
    uint64_t prod = (uint64_t)xs[k] * y;
    uint64_t t = (uint64_t)X[OFF + 2 * k] + (uint32_t)prod + cf;
    X[OFF + 2 * k] = (uint32_t)t;
    cf = t >> 32;
    t = (uint64_t)X[OFF + 2 * k + 1] + (prod >> 32) + cf;
    X[OFF + 2 * k + 1] = (uint32_t)t;
    cf = t >> 32;
  }
  X[OFF + 8] += (uint32_t)cf;
#endif
}

// r = a + b (8 limbs) ; returns nothing, carry-out impossible for p < 2^255 operands < p
NOVA_HD void add8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
#ifdef __CUDA_ARCH__
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] + b[i] + c;
    r[i] = (uint32_t)t;
    c = t >> 32;
  }
#endif
}

// r = a + b + carry32(ca + cb)
NOVA_HD void add8_cin(uint32_t (&r)[8], const uint32_t* a, const uint32_t* b, uint32_t ca,
                      uint32_t cb) {
#ifdef __CUDA_ARCH__
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "add.cc.u32 t, %24, %25;\n\t"
      "addc.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;\n\t"
      "}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]),
        "r"(ca), "r"(cb));
#else
  uint64_t c = ((uint64_t)ca + cb) >> 32;
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] + b[i] + c;
    r[i] = (uint32_t)t;
    c = t >> 32;
  }
#endif
}

// r = a - b (8 limbs); returns 0xffffffff if the subtraction borrowed, else 0
NOVA_HD uint32_t sub8(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t borrow;
#ifdef __CUDA_ARCH__
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(borrow)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t t = (uint64_t)a[i] - b[i] - c;
    r[i] = (uint32_t)t;
    c = (t >> 32) & 1;
  }
  borrow = c ? 0xffffffffu : 0u;
#endif
  return borrow;
}

template <class F>
NOVA_HD void load_p(uint32_t (&p)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) p[i] = F::p(i);
}

// ---------------------------------------------------------------------------------------
// field operations (all inputs and outputs fully reduced, in [0, p))
// ---------------------------------------------------------------------------------------

template <class F>
NOVA_HD fe_t fe_zero() {
  fe_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
  return r;
}

template <class F>
NOVA_HD fe_t fe_one() {  // Montgomery form of 1
  fe_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = F::r(i);
  return r;
}

NOVA_HD bool fe_is_zero(const fe_t& a) {
  return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}

NOVA_HD bool fe_eq(const fe_t& a, const fe_t& b) {
  return ((a.l[0] ^ b.l[0]) | (a.l[1] ^ b.l[1]) | (a.l[2] ^ b.l[2]) | (a.l[3] ^ b.l[3]) |
          (a.l[4] ^ b.l[4]) | (a.l[5] ^ b.l[5]) | (a.l[6] ^ b.l[6]) | (a.l[7] ^ b.l[7])) == 0;
}

// a < p as an integer: what SerdeObject::read_raw accepts as a coordinate (provider/ptau.rs:381-383)
template <class F>
NOVA_HD bool fe_is_canonical(const fe_t& a) {
  uint32_t p[8], t[8];
  load_p<F>(p);
  return sub8(t, a.l, p) != 0;
}

// conditional subtract of p: r in [0, 2p) -> [0, p)
template <class F>
NOVA_HD void fe_reduce_once(uint32_t (&r)[8]) {
  uint32_t p[8], t[8];
  load_p<F>(p);
  uint32_t borrow = sub8(t, r, p);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

template <class F>
NOVA_HD fe_t fe_add(const fe_t& a, const fe_t& b) {
  fe_t r;
  add8(r.l, a.l, b.l);  // < 2p < 2^256
  fe_reduce_once<F>(r.l);
  return r;
}

template <class F>
NOVA_HD fe_t fe_sub(const fe_t& a, const fe_t& b) {
  fe_t r;
  uint32_t p[8], t[8];
  load_p<F>(p);
  uint32_t borrow = sub8(r.l, a.l, b.l);
  add8(t, r.l, p);
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = borrow ? t[i] : r.l[i];
  return r;
}

template <class F>
NOVA_HD fe_t fe_neg(const fe_t& a) {
  fe_t z = fe_zero<F>();
  return fe_is_zero(a) ? z : fe_sub<F>(z, a);
}

template <class F>
NOVA_HD fe_t fe_dbl(const fe_t& a) {
  return fe_add<F>(a, a);
}

template <class F, int I>
NOVA_HD void mont_round(uint32_t (&E)[17], uint32_t (&O)[17], const fe_t& a, uint32_t bi) {
  // Round I adds a*b[I] and m*p at absolute limb position I.  The accumulator whose
  // (lo,hi) pairs start at position I is E for even I, O for odd I.
  if constexpr ((I & 1) == 0) {
    if constexpr (I == 0)
      chain_mad<I, false>(E, a.l[0], a.l[2], a.l[4], a.l[6], bi);
    else  // carry-in retires position I-1: carry32(E[I-1] + O[I-1])
      chain_mad<I, true>(E, a.l[0], a.l[2], a.l[4], a.l[6], bi, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(O, a.l[1], a.l[3], a.l[5], a.l[7], bi);
    uint32_t m = (E[I] + O[I]) * F::INV;
    chain_mad<I, false>(E, F::p(0), F::p(2), F::p(4), F::p(6), m);
    chain_mad<I + 1, false>(O, F::p(1), F::p(3), F::p(5), F::p(7), m);
  } else {
    chain_mad<I, true>(O, a.l[0], a.l[2], a.l[4], a.l[6], bi, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(E, a.l[1], a.l[3], a.l[5], a.l[7], bi);
    uint32_t m = (E[I] + O[I]) * F::INV;
    chain_mad<I, false>(O, F::p(0), F::p(2), F::p(4), F::p(6), m);
    chain_mad<I + 1, false>(E, F::p(1), F::p(3), F::p(5), F::p(7), m);
  }
}

// r = a * b * R^-1 mod p   (carry-chain form; fe_mul below selects the form the build uses)
template <class F>
NOVA_HD fe_t fe_mul_chain(const fe_t& a, const fe_t& b) {
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = O[i] = 0;
  mont_round<F, 0>(E, O, a, b.l[0]);
  mont_round<F, 1>(E, O, a, b.l[1]);
  mont_round<F, 2>(E, O, a, b.l[2]);
  mont_round<F, 3>(E, O, a, b.l[3]);
  mont_round<F, 4>(E, O, a, b.l[4]);
  mont_round<F, 5>(E, O, a, b.l[5]);
  mont_round<F, 6>(E, O, a, b.l[6]);
  mont_round<F, 7>(E, O, a, b.l[7]);
  fe_t r;
  add8_cin(r.l, &E[8], &O[8], E[7], O[7]);  // < 2p
  fe_reduce_once<F>(r.l);
  return r;
}

// r = (a*b + c*d) * R^-1 mod p with ONE Montgomery reduction: round I adds a*b[I] and c*d[I] before
// the m*p step, 192 wide products instead of the 256 of two fe_mul.  Needs 2p < R (all four moduli
// are < 2^255): the accumulated value stays below (a + c + p) 2^32-ish per round and ends < 2p.
// Every chain is the same chain_mad instantiation mont_round uses; the extra pair only raises the
// small carry counts that land in limb I+8 (<= 6 instead of <= 3).
template <class F, int I>
NOVA_HD void mont_round2(uint32_t (&E)[17], uint32_t (&O)[17], const fe_t& a, uint32_t bi, const fe_t& c,
                         uint32_t di) {
  if constexpr ((I & 1) == 0) {
    if constexpr (I == 0)
      chain_mad<I, false>(E, a.l[0], a.l[2], a.l[4], a.l[6], bi);
    else
      chain_mad<I, true>(E, a.l[0], a.l[2], a.l[4], a.l[6], bi, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(O, a.l[1], a.l[3], a.l[5], a.l[7], bi);
    chain_mad<I, false>(E, c.l[0], c.l[2], c.l[4], c.l[6], di);
    chain_mad<I + 1, false>(O, c.l[1], c.l[3], c.l[5], c.l[7], di);
    uint32_t m = (E[I] + O[I]) * F::INV;
    chain_mad<I, false>(E, F::p(0), F::p(2), F::p(4), F::p(6), m);
    chain_mad<I + 1, false>(O, F::p(1), F::p(3), F::p(5), F::p(7), m);
  } else {
    chain_mad<I, true>(O, a.l[0], a.l[2], a.l[4], a.l[6], bi, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(E, a.l[1], a.l[3], a.l[5], a.l[7], bi);
    chain_mad<I, false>(O, c.l[0], c.l[2], c.l[4], c.l[6], di);
    chain_mad<I + 1, false>(E, c.l[1], c.l[3], c.l[5], c.l[7], di);
    uint32_t m = (E[I] + O[I]) * F::INV;
    chain_mad<I, false>(O, F::p(0), F::p(2), F::p(4), F::p(6), m);
    chain_mad<I + 1, false>(E, F::p(1), F::p(3), F::p(5), F::p(7), m);
  }
}

template <class F>
NOVA_HD fe_t fe_mul2_add(const fe_t& a, const fe_t& b, const fe_t& c, const fe_t& d) {
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = O[i] = 0;
  mont_round2<F, 0>(E, O, a, b.l[0], c, d.l[0]);
  mont_round2<F, 1>(E, O, a, b.l[1], c, d.l[1]);
  mont_round2<F, 2>(E, O, a, b.l[2], c, d.l[2]);
  mont_round2<F, 3>(E, O, a, b.l[3], c, d.l[3]);
  mont_round2<F, 4>(E, O, a, b.l[4], c, d.l[4]);
  mont_round2<F, 5>(E, O, a, b.l[5], c, d.l[5]);
  mont_round2<F, 6>(E, O, a, b.l[6], c, d.l[6]);
  mont_round2<F, 7>(E, O, a, b.l[7], c, d.l[7]);
  fe_t r;
  add8_cin(r.l, &E[8], &O[8], E[7], O[7]);  // < 2p
  fe_reduce_once<F>(r.l);
  return r;
}

// ---------------------------------------------------------------------------------------
// Carry-save variant.  On B200 the carry-IN form IMAD.WIDE.U32.X issues at half the rate of
// the plain IMAD.WIDE.U32 (profiles/r01b_microbench.md), and 103 of the 128 wide products of
// fe_mul above sit inside carry chains.  Here every wide product is its own two-instruction
// chain (carry-out only) and the carry-out is counted into a small per-position counter K[]
// by an addc on the ALU pipe; counters are folded in when their position retires.
// Same even/odd accumulators, same result (bit-exact with fe_mul; tests/test_host_field.py).
// ---------------------------------------------------------------------------------------
NOVA_HD void mad_cs(uint32_t& lo, uint32_t& hi, uint32_t& k, uint32_t x, uint32_t y) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(lo), "+r"(hi), "+r"(k)
      : "r"(x), "r"(y));
#else
  uint64_t acc = ((uint64_t)hi << 32) | lo;
  uint64_t prod = (uint64_t)x * y;
  uint64_t s = acc + prod;
  k += s < prod ? 1u : 0u;
  lo = (uint32_t)s;
  hi = (uint32_t)(s >> 32);
#endif
}

// k_next += carry32(e + o) + kz   (kz in {0,1})
NOVA_HD void retire_cs(uint32_t& k_next, uint32_t e, uint32_t o, uint32_t kz) {
#ifdef __CUDA_ARCH__
  [[maybe_unused]] uint32_t t;  // scratch of the add.cc; only its carry is used
  asm("add.cc.u32 %1, %2, %3;\n\t"
      "addc.u32 %0, %0, %4;"
      : "+r"(k_next), "=r"(t)
      : "r"(e), "r"(o), "r"(kz));
#else
  k_next += (uint32_t)(((uint64_t)e + o) >> 32) + kz;
#endif
}

template <class F, int I>
NOVA_HD void mont_round_cs(uint32_t (&E)[16], uint32_t (&O)[16], uint32_t (&K)[17], const fe_t& a,
                           uint32_t bi) {
  // A: the accumulator whose (lo,hi) pairs start at position I; B: the one starting at I+1
  uint32_t(&A)[16] = (I & 1) ? O : E;
  uint32_t(&B)[16] = (I & 1) ? E : O;
  mad_cs(A[I + 0], A[I + 1], K[I + 2], a.l[0], bi);
  mad_cs(B[I + 1], B[I + 2], K[I + 3], a.l[1], bi);
  mad_cs(A[I + 2], A[I + 3], K[I + 4], a.l[2], bi);
  mad_cs(B[I + 3], B[I + 4], K[I + 5], a.l[3], bi);
  mad_cs(A[I + 4], A[I + 5], K[I + 6], a.l[4], bi);
  mad_cs(B[I + 5], B[I + 6], K[I + 7], a.l[5], bi);
  mad_cs(A[I + 6], A[I + 7], K[I + 8], a.l[6], bi);
  mad_cs(B[I + 7], B[I + 8], K[I + 9], a.l[7], bi);
  uint32_t m = (E[I] + O[I] + K[I]) * F::INV;
  mad_cs(A[I + 0], A[I + 1], K[I + 2], F::p(0), m);
  mad_cs(B[I + 1], B[I + 2], K[I + 3], F::p(1), m);
  mad_cs(A[I + 2], A[I + 3], K[I + 4], F::p(2), m);
  mad_cs(B[I + 3], B[I + 4], K[I + 5], F::p(3), m);
  mad_cs(A[I + 4], A[I + 5], K[I + 6], F::p(4), m);
  mad_cs(B[I + 5], B[I + 6], K[I + 7], F::p(5), m);
  mad_cs(A[I + 6], A[I + 7], K[I + 8], F::p(6), m);
  mad_cs(B[I + 7], B[I + 8], K[I + 9], F::p(7), m);
  // position I now sums to 0 mod 2^32: E[I] + O[I] + K[I] = c * 2^32 with
  // c = carry32(E[I] + O[I]) + (K[I] != 0); c moves to position I+1
  retire_cs(K[I + 1], E[I], O[I], K[I] ? 1u : 0u);
}

template <class F>
NOVA_HD fe_t fe_mul_cs(const fe_t& a, const fe_t& b) {
  uint32_t E[16], O[16], K[17];
#pragma unroll
  for (int i = 0; i < 16; i++) E[i] = O[i] = K[i] = 0;
  K[16] = 0;
  mont_round_cs<F, 0>(E, O, K, a, b.l[0]);
  mont_round_cs<F, 1>(E, O, K, a, b.l[1]);
  mont_round_cs<F, 2>(E, O, K, a, b.l[2]);
  mont_round_cs<F, 3>(E, O, K, a, b.l[3]);
  mont_round_cs<F, 4>(E, O, K, a, b.l[4]);
  mont_round_cs<F, 5>(E, O, K, a, b.l[5]);
  mont_round_cs<F, 6>(E, O, K, a, b.l[6]);
  mont_round_cs<F, 7>(E, O, K, a, b.l[7]);
  // result = positions 8..15 of E + O + K (< 2p < 2^256, so K[16] == 0 and no carry leaves)
  uint32_t hiE[8], hiO[8], hiK[8], t[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    hiE[i] = E[8 + i];
    hiO[i] = O[8 + i];
    hiK[i] = K[8 + i];
  }
  fe_t r;
  add8(t, hiE, hiO);
  add8(r.l, t, hiK);
  fe_reduce_once<F>(r.l);
  return r;
}

template <class F>
NOVA_HD fe_t fe_mul(const fe_t& a, const fe_t& b) {
#ifdef NOVA_MUL_CS
  return fe_mul_cs<F>(a, b);
#else
  return fe_mul_chain<F>(a, b);
#endif
}

// ---------------------------------------------------------------------------------------
// Dedicated squaring (A/B variant, -DNOVA_SQR_DEDICATED; timed on B200: -8 % wide products, +8 % time -> not the default,
// profiles/r02a_variants.md).
// a^2 = sum_i a_i^2 2^(64 i) + 2 sum_{i<j} a_i a_j 2^(32 (i+j)): 28 cross products + 8 squares, then the 64
// products of the Montgomery reduction = 100 wide products instead of 128.  Same even/odd accumulators as the
// multiplier: a_i a_j lands in E when i + j is even, in O when it is odd; for a fixed j the partners i < j of
// each parity form ONE carry chain of 1..4 products whose carry-out limb is still untouched at that point
// (the chains are issued in increasing j).  After doubling both accumulators (plain shifts, ALU pipe) the eight
// squares go into E as one chain, and eight reduce-only rounds use the multiplier's own chains.
// ---------------------------------------------------------------------------------------

// X[OFF .. OFF+2NP-1] += (x0 .. x_{NP-1}) * y, product k on limbs (OFF+2k, OFF+2k+1); X[OFF+2NP] += carry-out
template <int OFF, int NP, int N>
NOVA_HD void chain_mad_n(uint32_t (&X)[N], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
  static_assert(NP >= 1 && NP <= 4 && OFF + 2 * NP < N, "chain does not fit the accumulator");
#ifdef __CUDA_ARCH__
  if constexpr (NP == 1) {
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2])
        : "r"(x0), "r"(y));
  } else if constexpr (NP == 2) {
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32 %4, %4, 0;"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2]), "+r"(X[OFF + 3]), "+r"(X[OFF + 4])
        : "r"(x0), "r"(x1), "r"(y));
  } else if constexpr (NP == 3) {
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32 %6, %6, 0;"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2]), "+r"(X[OFF + 3]), "+r"(X[OFF + 4]),
          "+r"(X[OFF + 5]), "+r"(X[OFF + 6])
        : "r"(x0), "r"(x1), "r"(x2), "r"(y));
  } else {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(X[OFF + 0]), "+r"(X[OFF + 1]), "+r"(X[OFF + 2]), "+r"(X[OFF + 3]), "+r"(X[OFF + 4]),
          "+r"(X[OFF + 5]), "+r"(X[OFF + 6]), "+r"(X[OFF + 7]), "+r"(X[OFF + 8])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
  }
#else
  const uint32_t xs[4] = {x0, x1, x2, x3};
  uint64_t cf = 0;
  for (int k = 0; k < NP; k++) {
    uint64_t prod = (uint64_t)xs[k] * y;
    uint64_t t = (uint64_t)X[OFF + 2 * k] + (uint32_t)prod + cf;
    X[OFF + 2 * k] = (uint32_t)t;
    cf = t >> 32;
    t = (uint64_t)X[OFF + 2 * k + 1] + (prod >> 32) + cf;
    X[OFF + 2 * k + 1] = (uint32_t)t;
    cf = t >> 32;
  }
  X[OFF + 2 * NP] += (uint32_t)cf;
#endif
}

// X[0 .. 15] += a_k^2 on limbs (2k, 2k+1), k = 0..7, as one carry chain; X[16] += carry-out
NOVA_HD void chain_sqr_diag(uint32_t (&X)[17], const fe_t& a) {
#ifdef __CUDA_ARCH__
  asm("mad.lo.cc.u32 %0, %17, %17, %0;\n\t"
      "madc.hi.cc.u32 %1, %17, %17, %1;\n\t"
      "madc.lo.cc.u32 %2, %18, %18, %2;\n\t"
      "madc.hi.cc.u32 %3, %18, %18, %3;\n\t"
      "madc.lo.cc.u32 %4, %19, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %19, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %20, %20, %6;\n\t"
      "madc.hi.cc.u32 %7, %20, %20, %7;\n\t"
      "madc.lo.cc.u32 %8, %21, %21, %8;\n\t"
      "madc.hi.cc.u32 %9, %21, %21, %9;\n\t"
      "madc.lo.cc.u32 %10, %22, %22, %10;\n\t"
      "madc.hi.cc.u32 %11, %22, %22, %11;\n\t"
      "madc.lo.cc.u32 %12, %23, %23, %12;\n\t"
      "madc.hi.cc.u32 %13, %23, %23, %13;\n\t"
      "madc.lo.cc.u32 %14, %24, %24, %14;\n\t"
      "madc.hi.cc.u32 %15, %24, %24, %15;\n\t"
      "addc.u32 %16, %16, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]),
        "+r"(X[8]), "+r"(X[9]), "+r"(X[10]), "+r"(X[11]), "+r"(X[12]), "+r"(X[13]), "+r"(X[14]), "+r"(X[15]),
        "+r"(X[16])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
#else
  uint64_t cf = 0;
  for (int k = 0; k < 8; k++) {
    uint64_t prod = (uint64_t)a.l[k] * a.l[k];
    uint64_t t = (uint64_t)X[2 * k] + (uint32_t)prod + cf;
    X[2 * k] = (uint32_t)t;
    cf = t >> 32;
    t = (uint64_t)X[2 * k + 1] + (prod >> 32) + cf;
    X[2 * k + 1] = (uint32_t)t;
    cf = t >> 32;
  }
  X[16] += (uint32_t)cf;
#endif
}

// cross products of column j: partners i < j with i = j (mod 2) go to E, the others to O
template <int J>
NOVA_HD void sqr_cross(uint32_t (&E)[17], uint32_t (&O)[17], const fe_t& a) {
  constexpr int SAME0 = J & 1;            // first partner of the same parity as J
  constexpr int NSAME = J / 2;            // i = SAME0, SAME0 + 2, ... < J
  constexpr int OTH0 = 1 - (J & 1);
  constexpr int NOTH = (J + 1) / 2;
  const uint32_t y = a.l[J];
  auto limb = [&](int i) { return i < 8 ? a.l[i] : 0u; };
  if constexpr (NSAME > 0)  // positions i + J even
    chain_mad_n<SAME0 + J, NSAME>(E, limb(SAME0), limb(SAME0 + 2), limb(SAME0 + 4), limb(SAME0 + 6), y);
  if constexpr (NOTH > 0)   // positions i + J odd
    chain_mad_n<OTH0 + J, NOTH>(O, limb(OTH0), limb(OTH0 + 2), limb(OTH0 + 4), limb(OTH0 + 6), y);
}

// reduce-only round I: m = (position I incl. the carry retired from position I-1) * INV, then + m p
template <class F, int I>
NOVA_HD void mont_reduce_round(uint32_t (&E)[17], uint32_t (&O)[17]) {
  uint32_t c = 0;
  if constexpr (I > 0) c = (uint32_t)(((uint64_t)E[I - 1] + O[I - 1]) >> 32);
  const uint32_t m = (E[I] + O[I] + c) * F::INV;
  if constexpr ((I & 1) == 0) {
    if constexpr (I == 0)
      chain_mad<I, false>(E, F::p(0), F::p(2), F::p(4), F::p(6), m);
    else
      chain_mad<I, true>(E, F::p(0), F::p(2), F::p(4), F::p(6), m, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(O, F::p(1), F::p(3), F::p(5), F::p(7), m);
  } else {
    chain_mad<I, true>(O, F::p(0), F::p(2), F::p(4), F::p(6), m, E[I - 1], O[I - 1]);
    chain_mad<I + 1, false>(E, F::p(1), F::p(3), F::p(5), F::p(7), m);
  }
}

template <class F>
NOVA_HD fe_t fe_sqr_dedicated(const fe_t& a) {
  uint32_t E[17], O[17];
#pragma unroll
  for (int i = 0; i < 17; i++) E[i] = O[i] = 0;
  sqr_cross<1>(E, O, a);
  sqr_cross<2>(E, O, a);
  sqr_cross<3>(E, O, a);
  sqr_cross<4>(E, O, a);
  sqr_cross<5>(E, O, a);
  sqr_cross<6>(E, O, a);
  sqr_cross<7>(E, O, a);
#pragma unroll
  for (int i = 16; i > 0; i--) {  // double both halves of the cross sum (E + O < 2^511)
    E[i] = (E[i] << 1) | (E[i - 1] >> 31);
    O[i] = (O[i] << 1) | (O[i - 1] >> 31);
  }
  E[0] <<= 1;
  O[0] <<= 1;
  chain_sqr_diag(E, a);
  // T = E + O (< p^2).  The reduce-only chains end with a non-propagating `limb I+8 += carry`, which is only
  // safe while that limb holds a small carry count -- true in the multiplier, not on top of T's data limbs.  So
  // T's high half is set aside first and the reduction runs on the LOW halves of E and O with zeroed high limbs
  // (the invariant of the integrated multiplier is restored: limb I+8 is untouched when round I starts); the
  // high half is added afterwards:  (T + sum m_i p 2^(32 i)) / 2^256 = (E_hi + O_hi) + (E_lo + O_lo + sum m_i p 2^(32 i)) / 2^256.
  uint32_t hiE[8], hiO[8], lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {  // set T's high half aside and reduce the low half in place
    hiE[i] = E[8 + i];
    hiO[i] = O[8 + i];
    E[8 + i] = 0;
    O[8 + i] = 0;
  }
  E[16] = O[16] = 0;  // (E, O <= T < 2^512: limb 16 carried nothing)
  mont_reduce_round<F, 0>(E, O);
  mont_reduce_round<F, 1>(E, O);
  mont_reduce_round<F, 2>(E, O);
  mont_reduce_round<F, 3>(E, O);
  mont_reduce_round<F, 4>(E, O);
  mont_reduce_round<F, 5>(E, O);
  mont_reduce_round<F, 6>(E, O);
  mont_reduce_round<F, 7>(E, O);
  add8_cin(lo, &E[8], &O[8], E[7], O[7]);
  add8(hi, hiE, hiO);  // every partial sum is <= the final value < 2p < 2^256: no carry leaves
  fe_t r;
  add8(r.l, hi, lo);
  fe_reduce_once<F>(r.l);
  return r;
}

template <class F>
NOVA_HD fe_t fe_sqr(const fe_t& a) {
#if defined(NOVA_SQR_DEDICATED) && !defined(NOVA_MUL_CS)
  return fe_sqr_dedicated<F>(a);
#else
  return fe_mul<F>(a, a);
#endif
}

// Montgomery <-> canonical
template <class F>
NOVA_HD fe_t fe_from_mont(const fe_t& a) {
  fe_t one;
#pragma unroll
  for (int i = 0; i < 8; i++) one.l[i] = (i == 0);
  return fe_mul<F>(a, one);
}

template <class F>
NOVA_HD fe_t fe_to_mont(const fe_t& a) {
  fe_t r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.l[i] = F::r2(i);
  return fe_mul<F>(a, r2);
}

// a^(p-2) by square-and-multiply (only used O(1) times per call, never in hot loops)
template <class F>
NOVA_HD fe_t fe_inv(const fe_t& a) {
  uint32_t e[8];
  uint32_t bw = 2;  // e = p - 2 with borrow propagation (Pasta moduli have low limb 1)
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t pi = F::p(i);
    e[i] = pi - bw;
    bw = (pi < bw) ? 1u : 0u;
  }
  fe_t acc = fe_one<F>();
  for (int i = 255; i >= 0; i--) {
    acc = fe_sqr<F>(acc);
    if ((e[i >> 5] >> (i & 31)) & 1) acc = fe_mul<F>(acc, a);
  }
  return acc;
}

// ---------------------------------------------------------------------------------------
// global-memory access: 2 x 128-bit per element
// ---------------------------------------------------------------------------------------
#if defined(__CUDACC__)
NOVA_D fe_t fe_load(const void* base, size_t idx) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * idx;
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  fe_t r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}
// same, but through the normal (coherent) path -- for buffers written earlier in the same kernel
NOVA_D fe_t fe_load_rw(const void* base, size_t idx) {
  const uint4* p = reinterpret_cast<const uint4*>(base) + 2 * idx;
  uint4 lo = p[0], hi = p[1];
  fe_t r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}
NOVA_D void fe_store(void* base, size_t idx, const fe_t& a) {
  uint4* p = reinterpret_cast<uint4*>(base) + 2 * idx;
  p[0] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
  p[1] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
}
#elif defined(NOVA_SIMT_HOST)
// CPU run of the __global__ wrappers (tests/hostcheck/simt_host.h): plain memory accesses
inline fe_t fe_load(const void* base, size_t idx) { return reinterpret_cast<const fe_t*>(base)[idx]; }
inline fe_t fe_load_rw(const void* base, size_t idx) { return reinterpret_cast<const fe_t*>(base)[idx]; }
inline void fe_store(void* base, size_t idx, const fe_t& a) { reinterpret_cast<fe_t*>(base)[idx] = a; }
#endif

}  // namespace nova
