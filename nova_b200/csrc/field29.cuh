// Carry-free 256-bit Montgomery arithmetic on 9 x 29-bit limbs (the MSM's inner arithmetic).
//
// Why: on B200 the carry-in multiply IMAD.WIDE.U32.X issues at half the rate of the plain
// IMAD.WIDE.U32 (profiles/r01b_microbench.md), which caps the 8x32-bit carry-chain multiplier of
// field.cuh at ~64 G mul/s.  With 29-bit limbs every partial product is < 2^58, so a whole column
// of the schoolbook product (<= 9 terms from a*b, <= 9 from m*p, plus a carry) fits a 64-bit
// accumulator: all 162 multiplies are plain `mad.wide.u32` (full rate) and no carry flag exists
// anywhere.  Plain C++ -- the same code runs on the host for the CPU unit tests.
//
// Representation: value = sum l[i] * 2^(29 i), Montgomery radix R' = 2^261 (x is stored as
// x * 2^261 mod p).  Values are kept LAZILY reduced: a multiplication accepts operands whose
// VALUES satisfy a*b < 128 p^2 (2^261 / p >= 128 for all four moduli) and whose LIMBS are
// < 2^30.1, and returns a value < 2p with exactly normalised limbs.  Additions are limb-wise;
// subtractions add a multiple of p written with "borrowed" limbs (every limb >= 2^31 - 4) so no
// limb goes negative, then run one parallel carry pass.  Callers (curve29.cuh) document the
// bound of every intermediate.
#pragma once
#include <cstdint>
#include "field.cuh"

namespace nova {

struct fe29_t {
  uint32_t l[9];
};

constexpr uint32_t M29 = (1u << 29) - 1;

// c += a * b with a, b 32-bit: exactly one IMAD.WIDE.U32 on the device (written as PTX because the
// C form `c += (uint64_t)a * CONST` makes ptxas emit an extra add of the constant's zero high half)
NOVA_HD void mac(uint64_t& c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
#else
  c += (uint64_t)a * b;
#endif
}

#if defined(__CUDA_ARCH__)
#define NOVA_MULFN __device__ __noinline__
#else
#define NOVA_MULFN inline
#endif

template <class F>
NOVA_HD fe29_t f29_zero() {
  fe29_t r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = 0;
  return r;
}
template <class F>
NOVA_HD fe29_t f29_one() {
  fe29_t r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = F::one29(i);
  return r;
}
NOVA_HD bool f29_is_literal_zero(const fe29_t& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) o |= a.l[i];
  return o == 0;
}

// one parallel carry pass: limbs < 2^32 in, limbs < 2^29 + 8 out (top limb keeps its excess)
NOVA_HD void f29_carry(fe29_t& a) {
  uint32_t c[9];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c[i] = a.l[i] >> 29;
    a.l[i] &= M29;
  }
#pragma unroll
  for (int i = 1; i < 9; i++) a.l[i] += c[i - 1];
}

// exact normalisation (sequential carries): every limb < 2^29 except possibly the top
NOVA_HD void f29_carry_exact(fe29_t& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a.l[i + 1] += a.l[i] >> 29;
    a.l[i] &= M29;
  }
}

// r = a + b (limb-wise; callers keep limb sums < 2^30.1 when the result feeds a multiplication)
NOVA_HD fe29_t f29_add(const fe29_t& a, const fe29_t& b) {
  fe29_t r;
#pragma unroll
  for (int i = 0; i < 9; i++) r.l[i] = a.l[i] + b.l[i];
  return r;
}
NOVA_HD fe29_t f29_dbl(const fe29_t& a) { return f29_add(a, a); }

// r = a - b + K*p, K in {2,4,8}; needs value(b) < K*p, limbs(b) < 2^31 - 4, limbs(a) < 2^30.
// Result: value < value(a) + K*p, limbs < 2^29 + 8.
template <class F, int K>
NOVA_HD fe29_t f29_sub(const fe29_t& a, const fe29_t& b) {
  fe29_t r;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t kp = K == 2 ? F::sub2p(i) : (K == 4 ? F::sub4p(i) : F::sub8p(i));
    r.l[i] = a.l[i] + kp - b.l[i];
  }
  f29_carry(r);
  return r;
}
template <class F, int K>
NOVA_HD fe29_t f29_neg(const fe29_t& a) {
  return f29_sub<F, K>(f29_zero<F>(), a);
}

// Montgomery product a*b / 2^261 mod p.  value(a)*value(b) < 128 p^2, limbs < 2^30.1.
// Output: value < 2p, limbs exactly normalised (< 2^29; top limb < 2^24).
// Called, not inlined, on the device: the by-value ABI passes all 27 words in registers, and one
// copy of the ~250-instruction body keeps the MSM kernels inside the 32 KB instruction cache
// (fully inlined, a mixed add is ~64 KB of code and the kernel becomes fetch-bound).
template <class F>
NOVA_MULFN fe29_t f29_mul(fe29_t a, fe29_t b) {
  uint64_t c[18];
#pragma unroll
  for (int k = 0; k < 18; k++) c[k] = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j < 9; j++) mac(c[i + j], a.l[i], b.l[j]);
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t m = ((uint32_t)c[i] * F::INV29) & M29;
#pragma unroll
    for (int j = 0; j < 9; j++) mac(c[i + j], m, F::p29(j));
    c[i + 1] += c[i] >> 29;  // low 29 bits of c[i] are now zero
  }
  fe29_t r;
#pragma unroll
  for (int k = 9; k < 17; k++) {
    r.l[k - 9] = (uint32_t)c[k] & M29;
    c[k + 1] += c[k] >> 29;
  }
  r.l[8] = (uint32_t)c[17];
  return r;
}

// a^2: off-diagonal products once with a doubled operand (45 + 81 multiplies instead of 162)
template <class F>
NOVA_MULFN fe29_t f29_sqr(fe29_t a) {
  uint64_t c[18];
#pragma unroll
  for (int k = 0; k < 18; k++) c[k] = 0;
  uint32_t d[9];
#pragma unroll
  for (int i = 0; i < 9; i++) d[i] = a.l[i] << 1;  // limbs < 2^30.1 -> < 2^31.1, still 32-bit
#pragma unroll
  for (int i = 0; i < 9; i++) {
    mac(c[2 * i], a.l[i], a.l[i]);
#pragma unroll
    for (int j = i + 1; j < 9; j++) mac(c[i + j], a.l[i], d[j]);
  }
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint32_t m = ((uint32_t)c[i] * F::INV29) & M29;
#pragma unroll
    for (int j = 0; j < 9; j++) mac(c[i + j], m, F::p29(j));
    c[i + 1] += c[i] >> 29;
  }
  fe29_t r;
#pragma unroll
  for (int k = 9; k < 17; k++) {
    r.l[k - 9] = (uint32_t)c[k] & M29;
    c[k + 1] += c[k] >> 29;
  }
  r.l[8] = (uint32_t)c[17];
  return r;
}

// Is value(a) == 0 mod p ?  a must have been through f29_carry (low limb exact) and value < 16p.
// Cheap filter on the low limb (k = a_0 * p^-1 mod 2^29 must be a small multiple count), exact
// comparison against k*p only when the filter fires (probability ~2^-25 for random values).
template <class F>
NOVA_HD bool f29_is_zero_modp(const fe29_t& a) {
  uint32_t k = (a.l[0] * F::PINV29) & M29;
  if (k >= 16) return false;
  fe29_t x = a;
  f29_carry_exact(x);
  uint64_t carry = 0;
  bool eq = true;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    uint64_t t = (uint64_t)k * F::p29(i) + carry;
    uint32_t limb = i < 8 ? ((uint32_t)t & M29) : (uint32_t)t;
    carry = t >> 29;
    eq = eq && (limb == x.l[i]);
  }
  return eq;
}

// ---- conversions with the boundary format (8 x u32 words, Montgomery R = 2^256, canonical) ----
// unpack 256 bits into 9 limbs (no arithmetic)
NOVA_HD fe29_t f29_unpack(const fe_t& w) {
  fe29_t r;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    int bit = 29 * i, word = bit >> 5, sh = bit & 31;
    uint64_t two = w.l[word];
    if (word + 1 < 8) two |= (uint64_t)w.l[word + 1] << 32;
    r.l[i] = (uint32_t)(two >> sh) & M29;
  }
  return r;
}
// pack exactly-normalised limbs of a value < 2^256 into 8 words
NOVA_HD fe_t f29_pack(const fe29_t& a) {
  fe_t w;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    // word k holds bits [32k, 32k+32)
    int lo_limb = (32 * k) / 29, sh = (32 * k) % 29;
    uint64_t v = (uint64_t)a.l[lo_limb] >> sh;
    int have = 29 - sh;
    v |= (uint64_t)a.l[lo_limb + 1] << have;
    if (have + 29 < 32 && lo_limb + 2 < 9) v |= (uint64_t)a.l[lo_limb + 2] << (have + 29);
    w.l[k] = (uint32_t)v;
  }
  return w;
}
// canonical value in [0, p): value < 4p, limbs normalised by f29_mul/f29_carry
template <class F>
NOVA_HD fe29_t f29_canonical(fe29_t a) {
  f29_carry_exact(a);
#pragma unroll
  for (int pass = 0; pass < 3; pass++) {
    // t = a - p with borrow; keep if non-negative
    fe29_t t;
    int64_t br = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      int64_t v = (int64_t)a.l[i] - (int64_t)F::p29(i) + br;
      if (i < 8) {
        t.l[i] = (uint32_t)v & M29;
        br = v >> 29;  // arithmetic shift: -1 on borrow
      } else {
        t.l[i] = (uint32_t)v;
        br = v < 0 ? -1 : 0;
      }
    }
    if (br == 0) a = t;
  }
  return a;
}
// boundary (std Montgomery, canonical) -> internal R' form, value < 2p
template <class F>
NOVA_HD fe29_t f29_from_std(const fe_t& w) {
  fe29_t c;
#pragma unroll
  for (int i = 0; i < 9; i++) c.l[i] = F::to29(i);
  return f29_mul<F>(f29_unpack(w), c);
}
// internal (value < 8p) -> boundary canonical std Montgomery
template <class F>
NOVA_HD fe_t f29_to_std(const fe29_t& a) {
  fe29_t c;
#pragma unroll
  for (int i = 0; i < 9; i++) c.l[i] = F::from29(i);
  return f29_pack(f29_canonical<F>(f29_mul<F>(a, c)));
}
// "raw" 256-bit storage of an internal value (canonical, still in R' form): what the window tables hold
template <class F>
NOVA_HD fe_t f29_store_raw(const fe29_t& a) {
  return f29_pack(f29_canonical<F>(a));
}
NOVA_HD fe29_t f29_load_raw(const fe_t& w) { return f29_unpack(w); }

}  // namespace nova
