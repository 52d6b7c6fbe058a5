// Point-arithmetic policies for the MSM kernels.
//
//   pa29<F>  XYZZ group law (src/provider/msm.rs:38-165 restated) on the carry-free 9x29-bit
//            multiplier of field29.cuh.  Coordinates are kept lazily reduced; the invariant of a
//            stored point is   x < 8p, y < 4p, zz < 2p, zzz < 2p   with carry-normalised limbs,
//            and every intermediate below is annotated with its bound (in multiples of p); a
//            product needs (bound_a * bound_b) <= 128.
//   pa32<F>  the same interface over the canonical 8x32-bit arithmetic of field.cuh/curve.cuh
//            (kept for A/B measurement: build with -DNOVA_MSM_ARITH32).
//
// Window tables hold affine points as 2 x 256-bit words: pa32 in the boundary format
// (Montgomery R = 2^256), pa29 in its internal Montgomery form (R' = 2^261), canonical.
#pragma once
#include "curve.cuh"
#include "field29.cuh"

namespace nova {

template <class F>
struct pa29 {
  using fe = fe29_t;
  struct pt {
    fe x, y, zz, zzz;
  };
  struct aff {
    fe x, y;
  };
  static constexpr int PT_WORDS = 36;

  static NOVA_HD pt identity() {
    pt r;
    r.x = f29_one<F>();
    r.y = f29_one<F>();
    r.zz = f29_zero<F>();
    r.zzz = f29_zero<F>();
    return r;
  }
  // zz is either the literal zero written by identity() / the P - P branch, or a product of
  // values that are non-zero mod p (never 0 or p), so a literal test is exact
  static NOVA_HD bool is_identity(const pt& p) { return f29_is_literal_zero(p.zz); }
  static NOVA_HD bool aff_is_identity(const aff& a) {
    return f29_is_literal_zero(a.x) && f29_is_literal_zero(a.y);  // table entries are canonical
  }
  static NOVA_HD void neg_aff(aff& a) { a.y = f29_neg<F, 2>(a.y); }  // y canonical (< p) -> < 2p

  // dbl-2008-s-1 (msm.rs:65-88).  In: x<8, y<4, zz<2, zzz<2.
  static
#if defined(__CUDA_ARCH__)
      __device__ __noinline__
#else
      inline
#endif
      void
      dbl(pt& p) {
    if (is_identity(p)) return;
    fe u = f29_dbl(p.y);                       // < 8
    fe v = f29_sqr<F>(u);                      // 64 -> < 2
    fe w = f29_mul<F>(u, v);                   // 16 -> < 2
    fe s = f29_mul<F>(p.x, v);                 // 16 -> < 2
    fe xx = f29_sqr<F>(p.x);                   // 64 -> < 2
    fe m = f29_add(f29_dbl(xx), xx);           // < 6, limbs < 3*2^29: normalise before squaring
    f29_carry(m);
    fe x3 = f29_sub<F, 4>(f29_sqr<F>(m), f29_dbl(s));        // 36 -> <2 ; - (<4) + 4 -> < 6
    fe y3 = f29_sub<F, 2>(f29_mul<F>(m, f29_sub<F, 8>(s, x3)),  // (s - x3 + 8) < 10 ; 6*10 = 60
                          f29_mul<F>(w, p.y));                  // 2*4 ; result < 4
    p.x = x3;
    p.y = y3;
    p.zz = f29_mul<F>(v, p.zz);
    p.zzz = f29_mul<F>(w, p.zzz);
  }

  // madd-2008-s (msm.rs:126-165): acc += a, a a non-identity table point (x canonical, y < 2p)
  static NOVA_HD void madd(pt& acc, const aff& a) {
    if (is_identity(acc)) {
      acc.x = a.x;
      acc.y = a.y;
      acc.zz = f29_one<F>();
      acc.zzz = f29_one<F>();
      return;
    }
    fe u2 = f29_mul<F>(a.x, acc.zz);           // 1*2 -> < 2
    fe s2 = f29_mul<F>(a.y, acc.zzz);          // 2*2 -> < 2
    fe P = f29_sub<F, 8>(u2, acc.x);           // < 10
    fe R = f29_sub<F, 4>(s2, acc.y);           // < 6
    if (f29_is_zero_modp<F>(P)) {              // same x: doubling or inverse (msm.rs:147-154)
      if (f29_is_zero_modp<F>(R))
        dbl(acc);
      else
        acc = identity();
      return;
    }
    fe pp = f29_sqr<F>(P);                     // 100 -> < 2
    fe ppp = f29_mul<F>(P, pp);                // 20 -> < 2
    fe q = f29_mul<F>(acc.x, pp);              // 16 -> < 2
    fe x3 = f29_sub<F, 4>(f29_sub<F, 2>(f29_sqr<F>(R), ppp), f29_dbl(q));  // 36 ; <4 ; -(<4)+4 -> < 8
    fe y3 = f29_sub<F, 2>(f29_mul<F>(R, f29_sub<F, 8>(q, x3)),             // 6*10 = 60
                          f29_mul<F>(acc.y, ppp));                          // 4*2 ; -> < 4
    acc.x = x3;
    acc.y = y3;
    acc.zz = f29_mul<F>(acc.zz, pp);
    acc.zzz = f29_mul<F>(acc.zzz, ppp);
  }

  // add-2008-s (msm.rs:91-123)
  static NOVA_HD void add(pt& acc, const pt& o) {
    if (is_identity(o)) return;
    if (is_identity(acc)) {
      acc = o;
      return;
    }
    fe u1 = f29_mul<F>(acc.x, o.zz);           // 8*2
    fe u2 = f29_mul<F>(o.x, acc.zz);
    fe s1 = f29_mul<F>(acc.y, o.zzz);          // 4*2
    fe s2 = f29_mul<F>(o.y, acc.zzz);
    fe P = f29_sub<F, 2>(u2, u1);              // < 4
    fe R = f29_sub<F, 2>(s2, s1);              // < 4
    if (f29_is_zero_modp<F>(P)) {
      if (f29_is_zero_modp<F>(R))
        dbl(acc);
      else
        acc = identity();
      return;
    }
    fe pp = f29_sqr<F>(P);                     // 16
    fe ppp = f29_mul<F>(P, pp);                // 8
    fe q = f29_mul<F>(u1, pp);                 // 4
    fe x3 = f29_sub<F, 4>(f29_sub<F, 2>(f29_sqr<F>(R), ppp), f29_dbl(q));  // < 8
    fe y3 = f29_sub<F, 2>(f29_mul<F>(R, f29_sub<F, 8>(q, x3)),             // 4*10
                          f29_mul<F>(s1, ppp));                             // < 4
    acc.x = x3;
    acc.y = y3;
    acc.zz = f29_mul<F>(f29_mul<F>(acc.zz, o.zz), pp);
    acc.zzz = f29_mul<F>(f29_mul<F>(acc.zzz, o.zzz), ppp);
  }

  static NOVA_HD pt mul_small(const pt& p, uint32_t k) {
    pt acc = identity();
    if (k == 0 || is_identity(p)) return acc;
    int top = 31;
    while (!((k >> top) & 1)) top--;
    for (int i = top; i >= 0; i--) {
      dbl(acc);
      if ((k >> i) & 1) add(acc, p);
    }
    return acc;
  }

  static NOVA_HD fe inv(const fe& a) {  // a^(p-2); a < 8p
    uint32_t e[8], bw = 2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t pi = F::p(i);
      e[i] = pi - bw;
      bw = (pi < bw) ? 1u : 0u;
    }
    fe acc = f29_one<F>();
    for (int i = 255; i >= 0; i--) {
      acc = f29_sqr<F>(acc);
      if ((e[i >> 5] >> (i & 31)) & 1) acc = f29_mul<F>(acc, a);
    }
    return acc;
  }

  // non-identity XYZZ -> affine in table form (x = X/ZZ, y = Y/ZZZ; 1/ZZ = ZZ^2 (1/ZZZ)^2)
  static NOVA_HD aff to_affine(const pt& q) {
    fe iz3 = inv(q.zzz);
    fe iz2 = f29_mul<F>(f29_sqr<F>(q.zz), f29_sqr<F>(iz3));
    aff a;
    a.x = f29_mul<F>(q.x, iz2);
    a.y = f29_mul<F>(q.y, iz3);
    return a;
  }
  static NOVA_HD aff from_std_affine(const fe_t& x, const fe_t& y) {
    aff a;
    a.x = f29_from_std<F>(x);
    a.y = f29_from_std<F>(y);
    return a;
  }
  // XYZZ -> boundary Jacobian (X', Y', Z') = (X ZZ ZZZ^2, Y ZZ^3 ZZZ^2, ZZ ZZZ); identity -> z = 0
  static NOVA_HD void to_jacobian_std(const pt& p, fe_t& X, fe_t& Y, fe_t& Z) {
    if (is_identity(p)) {
      X = fe_zero<F>();
      Y = fe_one<F>();
      Z = fe_zero<F>();
      return;
    }
    fe zzz2 = f29_sqr<F>(p.zzz);
    fe zz_zzz2 = f29_mul<F>(p.zz, zzz2);
    X = f29_to_std<F>(f29_mul<F>(p.x, zz_zzz2));
    fe zz2 = f29_sqr<F>(p.zz);
    Y = f29_to_std<F>(f29_mul<F>(p.y, f29_mul<F>(zz2, zz_zzz2)));
    Z = f29_to_std<F>(f29_mul<F>(p.zz, p.zzz));
  }
  static NOVA_HD pt from_jacobian_std(const fe_t& X, const fe_t& Y, const fe_t& Z) {
    if (fe_is_zero(Z)) return identity();
    pt p;
    fe z = f29_from_std<F>(Z);
    p.x = f29_from_std<F>(X);
    p.y = f29_from_std<F>(Y);
    p.zz = f29_sqr<F>(z);
    p.zzz = f29_mul<F>(p.zz, z);
    return p;
  }

#if defined(__CUDACC__)
  // window-table entries: 2 x 256-bit canonical words in internal Montgomery form
  static NOVA_D aff load_table(const void* tables, size_t idx) {
    aff a;
    a.x = f29_load_raw(fe_load(tables, 2 * idx));
    a.y = f29_load_raw(fe_load(tables, 2 * idx + 1));
    return a;
  }
  static NOVA_D void store_table(void* tables, size_t idx, const aff& a) {
    fe_store(tables, 2 * idx, f29_store_raw<F>(a.x));
    fe_store(tables, 2 * idx + 1, f29_store_raw<F>(a.y));
  }
  static NOVA_D void store_table_identity(void* tables, size_t idx) {
    fe_store(tables, 2 * idx, fe_zero<F>());
    fe_store(tables, 2 * idx + 1, fe_zero<F>());
  }
  // XYZZ records: 36 words = 9 x uint4
  static NOVA_D pt load(const void* base, size_t idx) {
    const uint4* q = reinterpret_cast<const uint4*>(base) + 9 * idx;
    uint32_t w[36];
#pragma unroll
    for (int k = 0; k < 9; k++) {
      uint4 v = q[k];
      w[4 * k] = v.x;
      w[4 * k + 1] = v.y;
      w[4 * k + 2] = v.z;
      w[4 * k + 3] = v.w;
    }
    pt p;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      p.x.l[i] = w[i];
      p.y.l[i] = w[9 + i];
      p.zz.l[i] = w[18 + i];
      p.zzz.l[i] = w[27 + i];
    }
    return p;
  }
  static NOVA_D pt shfl_down(const pt& p, int d, int width) {
    pt r;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], d, width);
      r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], d, width);
      r.zz.l[i] = __shfl_down_sync(0xffffffffu, p.zz.l[i], d, width);
      r.zzz.l[i] = __shfl_down_sync(0xffffffffu, p.zzz.l[i], d, width);
    }
    return r;
  }
  static NOVA_D void store(void* base, size_t idx, const pt& p) {
    uint32_t w[36];
#pragma unroll
    for (int i = 0; i < 9; i++) {
      w[i] = p.x.l[i];
      w[9 + i] = p.y.l[i];
      w[18 + i] = p.zz.l[i];
      w[27 + i] = p.zzz.l[i];
    }
    uint4* q = reinterpret_cast<uint4*>(base) + 9 * idx;
#pragma unroll
    for (int k = 0; k < 9; k++) q[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
  }
#endif
};

// ---------------------------------------------------------------------------------------------
template <class F>
struct pa32 {
  using fe = fe_t;
  using pt = xyzz_t;
  using aff = affine_t;
  static constexpr int PT_WORDS = 32;
  static NOVA_HD pt identity() { return xyzz_identity<F>(); }
  static NOVA_HD bool is_identity(const pt& p) { return xyzz_is_identity(p); }
  static NOVA_HD bool aff_is_identity(const aff& a) { return affine_is_identity(a); }
  static NOVA_HD void neg_aff(aff& a) { a.y = fe_neg<F>(a.y); }
  static NOVA_HD void dbl(pt& p) { xyzz_dbl<F>(p); }
  static NOVA_HD void madd(pt& acc, const aff& a) { xyzz_madd<F>(acc, a.x, a.y); }
  static NOVA_HD void add(pt& acc, const pt& o) { xyzz_add<F>(acc, o); }
  static NOVA_HD pt mul_small(const pt& p, uint32_t k) { return xyzz_mul_small<F>(p, k); }
  static NOVA_HD aff to_affine(const pt& q) {
    fe_t iz3 = fe_inv<F>(q.zzz);
    fe_t iz2 = fe_mul<F>(fe_sqr<F>(q.zz), fe_sqr<F>(iz3));
    aff a;
    a.x = fe_mul<F>(q.x, iz2);
    a.y = fe_mul<F>(q.y, iz3);
    return a;
  }
  static NOVA_HD aff from_std_affine(const fe_t& x, const fe_t& y) {
    aff a;
    a.x = x;
    a.y = y;
    return a;
  }
  static NOVA_HD void to_jacobian_std(const pt& p, fe_t& X, fe_t& Y, fe_t& Z) {
    xyzz_to_jacobian<F>(p, X, Y, Z);
  }
  static NOVA_HD pt from_jacobian_std(const fe_t& X, const fe_t& Y, const fe_t& Z) {
    if (fe_is_zero(Z)) return identity();
    pt p;
    p.x = X;
    p.y = Y;
    p.zz = fe_sqr<F>(Z);
    p.zzz = fe_mul<F>(p.zz, Z);
    return p;
  }
#if defined(__CUDACC__)
  static NOVA_D aff load_table(const void* tables, size_t idx) { return affine_load(tables, idx); }
  static NOVA_D void store_table(void* tables, size_t idx, const aff& a) {
    fe_store(tables, 2 * idx, a.x);
    fe_store(tables, 2 * idx + 1, a.y);
  }
  static NOVA_D void store_table_identity(void* tables, size_t idx) {
    fe_store(tables, 2 * idx, fe_zero<F>());
    fe_store(tables, 2 * idx + 1, fe_zero<F>());
  }
  static NOVA_D pt shfl_down(const pt& p, int d, int width) {
    pt r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], d, width);
      r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], d, width);
      r.zz.l[i] = __shfl_down_sync(0xffffffffu, p.zz.l[i], d, width);
      r.zzz.l[i] = __shfl_down_sync(0xffffffffu, p.zzz.l[i], d, width);
    }
    return r;
  }
  static NOVA_D pt load(const void* base, size_t idx) { return xyzz_load(base, idx); }
  static NOVA_D void store(void* base, size_t idx, const pt& p) { xyzz_store(base, idx, p); }
#endif
};

// Default: pa32.  Measured on B200 at 2^20 (profiles/r01c_arith29_experiment.md): k_accumulate
// 2.78 ms with pa32 (IMAD pipe 85 % busy, .X half-rate bound) vs 4.06 ms with pa29 (issue-bound:
// ~4400 instructions per mixed add at 12 warps/SM).  pa29 stays buildable (-DNOVA_MSM_ARITH29),
// host-tested and GPU-parity-tested, as the starting point for a lower-overhead carry-free variant.
#if defined(NOVA_MSM_ARITH29)
template <class F>
using msm_arith = pa29<F>;
#else
template <class F>
using msm_arith = pa32<F>;
#endif

}  // namespace nova
