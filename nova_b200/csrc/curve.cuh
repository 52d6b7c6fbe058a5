// Short-Weierstrass (a = 0) group arithmetic in extended-Jacobian XYZZ coordinates.
//
// Device restatement of the reference's bucket arithmetic, src/provider/msm.rs:38-183:
//   BucketXYZZ {x,y,zz,zzz}, identity <=> zz == 0            (msm.rs:38-62)
//   double_in_place   dbl-2008-s-1, a = 0                     (msm.rs:65-88)
//   add_assign_bucket add-2008-s  + P==Q / P==-Q cases        (msm.rs:91-123)
//   bucket_add_affine madd-2008-s + identity/equal cases      (msm.rs:126-165)
// The curve constant b never appears (a = 0 and additions do not use b), so one template
// over the base field serves BN254 G1, Grumpkin, Pallas and Vesta (msm.rs:35-36).
#pragma once
#include "field.cuh"

namespace nova {

struct xyzz_t {
  fe_t x, y, zz, zzz;
};

struct affine_t {
  fe_t x, y;  // identity <=> x == 0 && y == 0 (halo2curves affine identity encoding)
};

template <class F>
NOVA_HD xyzz_t xyzz_identity() {
  xyzz_t r;
  r.x = fe_one<F>();
  r.y = fe_one<F>();
  r.zz = fe_zero<F>();
  r.zzz = fe_zero<F>();
  return r;
}

NOVA_HD bool xyzz_is_identity(const xyzz_t& p) { return fe_is_zero(p.zz); }
NOVA_HD bool affine_is_identity(const affine_t& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }

// y^2 == x^3 + b, or the identity encoding (0,0): halo2curves `is_on_curve`, the check
// CommitmentKey::new applies to every base (src/provider/hyperkzg.rs:113-119).  The four curves
// have small integer b (3, -17, 5, 5: bn256_grumpkin.rs:35-41,80-86; pasta.rs:33-47), passed as a
// signed integer and converted here.
template <class F>
NOVA_HD fe_t fe_from_small_int(int v) {
  fe_t a = fe_zero<F>();
  a.l[0] = (uint32_t)(v < 0 ? -v : v);
  a = fe_to_mont<F>(a);
  return v < 0 ? fe_neg<F>(a) : a;
}
template <class F>
NOVA_HD bool affine_on_curve(const affine_t& p, const fe_t& b_mont) {
  if (affine_is_identity(p)) return true;
  fe_t lhs = fe_sqr<F>(p.y);
  fe_t rhs = fe_add<F>(fe_mul<F>(fe_sqr<F>(p.x), p.x), b_mont);
  return fe_eq(lhs, rhs);
}

// A point as it arrives from a file or another process: both coordinates canonical AND on the curve
// (read_points, provider/ptau.rs:372-392; in-memory halo2curves points are canonical by construction).
template <class F>
NOVA_HD bool affine_valid_raw(const affine_t& p, const fe_t& b_mont) {
  return fe_is_canonical<F>(p.x) && fe_is_canonical<F>(p.y) && affine_on_curve<F>(p, b_mont);
}

// dbl-2008-s-1 (a = 0): 2M + 5S.  Precondition: not identity, y != 0 is NOT required
// (y == 0 gives zz3 = 0, i.e. the identity, which is the correct answer for a 2-torsion point).
template <class F>
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void xyzz_dbl(xyzz_t& p) {
  if (xyzz_is_identity(p)) return;
  fe_t u = fe_dbl<F>(p.y);
  fe_t v = fe_sqr<F>(u);
  fe_t w = fe_mul<F>(u, v);
  fe_t s = fe_mul<F>(p.x, v);
  fe_t xx = fe_sqr<F>(p.x);
  fe_t m = fe_add<F>(fe_dbl<F>(xx), xx);
  fe_t x3 = fe_sub<F>(fe_sqr<F>(m), fe_dbl<F>(s));
  fe_t y3 = fe_sub<F>(fe_mul<F>(m, fe_sub<F>(s, x3)), fe_mul<F>(w, p.y));
  p.x = x3;
  p.y = y3;
  p.zz = fe_mul<F>(v, p.zz);
  p.zzz = fe_mul<F>(w, p.zzz);
}

// acc += (px, py) with (px,py) a non-identity affine point.  madd-2008-s: 7M + 2S.
template <class F>
NOVA_HD void xyzz_madd(xyzz_t& acc, const fe_t& px, const fe_t& py) {
  if (xyzz_is_identity(acc)) {
    acc.x = px;
    acc.y = py;
    acc.zz = fe_one<F>();
    acc.zzz = fe_one<F>();
    return;
  }
  fe_t u2 = fe_mul<F>(px, acc.zz);
  fe_t s2 = fe_mul<F>(py, acc.zzz);
  if (fe_eq(acc.x, u2)) {
    if (fe_eq(acc.y, s2))
      xyzz_dbl<F>(acc);
    else
      acc = xyzz_identity<F>();
    return;
  }
  fe_t p = fe_sub<F>(u2, acc.x);
  fe_t r = fe_sub<F>(s2, acc.y);
  fe_t pp = fe_sqr<F>(p);
  fe_t ppp = fe_mul<F>(p, pp);
  fe_t q = fe_mul<F>(acc.x, pp);
  fe_t x3 = fe_sub<F>(fe_sub<F>(fe_sqr<F>(r), ppp), fe_dbl<F>(q));
#if !defined(NOVA_MADD_SPLIT_Y3)
  // y3 = r (q - x3) + (-y1) ppp as ONE sum-of-products reduction (fe_mul2_add): 192 instead of
  // 256 wide products, -5 % of the mixed addition.  Measured on B200 (profiles/r02a_variants.md): k_accumulate
  // 2.627 -> 2.486 ms at 2^20, 9.08 -> 8.60 ms at 2^22; the two-product form stays buildable with
  // -DNOVA_MADD_SPLIT_Y3 (make variant VARIANT=split VFLAGS=-DNOVA_MADD_SPLIT_Y3).
  fe_t y3 = fe_mul2_add<F>(r, fe_sub<F>(q, x3), fe_neg<F>(acc.y), ppp);
#else
  fe_t y3 = fe_sub<F>(fe_mul<F>(r, fe_sub<F>(q, x3)), fe_mul<F>(acc.y, ppp));
#endif
  acc.x = x3;
  acc.y = y3;
  acc.zz = fe_mul<F>(acc.zz, pp);
  acc.zzz = fe_mul<F>(acc.zzz, ppp);
}

// acc += other (both XYZZ).  add-2008-s: 12M + 2S.
template <class F>
NOVA_HD void xyzz_add(xyzz_t& acc, const xyzz_t& o) {
  if (xyzz_is_identity(o)) return;
  if (xyzz_is_identity(acc)) {
    acc = o;
    return;
  }
  fe_t u1 = fe_mul<F>(acc.x, o.zz);
  fe_t u2 = fe_mul<F>(o.x, acc.zz);
  fe_t s1 = fe_mul<F>(acc.y, o.zzz);
  fe_t s2 = fe_mul<F>(o.y, acc.zzz);
  if (fe_eq(u1, u2)) {
    if (fe_eq(s1, s2))
      xyzz_dbl<F>(acc);
    else
      acc = xyzz_identity<F>();
    return;
  }
  fe_t p = fe_sub<F>(u2, u1);
  fe_t r = fe_sub<F>(s2, s1);
  fe_t pp = fe_sqr<F>(p);
  fe_t ppp = fe_mul<F>(p, pp);
  fe_t q = fe_mul<F>(u1, pp);
  fe_t x3 = fe_sub<F>(fe_sub<F>(fe_sqr<F>(r), ppp), fe_dbl<F>(q));
  fe_t y3 = fe_sub<F>(fe_mul<F>(r, fe_sub<F>(q, x3)), fe_mul<F>(s1, ppp));
  acc.x = x3;
  acc.y = y3;
  acc.zz = fe_mul<F>(fe_mul<F>(acc.zz, o.zz), pp);
  acc.zzz = fe_mul<F>(fe_mul<F>(acc.zzz, o.zzz), ppp);
}

template <class F>
NOVA_HD void xyzz_neg(xyzz_t& p) {
  p.y = fe_neg<F>(p.y);
}

// XYZZ -> Jacobian (X', Y', Z') without inversion: Z' = ZZ*ZZZ... we need Z'^2 ~ ZZ and
// Z'^3 ~ ZZZ.  With x = X/ZZ, y = Y/ZZZ and any Z' != 0:  X' = x Z'^2, Y' = y Z'^3.
// Choosing Z' = ZZZ/ZZ is not inversion-free; choosing Z' = ZZ*ZZZ gives
//   X' = X * ZZ * ZZZ^2 ,  Y' = Y * ZZ^3 * ZZZ^2 ,  Z' = ZZ * ZZZ.
// Identity -> (0, 1, 0)-style encoding with Z' = 0 (halo2curves projective identity has z = 0).
template <class F>
NOVA_HD void xyzz_to_jacobian(const xyzz_t& p, fe_t& X, fe_t& Y, fe_t& Z) {
  if (xyzz_is_identity(p)) {
    X = fe_zero<F>();
    Y = fe_one<F>();
    Z = fe_zero<F>();
    return;
  }
  fe_t zzz2 = fe_sqr<F>(p.zzz);
  fe_t zz_zzz2 = fe_mul<F>(p.zz, zzz2);           // ZZ * ZZZ^2
  X = fe_mul<F>(p.x, zz_zzz2);
  fe_t zz2 = fe_sqr<F>(p.zz);
  Y = fe_mul<F>(p.y, fe_mul<F>(zz2, zz_zzz2));    // Y * ZZ^3 * ZZZ^2
  Z = fe_mul<F>(p.zz, p.zzz);
}

// [k] P for a small non-negative integer k (double-and-add on XYZZ), used by the bucket
// reduction to weight a chunk's plain sum by the chunk's base index.
template <class F>
NOVA_HD xyzz_t xyzz_mul_small(const xyzz_t& p, uint32_t k) {
  xyzz_t acc = xyzz_identity<F>();
  if (k == 0 || xyzz_is_identity(p)) return acc;
  int top = 31;
  while (!((k >> top) & 1)) top--;
  for (int i = top; i >= 0; i--) {
    xyzz_dbl<F>(acc);
    if ((k >> i) & 1) xyzz_add<F>(acc, p);
  }
  return acc;
}

#if defined(__CUDACC__)
NOVA_D affine_t affine_load(const void* base, size_t idx) {
  affine_t a;
  a.x = fe_load(base, 2 * idx);
  a.y = fe_load(base, 2 * idx + 1);
  return a;
}
NOVA_D xyzz_t xyzz_load(const void* base, size_t idx) {
  xyzz_t p;
  p.x = fe_load_rw(base, 4 * idx);
  p.y = fe_load_rw(base, 4 * idx + 1);
  p.zz = fe_load_rw(base, 4 * idx + 2);
  p.zzz = fe_load_rw(base, 4 * idx + 3);
  return p;
}
NOVA_D void xyzz_store(void* base, size_t idx, const xyzz_t& p) {
  fe_store(base, 4 * idx, p.x);
  fe_store(base, 4 * idx + 1, p.y);
  fe_store(base, 4 * idx + 2, p.zz);
  fe_store(base, 4 * idx + 3, p.zzz);
}
#endif

}  // namespace nova
