// Quad-cooperative XYZZ arithmetic for the latency-bound tail of the MSM (bucket fix-up and
// reduction).  A lone warp needs ~9 us for one XYZZ addition on B200: its 14 field products are a
// dependent chain of ~1250 cycles each.  The tail kernels have only a few thousand logical
// threads, so four lanes can share one point operation: every lane holds the operands, each
// computes one of the (up to four) independent products of a stage, and the results are exchanged
// inside the quad.  An addition becomes 4 product stages instead of 14 sequential products, a
// doubling 3 instead of 7.
//
// The exchange is abstracted (`Comm`): on the device it is __shfl_sync inside aligned groups of four
// lanes; the CPU unit test (tests/hostcheck) runs four std::threads per quad with a barrier-based
// Comm, so the stage logic is validated without a GPU.  All four lanes of a quad must call every
// function together and end up with identical results.
//
// Formulas: add-2008-s / dbl-2008-s-1 as in curve.cuh (src/provider/msm.rs:65-123).
#pragma once
#include "curve.cuh"

namespace nova {

#if defined(__CUDACC__)
struct quad_comm_dev {
  // Only the quad's own four lanes are named in the shuffle mask, so different quads of a warp may
  // take different branches (one quad adds while its neighbour idles) without deadlock.
  unsigned mask;
  NOVA_D quad_comm_dev() : mask(0xFu << ((threadIdx.x & 31u) & ~3u)) {}
  NOVA_D int lane() const { return (int)(threadIdx.x & 3); }
  // value of `v` held by quad lane `src` (0..3)
  NOVA_D fe_t get(const fe_t& v, int src) const {
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(mask, v.l[i], src, 4);
    return r;
  }
};
#endif

// operand selection by quad lane (data select, no divergent control flow: all four lanes then run
// the SAME fe_mul on their own operands)
NOVA_HD fe_t sel4(int q, const fe_t& a0, const fe_t& a1, const fe_t& a2, const fe_t& a3) {
  fe_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t lo = q & 1 ? a1.l[i] : a0.l[i];
    uint32_t hi = q & 1 ? a3.l[i] : a2.l[i];
    r.l[i] = q & 2 ? hi : lo;
  }
  return r;
}

template <class F, class Comm>
NOVA_HD void coop_dbl(xyzz_t& p, const Comm& cm) {
  if (xyzz_is_identity(p)) return;  // uniform across the quad
  const int q = cm.lane();
  const fe_t z = fe_zero<F>();
  fe_t u = fe_dbl<F>(p.y);
  // stage 1: v = u^2 | xx = x^2
  fe_t a = sel4(q, u, p.x, z, z);
  fe_t r = fe_mul<F>(a, a);
  fe_t v = cm.get(r, 0), xx = cm.get(r, 1);
  fe_t m = fe_add<F>(fe_dbl<F>(xx), xx);
  // stage 2: w = u v | s = x v | mm = m^2 | zz3 = v zz
  r = fe_mul<F>(sel4(q, u, p.x, m, v), sel4(q, v, v, m, p.zz));
  fe_t w = cm.get(r, 0), s = cm.get(r, 1), mm = cm.get(r, 2), zz3 = cm.get(r, 3);
  fe_t x3 = fe_sub<F>(mm, fe_dbl<F>(s));
  // stage 3: t = m (s - x3) | wy = w y | zzz3 = w zzz
  r = fe_mul<F>(sel4(q, m, w, w, z), sel4(q, fe_sub<F>(s, x3), p.y, p.zzz, z));
  fe_t t = cm.get(r, 0), wy = cm.get(r, 1), zzz3 = cm.get(r, 2);
  p.x = x3;
  p.y = fe_sub<F>(t, wy);
  p.zz = zz3;
  p.zzz = zzz3;
}

template <class F, class Comm>
NOVA_HD void coop_add(xyzz_t& acc, const xyzz_t& o, const Comm& cm) {
  if (xyzz_is_identity(o)) return;
  if (xyzz_is_identity(acc)) {
    acc = o;
    return;
  }
  const int q = cm.lane();
  const fe_t z = fe_zero<F>();
  // stage 1: u1 = x1 zz2 | u2 = x2 zz1 | s1 = y1 zzz2 | s2 = y2 zzz1
  fe_t r = fe_mul<F>(sel4(q, acc.x, o.x, acc.y, o.y), sel4(q, o.zz, acc.zz, o.zzz, acc.zzz));
  fe_t u1 = cm.get(r, 0), u2 = cm.get(r, 1), s1 = cm.get(r, 2), s2 = cm.get(r, 3);
  if (fe_eq(u1, u2)) {  // uniform across the quad: same data in every lane
    if (fe_eq(s1, s2))
      coop_dbl<F>(acc, cm);
    else
      acc = xyzz_identity<F>();
    return;
  }
  fe_t P = fe_sub<F>(u2, u1), R = fe_sub<F>(s2, s1);
  // stage 2: pp = P^2 | rr = R^2 | zz12 = zz1 zz2 | zzz12 = zzz1 zzz2
  r = fe_mul<F>(sel4(q, P, R, acc.zz, acc.zzz), sel4(q, P, R, o.zz, o.zzz));
  fe_t pp = cm.get(r, 0), rr = cm.get(r, 1), zz12 = cm.get(r, 2), zzz12 = cm.get(r, 3);
  // stage 3: ppp = P pp | qq = u1 pp | zz3 = zz12 pp
  r = fe_mul<F>(sel4(q, P, u1, zz12, z), sel4(q, pp, pp, pp, z));
  fe_t ppp = cm.get(r, 0), qq = cm.get(r, 1), zz3 = cm.get(r, 2);
  fe_t x3 = fe_sub<F>(fe_sub<F>(rr, ppp), fe_dbl<F>(qq));
  // stage 4: t1 = R (qq - x3) | t2 = s1 ppp | zzz3 = zzz12 ppp
  r = fe_mul<F>(sel4(q, R, s1, zzz12, z), sel4(q, fe_sub<F>(qq, x3), ppp, ppp, z));
  fe_t t1 = cm.get(r, 0), t2 = cm.get(r, 1), zzz3 = cm.get(r, 2);
  acc.x = x3;
  acc.y = fe_sub<F>(t1, t2);
  acc.zz = zz3;
  acc.zzz = zzz3;
}

}  // namespace nova
