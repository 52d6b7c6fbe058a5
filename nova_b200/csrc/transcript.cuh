// Device-side Fiat-Shamir for the sum-check round loop.
//
// Restates, for sm_100a, the O(1) host algebra that sits between two O(N) kernels of a sum-check
// round in the reference, so that a whole round loop can be enqueued without a host round trip:
//   * Keccak256Transcript::{absorb, squeeze}      src/provider/keccak.rs:98-160 (non-evm variant):
//       squeeze hashes  buffered absorbs || "NoDS" || round_le64 || state[64] || label || {0,1}
//       twice (prefix byte 0 -> low half, 1 -> high half), the 64 bytes become the new state and
//       the challenge is from_uniform (64-byte little-endian integer mod p)
//   * UniPoly::{from_evals_deg2, from_evals_deg3, evaluate, compress} and its transcript bytes
//                                                 src/spartan/polys/univariate.rs:89-154, 177-190
//   * the claim derivation / bound of EqSumCheckInstance   src/spartan/sumcheck.rs:680-747, 1226-1231
//     with  q*t(1) = (claim - s(0)) / tau  -- the same field element as the reference's
//     (claim - s(0)) / (tau q) * q, but the divisor no longer depends on the challenges, so all
//     inverses are known before the loop starts.  tau = 0 rounds take the reference's third-sum
//     fall-back (sumcheck.rs:696-698, 1086-1213): the caller supplies t(-1).
//
// Everything is plain integer code (`NOVA_HD`): the CPU unit test compiles this header for the host
// (tests/hostcheck) and drives complete proofs through it, compared with the CPU restatement.
#pragma once
#include "field.cuh"

namespace nova {

// The round kernels are one warp running ~25 field products ONCE: with the multiplier inlined at every site they were
// 138 KB of straight-line code, fetched cold from L2 (instruction-cache misses dominated the kernel).  A called
// multiplier keeps the body resident after its first use.  -DNOVA_ROUND_INLINE_MUL restores the inlined form (A/B).
#if defined(__CUDACC__) && !defined(NOVA_ROUND_INLINE_MUL)
template <class F>
__host__ __device__ __noinline__ fe_t fe_mulx(const fe_t& a, const fe_t& b) {
  return fe_mul<F>(a, b);
}
#else
template <class F>
NOVA_HD fe_t fe_mulx(const fe_t& a, const fe_t& b) {
  return fe_mul<F>(a, b);
}
#endif
#if !defined(SCB_STAMP)
#define SCB_STAMP(i)  // tools/roundbench.cu: clock64() stamps at the section boundaries of the round kernels
#endif

// ---------------------------------------------------------------------------------------------
// Keccak-f[1600] and Keccak-256 (the original padding 0x01 .. 0x80, rate 136 -- sha3 crate
// `Keccak256`, keccak.rs:9)
// ---------------------------------------------------------------------------------------------
NOVA_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

NOVA_HD void keccak_f1600(uint64_t (&a)[25]) {
  constexpr uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  constexpr int ROTC[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
  constexpr int PILN[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#pragma unroll
  for (int round = 0; round < 24; round++) {
    uint64_t c[5];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      uint64_t d = c[(x + 4) % 5] ^ rotl64(c[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
    }
    uint64_t t = a[1];
#pragma unroll
    for (int i = 0; i < 24; i++) {
      uint64_t tmp = a[PILN[i]];
      a[PILN[i]] = rotl64(t, ROTC[i]);
      t = tmp;
    }
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
      uint64_t r0 = a[y], r1 = a[y + 1], r2 = a[y + 2], r3 = a[y + 3], r4 = a[y + 4];
      a[y] = r0 ^ (~r1 & r2);
      a[y + 1] = r1 ^ (~r2 & r3);
      a[y + 2] = r2 ^ (~r3 & r4);
      a[y + 3] = r3 ^ (~r4 & r0);
      a[y + 4] = r4 ^ (~r0 & r1);
    }
    a[0] ^= RC[round];
  }
}

// A message assembled as little-endian 64-bit words (zero beyond `len`), so that the absorb loop
// reads whole lanes with static register indices.
constexpr int KECCAK_RATE_WORDS = 17;  // 136 bytes
constexpr int MSG_MAX_BLOCKS = 16;
constexpr int MSG_WORDS = KECCAK_RATE_WORDS * MSG_MAX_BLOCKS;
constexpr uint32_t MSG_MAX_BYTES = MSG_WORDS * 8 - 1;  // one byte is left for the padding
struct msg_buf {
  uint64_t w[MSG_WORDS];
  uint32_t len;
};
NOVA_HD void msg_reset(msg_buf& m) {
  for (int i = 0; i < MSG_WORDS; i++) m.w[i] = 0;
  m.len = 0;
}
NOVA_HD void msg_put(msg_buf& m, uint8_t b) {
  m.w[m.len >> 3] |= (uint64_t)b << (8 * (m.len & 7));
  m.len++;
}
NOVA_HD void msg_put_bytes(msg_buf& m, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) msg_put(m, p[i]);
}
// four bytes at once (little-endian), at any byte position: one or two word updates instead of four
NOVA_HD void msg_put_u32le(msg_buf& m, uint32_t v) {
  const uint32_t w = m.len >> 3, sh = 8 * (m.len & 7);
  m.w[w] |= (uint64_t)v << sh;
  if (sh > 32) m.w[w + 1] |= (uint64_t)v >> (64 - sh);
  m.len += 4;
}
NOVA_HD void msg_put_u64le(msg_buf& m, uint64_t v) {
  msg_put_u32le(m, (uint32_t)v);
  msg_put_u32le(m, (uint32_t)(v >> 32));
}
// zeroes the words a message of `total_len` bytes (plus its padding) occupies -- all the hash will read
NOVA_HD void msg_reset_for(msg_buf& m, uint32_t total_len) {
  const uint32_t nwords = (total_len / (8 * KECCAK_RATE_WORDS) + 1) * KECCAK_RATE_WORDS;
  for (uint32_t i = 0; i < nwords; i++) m.w[i] = 0;
  m.len = 0;
}

// Keccak-256 of the message with byte `flip_pos` XORed with `flip` (the two squeeze hashes differ
// in their last byte only, so both read the same buffer); out = the 32-byte digest as 4 LE words.
NOVA_HD void keccak256_msg(const msg_buf& m, uint32_t flip_pos, uint8_t flip, uint64_t (&out)[4]) {
  uint64_t a[25];
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = 0;
  const uint32_t nblocks = m.len / (8 * KECCAK_RATE_WORDS) + 1;
  const uint32_t pad_w = m.len >> 3, last_w = nblocks * KECCAK_RATE_WORDS - 1, flip_w = flip_pos >> 3;
  for (uint32_t b = 0; b < nblocks; b++) {
#pragma unroll
    for (int j = 0; j < KECCAK_RATE_WORDS; j++) {
      uint32_t wi = b * KECCAK_RATE_WORDS + j;
      uint64_t v = m.w[wi];
      if (wi == pad_w) v ^= (uint64_t)0x01 << (8 * (m.len & 7));
      if (wi == last_w) v ^= (uint64_t)0x80 << 56;
      if (wi == flip_w) v ^= (uint64_t)flip << (8 * (flip_pos & 7));
      a[j] ^= v;
    }
    keccak_f1600(a);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = a[i];
}

#if defined(__CUDACC__) || defined(NOVA_SIMT_HOST)
// ---------------------------------------------------------------------------------------------
// Keccak-f[1600] spread over one warp: lane i < 25 holds state word i (= x + 5 y); every step of a round is one or two
// 64-bit shuffles instead of a 25-word loop on a single lane, so a permutation is ~24 x 9 dependent shuffles (~3 us)
// rather than ~6 k dependent instructions on one lane (~15 us) -- the one-lane hash was half of a sum-check round
// (profiles/r02a).  All 32 lanes must call; lanes 25..31 mirror lane 0's pattern and hold garbage.
// ---------------------------------------------------------------------------------------------
NOVA_D uint64_t warp_get64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
NOVA_D uint64_t rotl64v(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

NOVA_D void keccak_f1600_warp(uint64_t& a) {
  constexpr uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  constexpr int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  const int lane = (int)(threadIdx.x & 31u);
  const int l = lane < 25 ? lane : 0;
  const int x = l % 5, y = l / 5;
  const int row = 5 * y;
  const int up1 = (l + 5) % 25, up2 = (l + 10) % 25, up3 = (l + 15) % 25, up4 = (l + 20) % 25;  // same column
  const int col_m1 = row + (x + 4) % 5, col_p1 = row + (x + 1) % 5, col_p2 = row + (x + 2) % 5;
  // rho + pi as a gather: B[X][Y] = rot(A[x][y]) with X = y, Y = 2x + 3y  =>  x = X + 3Y, y = X (mod 5)
  const int src_pi = ((x + 3 * y) % 5) + 5 * x;
  int rot = 0;
#pragma unroll
  for (int k = 0; k < 25; k++)
    if (k == src_pi) rot = ROT[k];
  for (int round = 0; round < 24; round++) {
    uint64_t c = a ^ warp_get64(a, up1) ^ warp_get64(a, up2) ^ warp_get64(a, up3) ^ warp_get64(a, up4);
    a ^= warp_get64(c, col_m1) ^ rotl64v(warp_get64(c, col_p1), 1);
    uint64_t b = rotl64v(warp_get64(a, src_pi), rot);
    a = b ^ (~warp_get64(b, col_p1) & warp_get64(b, col_p2));
    if (lane == 0) a ^= RC[round];
  }
}

// Two states in lockstep: the shuffles of both permutations are in flight together, so the pair costs about what one
// costs (the warp is latency-bound).
NOVA_D void keccak_f1600_warp2(uint64_t& a0, uint64_t& a1) {
  constexpr uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  constexpr int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  const int lane = (int)(threadIdx.x & 31u);
  const int l = lane < 25 ? lane : 0;
  const int x = l % 5, y = l / 5;
  const int row = 5 * y;
  const int up1 = (l + 5) % 25, up2 = (l + 10) % 25, up3 = (l + 15) % 25, up4 = (l + 20) % 25;
  const int col_m1 = row + (x + 4) % 5, col_p1 = row + (x + 1) % 5, col_p2 = row + (x + 2) % 5;
  const int src_pi = ((x + 3 * y) % 5) + 5 * x;
  int rot = 0;
#pragma unroll
  for (int k = 0; k < 25; k++)
    if (k == src_pi) rot = ROT[k];
  for (int round = 0; round < 24; round++) {
    uint64_t c0 = a0 ^ warp_get64(a0, up1) ^ warp_get64(a0, up2) ^ warp_get64(a0, up3) ^ warp_get64(a0, up4);
    uint64_t c1 = a1 ^ warp_get64(a1, up1) ^ warp_get64(a1, up2) ^ warp_get64(a1, up3) ^ warp_get64(a1, up4);
    a0 ^= warp_get64(c0, col_m1) ^ rotl64v(warp_get64(c0, col_p1), 1);
    a1 ^= warp_get64(c1, col_m1) ^ rotl64v(warp_get64(c1, col_p1), 1);
    uint64_t b0 = rotl64v(warp_get64(a0, src_pi), rot);
    uint64_t b1 = rotl64v(warp_get64(a1, src_pi), rot);
    a0 = b0 ^ (~warp_get64(b0, col_p1) & warp_get64(b0, col_p2));
    a1 = b1 ^ (~warp_get64(b1, col_p1) & warp_get64(b1, col_p2));
    if (lane == 0) {
      a0 ^= RC[round];
      a1 ^= RC[round];
    }
  }
}

// Both squeeze hashes of a transcript (keccak.rs:131-160: the same message with the last byte 0 and 1) by one warp: the
// blocks before the one holding that byte are absorbed once, from there on the two states run in lockstep.
// out0 / out1: the digests for flip = 0 / 1.
NOVA_D void keccak256_msg_warp_pair(const msg_buf& m, uint32_t flip_pos, uint64_t (&out0)[4], uint64_t (&out1)[4]) {
  const int lane = (int)(threadIdx.x & 31u);
  uint64_t a0 = 0, a1 = 0;
  const uint32_t nblocks = m.len / (8 * KECCAK_RATE_WORDS) + 1;
  const uint32_t pad_w = m.len >> 3, last_w = nblocks * KECCAK_RATE_WORDS - 1, flip_w = flip_pos >> 3;
  const uint32_t fork = flip_w / KECCAK_RATE_WORDS;  // first block in which the two messages differ
  for (uint32_t b = 0; b < nblocks; b++) {
    if (b == fork) a1 = a0;
    if (lane < KECCAK_RATE_WORDS) {
      uint32_t wi = b * KECCAK_RATE_WORDS + lane;
      uint64_t v = m.w[wi];
      if (wi == pad_w) v ^= (uint64_t)0x01 << (8 * (m.len & 7));
      if (wi == last_w) v ^= (uint64_t)0x80 << 56;
      a0 ^= v;
      if (b >= fork) {
        if (wi == flip_w) v ^= (uint64_t)1 << (8 * (flip_pos & 7));
        a1 ^= v;
      }
    }
    if (b < fork) keccak_f1600_warp(a0);
    else keccak_f1600_warp2(a0, a1);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    out0[i] = warp_get64(a0, i);
    out1[i] = warp_get64(a1, i);
  }
}

// Keccak-256 of `m` (same padding / flip convention as keccak256_msg) computed by the whole warp; every lane returns
// the four digest words.
NOVA_D void keccak256_msg_warp(const msg_buf& m, uint32_t flip_pos, uint8_t flip, uint64_t (&out)[4]) {
  const int lane = (int)(threadIdx.x & 31u);
  uint64_t a = 0;
  const uint32_t nblocks = m.len / (8 * KECCAK_RATE_WORDS) + 1;
  const uint32_t pad_w = m.len >> 3, last_w = nblocks * KECCAK_RATE_WORDS - 1, flip_w = flip_pos >> 3;
  for (uint32_t b = 0; b < nblocks; b++) {
    if (lane < KECCAK_RATE_WORDS) {
      uint32_t wi = b * KECCAK_RATE_WORDS + lane;
      uint64_t v = m.w[wi];
      if (wi == pad_w) v ^= (uint64_t)0x01 << (8 * (m.len & 7));
      if (wi == last_w) v ^= (uint64_t)0x80 << 56;
      if (wi == flip_w) v ^= (uint64_t)flip << (8 * (flip_pos & 7));
      a ^= v;
    }
    keccak_f1600_warp(a);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = warp_get64(a, i);
}
#endif

// ---------------------------------------------------------------------------------------------
// field helpers
// ---------------------------------------------------------------------------------------------
// from_uniform (traits.rs:315-319 -> halo2curves from_uniform_bytes): the 64-byte little-endian
// integer lo + 2^256 hi mod p, returned in Montgomery form.
template <class F>
NOVA_HD fe_t fe_from_uniform(const uint64_t (&w)[8]) {
  fe_t lo, hi, r2;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    lo.l[2 * i] = (uint32_t)w[i];
    lo.l[2 * i + 1] = (uint32_t)(w[i] >> 32);
    hi.l[2 * i] = (uint32_t)w[4 + i];
    hi.l[2 * i + 1] = (uint32_t)(w[4 + i] >> 32);
  }
  // 2^256 / p < 6 for all four moduli: bring both halves into [0, p) first
  for (int k = 0; k < 6; k++) {
    fe_reduce_once<F>(lo.l);
    fe_reduce_once<F>(hi.l);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) r2.l[i] = F::r2(i);
  // lo R + (hi R) R  =  (lo + 2^256 hi) R   (mod p)
  return fe_add<F>(fe_mulx<F>(lo, r2), fe_mulx<F>(fe_mulx<F>(hi, r2), r2));
}

// x / 2 on any residue representation (Montgomery included): (x + (x odd ? p : 0)) >> 1
template <class F>
NOVA_HD fe_t fe_half(const fe_t& a) {
  uint64_t carry = 0;
  uint32_t t[9];
  const uint32_t odd = a.l[0] & 1u;
  for (int i = 0; i < 8; i++) {
    uint64_t s = (uint64_t)a.l[i] + (odd ? F::p(i) : 0u) + carry;
    t[i] = (uint32_t)s;
    carry = s >> 32;
  }
  t[8] = (uint32_t)carry;
  fe_t r;
  for (int i = 0; i < 8; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
  return r;
}

// ---------------------------------------------------------------------------------------------
// one sum-check round between the reduction kernel and the bind kernels
// ---------------------------------------------------------------------------------------------
// Mirrors the layout documented in include/nova_b200.h (b200_sc_state, 144 bytes).
struct sc_state {
  fe_t claim;          // running claim e (Montgomery)
  fe_t q;              // EqSumCheckInstance::eval_eq_left (Montgomery); 1 for the plain kinds
  uint64_t round;      // Keccak256Transcript::round
  uint8_t tstate[64];  // Keccak256Transcript::state
  uint64_t rounds_done;
};

enum sc_round_kind {
  SC_ROUND_QUAD_PROD = 0,     // res = [eval_point_0, bound_coeff]            sumcheck.rs:213-220
  SC_ROUND_CUBIC3_EQ = 1,     // res = [t(0), t(inf)], tau != 0               sumcheck.rs:459-470, 680-715
  SC_ROUND_CUBIC3_EQ_M1 = 2,  // res = [t(0), t(inf), t(-1)], tau == 0         sumcheck.rs:696-698
};

struct sc_round_poly {
  fe_t c[4];   // coefficients, constant term first (Montgomery)
  int deg;     // 2 or 3
  fe_t tau;    // eq kinds: this round's tau
};

// Builds the round polynomial from the reduction results and the running state.
template <class F>
NOVA_HD void sc_round_build(int kind, const sc_state& st, const fe_t* res, const fe_t& tau,
                            const fe_t& tau_inv, sc_round_poly& out) {
  const fe_t e0_in = res[0];
  if (kind == SC_ROUND_QUAD_PROD) {
    // evals [e0, claim - e0, bc] -> c + b x + a x^2 with c = e0, a = bc, b = e1 - a - c
    fe_t e1 = fe_sub<F>(st.claim, e0_in);
    out.deg = 2;
    out.c[0] = e0_in;
    out.c[2] = res[1];
    out.c[1] = fe_sub<F>(fe_sub<F>(e1, res[1]), e0_in);
    out.c[3] = fe_zero<F>();
    out.tau = fe_zero<F>();
    return;
  }
  const fe_t one = fe_one<F>();
  const fe_t two_tau = fe_dbl<F>(tau);
  const fe_t e0c = fe_sub<F>(one, tau);                              // eq(tau, 0)
  const fe_t slope = fe_sub<F>(two_tau, one);                        // 2 tau - 1
  const fe_t em1c = fe_sub<F>(fe_dbl<F>(one), fe_add<F>(two_tau, tau));  // eq(tau, -1) = 2 - 3 tau
  const fe_t qt0 = fe_mulx<F>(st.q, res[0]);
  const fe_t qtinf = fe_mulx<F>(st.q, res[1]);
  const fe_t s0 = fe_mulx<F>(e0c, qt0);
  const fe_t lead = fe_mulx<F>(slope, qtinf);
  fe_t em1;
  if (kind == SC_ROUND_CUBIC3_EQ_M1) {
    em1 = fe_mulx<F>(em1c, fe_mulx<F>(st.q, res[2]));
  } else {
    // q t(1) = (claim - s0) / tau ;  t(-1) = 2 t(inf) + 2 t(0) - t(1)
    fe_t qt1 = fe_mulx<F>(fe_sub<F>(st.claim, s0), tau_inv);
    em1 = fe_mulx<F>(em1c, fe_sub<F>(fe_dbl<F>(fe_add<F>(qtinf, qt0)), qt1));
  }
  // evals [d, a+b+c+d, a, s(-1)] -> d + c x + b x^2 + a x^3
  fe_t e1 = fe_sub<F>(st.claim, s0);
  fe_t b = fe_sub<F>(fe_half<F>(fe_add<F>(e1, em1)), s0);
  out.deg = 3;
  out.c[0] = s0;
  out.c[3] = lead;
  out.c[2] = b;
  out.c[1] = fe_sub<F>(fe_sub<F>(fe_sub<F>(e1, lead), s0), b);
  out.tau = tau;
}

// Compressed coefficients (all but the linear term) as canonical little-endian bytes: what the
// proof stores and what `absorb(b"p", &poly)` hashes (univariate.rs:177-190).  Returns their count.
template <class F>
NOVA_HD int sc_round_compressed(const sc_round_poly& p, fe_t (&canon)[3]) {
  canon[0] = fe_from_mont<F>(p.c[0]);
  canon[1] = fe_from_mont<F>(p.c[2]);
  if (p.deg == 3) canon[2] = fe_from_mont<F>(p.c[3]);
  return p.deg;  // deg 2 -> 2 coefficients, deg 3 -> 3
}

// Assembles the squeeze message; returns the position of the trailing prefix byte (written as 0).
// `pending` = bytes absorbed since the last squeeze (first round only), at most
// MSG_MAX_BYTES - 300.
NOVA_HD uint32_t sc_round_message(msg_buf& m, const uint8_t* pending, uint32_t pending_len,
                                  uint8_t absorb_label, const fe_t* canon, int ncoef,
                                  const sc_state& st, uint8_t squeeze_label) {
  msg_reset_for(m, pending_len + 1 + 32 * (uint32_t)ncoef + 4 + 8 + 64 + 2);
  msg_put_bytes(m, pending, pending_len);
  msg_put(m, absorb_label);
  for (int k = 0; k < ncoef; k++)
    for (int i = 0; i < 8; i++) msg_put_u32le(m, canon[k].l[i]);
  msg_put_u32le(m, 0x53446f4eu);  // "NoDS"
  msg_put_u64le(m, st.round);
  for (int i = 0; i < 64; i += 4)
    msg_put_u32le(m, (uint32_t)st.tstate[i] | ((uint32_t)st.tstate[i + 1] << 8) | ((uint32_t)st.tstate[i + 2] << 16) |
                         ((uint32_t)st.tstate[i + 3] << 24));
  msg_put(m, squeeze_label);
  uint32_t pos = m.len;
  msg_put(m, 0);
  return pos;
}

// Challenge, new claim, eq bound, transcript state.
template <class F>
NOVA_HD fe_t sc_round_finish(int kind, sc_state& st, const sc_round_poly& p, const uint64_t (&digest)[8]) {
  fe_t r = fe_from_uniform<F>(digest);
  // UniPoly::evaluate (Horner form of the same polynomial value)
  fe_t acc = p.c[p.deg];
  for (int k = p.deg - 1; k >= 0; k--) acc = fe_add<F>(fe_mulx<F>(acc, r), p.c[k]);
  st.claim = acc;
  if (kind != SC_ROUND_QUAD_PROD) {
    // eval_eq_left *= 1 - tau - r + 2 r tau            (sumcheck.rs:1226-1231)
    fe_t rt = fe_mulx<F>(r, p.tau);
    fe_t f = fe_add<F>(fe_sub<F>(fe_sub<F>(fe_one<F>(), p.tau), r), fe_dbl<F>(rt));
    st.q = fe_mulx<F>(st.q, f);
  }
  for (int i = 0; i < 8; i++)
    for (int b = 0; b < 8; b++) st.tstate[8 * i + b] = (uint8_t)(digest[i] >> (8 * b));
  st.round += 1;
  st.rounds_done += 1;
  return r;
}

#if defined(__CUDACC__) || defined(NOVA_SIMT_HOST)  // NOVA_SIMT_HOST: tests/hostcheck/simt_host.h
// One warp: the (cheap, redundant) field algebra runs on every lane so that no result has to be broadcast; the two
// squeeze hashes are computed by the whole warp (keccak_f1600_warp).  <<<1, 32>>>.
template <class F>
__global__ void __launch_bounds__(32) k_sc_round(int kind, sc_state* __restrict__ state,
                                                 const void* __restrict__ res,
                                                 const void* __restrict__ tau_ptr,
                                                 const void* __restrict__ tau_inv_ptr,
                                                 const uint8_t* __restrict__ pending, uint32_t pending_len,
                                                 uint8_t absorb_label, uint8_t squeeze_label,
                                                 void* __restrict__ out_poly, void* __restrict__ out_r) {
  __shared__ msg_buf msg;
  __shared__ uint32_t flip_pos_sh;
  const unsigned lane = threadIdx.x;
  sc_state st = *state;
  fe_t r3[3];
  r3[0] = fe_load_rw(res, 0);
  r3[1] = fe_load_rw(res, 1);
  r3[2] = kind == SC_ROUND_CUBIC3_EQ_M1 ? fe_load_rw(res, 2) : fe_zero<F>();
  fe_t tau = fe_zero<F>(), tau_inv = fe_zero<F>();
  if (kind != SC_ROUND_QUAD_PROD) {
    tau = fe_load_rw(tau_ptr, 0);
    if (kind == SC_ROUND_CUBIC3_EQ) tau_inv = fe_load_rw(tau_inv_ptr, 0);
  }
  sc_round_poly poly;
  sc_round_build<F>(kind, st, r3, tau, tau_inv, poly);
  fe_t canon[3];
  const int ncoef = sc_round_compressed<F>(poly, canon);
  if (lane == 0) {
    flip_pos_sh = sc_round_message(msg, pending, pending_len, absorb_label, canon, ncoef, st, squeeze_label);
    for (int k = 0; k < ncoef; k++) fe_store(out_poly, k, canon[k]);
  }
  __syncwarp();
  // the two squeeze hashes (they differ in the last message byte), each computed by the whole warp
  uint64_t digest[8];
  {
    uint64_t d0[4], d1[4];
    keccak256_msg_warp_pair(msg, flip_pos_sh, d0, d1);
    for (int i = 0; i < 4; i++) {
      digest[i] = d0[i];
      digest[4 + i] = d1[i];
    }
  }
  fe_t r = sc_round_finish<F>(kind, st, poly, digest);
  if (lane == 0) {
    *state = st;
    fe_store(out_r, 0, r);
  }
}

// out[i] = 1 / in[i] (0 -> 0), one thread per element: the handful of per-round constants of an eq
// sum-check (1/tau_i), computed once before the round loop.
template <class F>
__global__ void __launch_bounds__(64) k_fe_inv_each(const void* __restrict__ in, size_t n,
                                                    void* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_t a = fe_load_rw(in, i);
  fe_store(out, i, fe_is_zero(a) ? a : fe_inv<F>(a));
}
#endif

}  // namespace nova
