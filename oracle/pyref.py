"""Python big-integer oracle for the Nova prover hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product path (nova_b200/) never does.

Everything here is a first-principles restatement (no halo2curves available: the crate
`halo2curves = "0.9.0"`, Cargo.toml:38, is not vendored in /root/reference) of:

  * the four prime fields named at  src/provider/bn256_grumpkin.rs:39-40,84-85 and
    src/provider/pasta.rs:37-38,45-46  (moduli as hex literals there);
  * `from_uniform` = 64-byte little-endian integer mod p  (src/provider/traits.rs:315-319,
    pinned by src/provider/curve_property_tests.rs:40-72);
  * `to_repr` / `to_bytes` = 32-byte little-endian canonical (src/provider/traits.rs:323-327);
  * the short-Weierstrass a=0 group law for BN254 G1, Grumpkin, Pallas, Vesta;
  * the Keccak256 Fiat-Shamir transcript (src/provider/keccak.rs:60-160), used ONLY to pin the
    field code against the golden vectors at src/provider/keccak.rs:241-258.

Parity status: PINNED for field encodings (keccak golden vectors, tests/test_oracle_golden.py),
pinned structurally (MSM == naive sum, src/provider/msm.rs:722-821) for group results; the
reference holds no literal commitment bytes.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------------------
# Fields
# ----------------------------------------------------------------------------------------
BN254_FR = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
BN254_FQ = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
PALLAS_FP = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_FQ = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001

# field ids used across the C ABI (include/nova_b200.h)
FIELD_BN254_FR, FIELD_BN254_FQ, FIELD_PALLAS_FP, FIELD_PALLAS_FQ = 0, 1, 2, 3
FIELD_MODULUS = {
    FIELD_BN254_FR: BN254_FR,
    FIELD_BN254_FQ: BN254_FQ,
    FIELD_PALLAS_FP: PALLAS_FP,
    FIELD_PALLAS_FQ: PALLAS_FQ,
}
FIELD_NAME = {0: "bn254_fr", 1: "bn254_fq", 2: "pallas_fp", 3: "pallas_fq"}

R_BITS = 256
R = 1 << R_BITS


def from_uniform(p: int, b: bytes) -> int:
    """src/provider/traits.rs:315-319 -> halo2curves `from_uniform_bytes`: LE integer mod p."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % p


def to_repr(x: int) -> bytes:
    """32-byte little-endian canonical encoding (src/provider/traits.rs:323-327)."""
    return int(x).to_bytes(32, "little")


def to_mont(p: int, x: int) -> int:
    return (x << R_BITS) % p


def from_mont(p: int, x: int) -> int:
    return (x * pow(R, -1, p)) % p


def mont_bytes(p: int, x: int) -> bytes:
    """In-memory layout of a halo2curves field element: 4 x u64 LE limbs, Montgomery R=2^256."""
    return to_mont(p, x % p).to_bytes(32, "little")


def from_mont_bytes(p: int, b: bytes) -> int:
    return from_mont(p, int.from_bytes(b, "little"))


# ----------------------------------------------------------------------------------------
# Curves (all y^2 = x^3 + b, a = 0)
# ----------------------------------------------------------------------------------------
CURVE_BN254_G1, CURVE_GRUMPKIN, CURVE_PALLAS, CURVE_VESTA = 0, 1, 2, 3


def _sqrt(p: int, a: int) -> int:
    """Tonelli-Shanks."""
    a %= p
    if a == 0:
        return 0
    assert pow(a, (p - 1) // 2, p) == 1, "not a square"
    if p % 4 == 3:
        return pow(a, (p + 1) // 4, p)
    q, s = p - 1, 0
    while q % 2 == 0:
        q //= 2
        s += 1
    z = 2
    while pow(z, (p - 1) // 2, p) != p - 1:
        z += 1
    m, c, t, r = s, pow(z, q, p), pow(a, q, p), pow(a, (q + 1) // 2, p)
    while t != 1:
        i, t2 = 0, t
        while t2 != 1:
            t2 = t2 * t2 % p
            i += 1
        b = pow(c, 1 << (m - i - 1), p)
        m, c, t, r = i, b * b % p, t * b * b % p, r * b % p
    return r


class Curve:
    def __init__(self, cid, name, base_field, scalar_field, b, gen):
        self.id = cid
        self.name = name
        self.base_field = base_field      # field id of coordinates
        self.scalar_field = scalar_field  # field id of scalars
        self.p = FIELD_MODULUS[base_field]
        self.q = FIELD_MODULUS[scalar_field]
        self.b = b % self.p
        self.gen = gen
        assert self.on_curve(gen)

    def on_curve(self, P):
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.b) % self.p == 0

    # affine group law; None is the identity
    def neg(self, P):
        return None if P is None else (P[0], (-P[1]) % self.p)

    def add(self, P, Q):
        p = self.p
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return (x3, (lam * (x1 - x3) - y1) % p)

    def mul(self, k, P):
        k %= self.q
        acc = None
        add = P
        while k:
            if k & 1:
                acc = self.add(acc, add)
            add = self.add(add, add)
            k >>= 1
        return acc

    def msm_naive(self, scalars, bases):
        """Sum_i s_i * P_i  (src/provider/msm.rs:422-429 `msm_simple`, the definition)."""
        acc = None
        for s, P in zip(scalars, bases):
            acc = self.add(acc, self.mul(s, P))
        return acc

    # --- fast Jacobian arithmetic for building large deterministic base sets --------------
    def bases_arith(self, n, k0=0x5EED):
        """P_i = [k0]G + i*G for i in 0..n (same construction as
        src/provider/curve_property_tests.rs:186-194), batch-normalised."""
        p = self.p
        P0 = self.mul(k0, self.gen)
        G = self.gen
        # incremental affine adds with batched inversion in blocks
        out = []
        cur = P0
        # simple: Jacobian accumulate then batch normalise
        jac = []
        X, Y, Z = cur[0], cur[1], 1
        gx, gy = G
        for _ in range(n):
            jac.append((X, Y, Z))
            # mixed add (X,Y,Z) + (gx,gy)
            Z2 = Z * Z % p
            U2 = gx * Z2 % p
            S2 = gy * Z2 * Z % p
            H = (U2 - X) % p
            Rr = (S2 - Y) % p
            if H == 0:
                # doubling / inverse: fall back to affine
                aff = self.add(self._jac_to_aff((X, Y, Z)), G)
                X, Y, Z = (aff[0], aff[1], 1) if aff else (1, 1, 0)
                continue
            H2 = H * H % p
            H3 = H2 * H % p
            X3 = (Rr * Rr - H3 - 2 * X * H2) % p
            Y3 = (Rr * (X * H2 - X3) - Y * H3) % p
            Z3 = Z * H % p
            X, Y, Z = X3, Y3, Z3
        # batch inversion of Z's
        zs = [j[2] for j in jac]
        pref = [1] * (n + 1)
        for i, z in enumerate(zs):
            pref[i + 1] = pref[i] * z % p
        inv = pow(pref[n], -1, p)
        for i in range(n - 1, -1, -1):
            zi = inv * pref[i] % p
            inv = inv * zs[i] % p
            zi2 = zi * zi % p
            out.append((jac[i][0] * zi2 % p, jac[i][1] * zi2 * zi % p))
        out.reverse()
        return out

    def _jac_to_aff(self, J):
        X, Y, Z = J
        if Z % self.p == 0:
            return None
        zi = pow(Z, -1, self.p)
        return (X * zi * zi % self.p, Y * zi * zi * zi % self.p)

    # --- boundary encodings ---------------------------------------------------------------
    def affine_bytes(self, P) -> bytes:
        """halo2curves affine {x,y} 64 B, Montgomery limbs; identity = zero coordinates."""
        if P is None:
            return bytes(64)
        return mont_bytes(self.p, P[0]) + mont_bytes(self.p, P[1])

    def affine_from_bytes(self, b: bytes):
        x = from_mont_bytes(self.p, b[:32])
        y = from_mont_bytes(self.p, b[32:64])
        if x == 0 and y == 0:
            return None
        return (x, y)

    def jacobian_from_bytes(self, b: bytes):
        """{x,y,z} 96 B Jacobian Montgomery -> affine tuple / None."""
        X = from_mont_bytes(self.p, b[:32])
        Y = from_mont_bytes(self.p, b[32:64])
        Z = from_mont_bytes(self.p, b[64:96])
        return self._jac_to_aff((X, Y, Z))


def _grumpkin_gen():
    # y^2 = x^3 - 17 over BN254 Fr, generator x = 1 (halo2curves grumpkin); the root whose
    # low hex digits are ...d823f272c (SURVEY.md §8 "field/curve facts").
    y = _sqrt(BN254_FR, 1 - 17)
    if (y & 0xFFF) != 0x72C:
        y = BN254_FR - y
    assert (y & 0xFFFFFFFFF) == 0xD823F272C
    return (1, y)


CURVES = {
    CURVE_BN254_G1: Curve(CURVE_BN254_G1, "bn254_g1", FIELD_BN254_FQ, FIELD_BN254_FR, 3, (1, 2)),
    CURVE_GRUMPKIN: Curve(CURVE_GRUMPKIN, "grumpkin", FIELD_BN254_FR, FIELD_BN254_FQ, -17, _grumpkin_gen()),
    CURVE_PALLAS: Curve(CURVE_PALLAS, "pallas", FIELD_PALLAS_FP, FIELD_PALLAS_FQ, 5, (PALLAS_FP - 1, 2)),
    CURVE_VESTA: Curve(CURVE_VESTA, "vesta", FIELD_PALLAS_FQ, FIELD_PALLAS_FP, 5, (PALLAS_FQ - 1, 2)),
}

# ----------------------------------------------------------------------------------------
# Keccak-256 (original Keccak padding 0x01, not SHA3's 0x06) + Nova transcript
# ----------------------------------------------------------------------------------------
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [
    [0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14],
]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(A):
    for rc in _RC:
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= rc
    return A


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        blk = msg[off:off + rate]
        for i in range(rate // 8):
            x, y = i % 5, i // 5
            A[x][y] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = _keccak_f(A)
    out = b""
    for i in range(4):
        x, y = i % 5, i // 5
        out += A[x][y].to_bytes(8, "little")
    return out


class Keccak256Transcript:
    """Non-EVM variant of src/provider/keccak.rs:98-160."""

    def __init__(self, p: int, label: bytes):
        self.p = p
        self.round = 0
        self.buf = b""
        self.state = self._updated_state(b"", b"NoTR" + label)

    @staticmethod
    def _updated_state(buf: bytes, inp: bytes) -> bytes:
        # keccak.rs:66-95: H(buf || inp || 0) || H(buf || inp || 1)
        return keccak256(buf + inp + b"\x00") + keccak256(buf + inp + b"\x01")

    def absorb_bytes(self, label: bytes, repr_: bytes):
        self.buf += label + repr_

    def absorb_scalar(self, label: bytes, x: int):
        self.absorb_bytes(label, to_repr(x % self.p))

    def squeeze(self, label: bytes) -> int:
        inp = b"NoDS" + self.round.to_bytes(8, "little") + self.state + label
        out = self._updated_state(self.buf, inp)
        self.round += 1
        self.state = out
        self.buf = b""
        return from_uniform(self.p, out)


# ----------------------------------------------------------------------------------------
# Deterministic PRNG shared by every harness (SplitMix64) -- OUR generator, not halo2curves'
# ----------------------------------------------------------------------------------------
class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & _M64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def bytes(self, n: int) -> bytes:
        out = b""
        while len(out) < n:
            out += self.next().to_bytes(8, "little")
        return out[:n]

    def field(self, p: int) -> int:
        """uniform via the from_uniform rule (64 bytes reduced mod p)."""
        return from_uniform(p, self.bytes(64))


# ----------------------------------------------------------------------------------------
# Sum-check restatement on Python integers (src/spartan/sumcheck.rs, polys/*.rs).  Canonical
# integers mod p throughout; slow, for sizes up to ~2^12.
# ----------------------------------------------------------------------------------------
def eq_evals(p: int, r):
    """EqPolynomial::evals_from_points (polys/eq.rs:54-73)."""
    ev = [0] * (1 << len(r))
    ev[0] = 1
    size = 1
    for rk in reversed(r):
        for i in range(size):
            ev[size + i] = ev[i] * rk % p
            ev[i] = (ev[i] - ev[size + i]) % p
        size *= 2
    return ev


def mle_evaluate(p: int, Z, r):
    """MultilinearPolynomial::evaluate (polys/multilinear.rs:89-127) = <Z, eq(r)>."""
    e = eq_evals(p, r)
    return sum(z * w for z, w in zip(Z, e)) % p


def bind_top(p: int, Z, r):
    """bind_poly_var_top (polys/multilinear.rs:65-84)."""
    h = len(Z) // 2
    return [(Z[i] + r * (Z[i + h] - Z[i])) % p for i in range(h)]


class UniPoly:
    """polys/univariate.rs:89-154, 177-205."""

    def __init__(self, p, coeffs):
        self.p, self.coeffs = p, [c % p for c in coeffs]

    @classmethod
    def from_evals_deg2(cls, p, ev):
        c, abc, a = ev[0], ev[1], ev[2]
        return cls(p, [c, abc - a - c, a])

    @classmethod
    def from_evals_deg3(cls, p, ev):
        d, abcd, a = ev[0], ev[1], ev[2]
        b = ((abcd + ev[3]) * pow(2, -1, p) - d) % p
        c = (abcd - a - d - b) % p
        return cls(p, [d, c, b, a])

    def evaluate(self, r):
        return sum(c * pow(r, i, self.p) for i, c in enumerate(self.coeffs)) % self.p

    def compressed(self):
        return [self.coeffs[0]] + self.coeffs[2:]

    def to_transcript_bytes(self):
        return b"".join(to_repr(c) for c in self.compressed())


class EqSumCheckInstance:
    """spartan/sumcheck.rs:593-1251 (Gruen split-eq + BDDT claim-derived points)."""

    def __init__(self, p, taus):
        self.p = p
        l = len(taus)
        self.init_num_vars = l
        self.first_half = l // 2
        self.second_half = l - self.first_half
        self.round = 1
        self.taus = list(taus)
        self.eval_eq_left = 1

        def compute(ts):  # sumcheck.rs:614-634
            res = [[1]]
            for t in ts:
                prev = res[-1]
                hi = [v * t % p for v in prev]
                lo = [(a - b) % p for a, b in zip(prev, hi)]
                res.append(lo + hi)
            return res

        left = list(reversed(taus[1:self.first_half])) if self.first_half >= 1 else []
        right = list(reversed(taus[self.first_half:]))
        self.poly_eq_left = compute(left)
        self.poly_eq_right = compute(right)
        self.eq_tau_0_a_inf = [((1 - t) % p, (2 * t - 1) % p, (2 - 3 * t) % p) for t in taus]

    # -- which tables a round uses (sumcheck.rs:1233-1251) --
    def tables(self):
        if self.round < self.first_half:
            return (self.poly_eq_left[self.first_half - self.round], self.poly_eq_right[self.second_half],
                    self.second_half)
        return (None, self.poly_eq_right[self.init_num_vars - self.round], 0)

    def factor(self, idx):
        L, R, sh = self.tables()
        if L is None:
            return R[idx]
        return L[idx >> sh] * R[idx & ((1 << sh) - 1)] % self.p

    def derive_deg2(self, t0, tinf, claim):  # sumcheck.rs:680-715
        p = self.p
        e0, slope, em1 = self.eq_tau_0_a_inf[self.round - 1]
        l1p = (e0 + slope) * self.eval_eq_left % p
        if l1p == 0:
            return None
        s0 = e0 * self.eval_eq_left * t0 % p
        t1 = (claim - s0) * pow(l1p, -1, p) % p
        s_lead = slope * self.eval_eq_left * tinf % p
        tm1 = (2 * tinf + 2 * t0 - t1) % p
        return s0, s_lead, em1 * self.eval_eq_left * tm1 % p

    def derive_deg1(self, t0, claim):  # sumcheck.rs:717-747
        p = self.p
        e0, slope, em1 = self.eq_tau_0_a_inf[self.round - 1]
        l1p = (e0 + slope) * self.eval_eq_left % p
        if l1p == 0:
            return None
        s0 = e0 * self.eval_eq_left * t0 % p
        t1 = (claim - s0) * pow(l1p, -1, p) % p
        tm1 = (2 * t0 - t1) % p
        return s0, 0, em1 * self.eval_eq_left * tm1 % p

    def evaluation_points_cubic_with_three_inputs(self, A, B, C, claim):  # sumcheck.rs:900-966
        p = self.p
        h = len(A) // 2
        t0 = tinf = 0
        for i in range(h):
            f = self.factor(i)
            t0 += (A[i] * B[i] - C[i]) * f
            tinf += (A[h + i] - A[i]) * (B[h + i] - B[i]) * f
        t0 %= p
        tinf %= p
        d = self.derive_deg2(t0, tinf, claim)
        if d is not None:
            return d
        e0, slope, em1 = self.eq_tau_0_a_inf[self.round - 1]  # fallback sumcheck.rs:1082-1130
        tm1 = sum(((2 * A[i] - A[h + i]) * (2 * B[i] - B[h + i]) - (2 * C[i] - C[h + i])) * self.factor(i)
                  for i in range(h)) % p
        q = self.eval_eq_left
        return e0 * q * t0 % p, slope * q * tinf % p, em1 * q * tm1 % p

    def evaluation_points_cubic_with_two_inputs(self, A, B, claim):  # sumcheck.rs:972-1033
        p = self.p
        h = len(A) // 2
        t0 = tinf = 0
        for i in range(h):
            f = self.factor(i)
            t0 += (A[i] * B[i] - 1) * f
            tinf += (A[h + i] - A[i]) * (B[h + i] - B[i]) * f
        t0 %= p
        tinf %= p
        d = self.derive_deg2(t0, tinf, claim)
        if d is not None:
            return d
        e0, slope, em1 = self.eq_tau_0_a_inf[self.round - 1]  # fallback sumcheck.rs:1134-1178
        tm1 = sum(((2 * A[i] - A[h + i]) * (2 * B[i] - B[h + i]) - 1) * self.factor(i) for i in range(h)) % p
        q = self.eval_eq_left
        return e0 * q * t0 % p, slope * q * tinf % p, em1 * q * tm1 % p

    def evaluation_points_quadratic_with_one_input(self, A, claim):  # sumcheck.rs:1039-1080
        p = self.p
        h = len(A) // 2
        t0 = sum(A[i] * self.factor(i) for i in range(h)) % p
        d = self.derive_deg1(t0, claim)
        if d is not None:
            return d
        e0, slope, em1 = self.eq_tau_0_a_inf[self.round - 1]
        tm1 = sum((2 * A[i] - A[h + i]) * self.factor(i) for i in range(h)) % p
        q = self.eval_eq_left
        return e0 * q * t0 % p, 0, em1 * q * tm1 % p

    def bound(self, r):  # sumcheck.rs:1226-1231
        tau = self.taus[self.round - 1]
        self.eval_eq_left = self.eval_eq_left * (1 - tau - r + 2 * r * tau) % self.p
        self.round += 1


def prove_quad_prod(p, claim, num_rounds, A, B, transcript):
    """SumcheckProof::prove_quad_prod (sumcheck.rs:199-242).  Returns (compressed polys, r, finals)."""
    A, B = list(A), list(B)
    rs, polys = [], []
    for _ in range(num_rounds):
        h = len(A) // 2
        e0 = sum(A[i] * B[i] for i in range(h)) % p
        bc = sum((A[h + i] - A[i]) * (B[h + i] - B[i]) for i in range(h)) % p
        poly = UniPoly.from_evals_deg2(p, [e0, (claim - e0) % p, bc])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        polys.append(poly.compressed())
        claim = poly.evaluate(r)
        A, B = bind_top(p, A, r), bind_top(p, B, r)
    return polys, rs, [A[0], B[0]]


def prove_cubic_with_three_inputs(p, claim, taus, A, B, C, transcript):
    """SumcheckProof::prove_cubic_with_three_inputs (sumcheck.rs:446-507)."""
    A, B, C = list(A), list(B), list(C)
    rs, polys = [], []
    eq = EqSumCheckInstance(p, taus)
    for _ in range(len(taus)):
        e0, lead, em1 = eq.evaluation_points_cubic_with_three_inputs(A, B, C, claim)
        poly = UniPoly.from_evals_deg3(p, [e0, (claim - e0) % p, lead, em1])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        polys.append(poly.compressed())
        claim = poly.evaluate(r)
        A, B, C = bind_top(p, A, r), bind_top(p, B, r), bind_top(p, C, r)
        eq.bound(r)
    return polys, rs, [A[0], B[0], C[0]]


def prove_batched_cubic(p, claim, taus, As, Bs, Cs, alphas, transcript):
    """SumcheckProof::prove_batched_cubic (sumcheck.rs:513-577) with
    EqSumCheckInstance::evaluation_points_batched_cubic (sumcheck.rs:755-835) and its tau = 0 fall-back
    (:838-890): sum_x eq(tau, x) sum_i alpha_i (A_i B_i - C_i)(x) for K instance triples.
    Returns (compressed polys, r, [[A_i(r), B_i(r), C_i(r)] for i])."""
    As, Bs, Cs = [list(v) for v in As], [list(v) for v in Bs], [list(v) for v in Cs]
    k = len(As)
    assert k > 0 and k == len(Bs) == len(Cs) == len(alphas)
    eq = EqSumCheckInstance(p, taus)
    rs, polys = [], []
    for _ in range(len(taus)):
        h = len(As[0]) // 2
        t0 = tinf = 0
        for idx in range(h):  # the reference's per-index accumulation over the K instances
            e0 = q = 0
            for i in range(k):
                e0 += alphas[i] * (As[i][idx] * Bs[i][idx] - Cs[i][idx])
                q += alphas[i] * (As[i][h + idx] - As[i][idx]) * (Bs[i][h + idx] - Bs[i][idx])
            f = eq.factor(idx)
            t0 += e0 * f
            tinf += q * f
        t0 %= p
        tinf %= p
        d = eq.derive_deg2(t0, tinf, claim)
        if d is None:  # tau = 0: third sum
            e0c, slope, em1c = eq.eq_tau_0_a_inf[eq.round - 1]
            tm1 = 0
            for idx in range(h):
                acc = 0
                for i in range(k):
                    ma, mb, mc = (2 * V[i][idx] - V[i][h + idx] for V in (As, Bs, Cs))
                    acc += alphas[i] * (ma * mb - mc)
                tm1 += acc * eq.factor(idx)
            ql = eq.eval_eq_left
            d = (e0c * ql * t0 % p, slope * ql * tinf % p, em1c * ql * tm1 % p)
        e0, lead, em1 = d
        poly = UniPoly.from_evals_deg3(p, [e0, (claim - e0) % p, lead, em1])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        polys.append(poly.compressed())
        claim = poly.evaluate(r)
        As, Bs, Cs = ([bind_top(p, v, r) for v in V] for V in (As, Bs, Cs))
        eq.bound(r)
    return polys, rs, [[As[i][0], Bs[i][0], Cs[i][0]] for i in range(k)]


def update_claim(p, claim, evals, r):
    """SumcheckProof::update_claim (sumcheck.rs:68-75)."""
    e0, c3, em1 = evals
    e1 = (claim - e0) % p
    half = pow(2, -1, p)
    a1 = ((e1 - em1) * half - c3) % p
    a2 = ((e1 + em1) * half - e0) % p
    return (e0 + r * (a1 + r * (a2 + r * c3))) % p


def prove_batch_eval(p, claims, num_rounds, polys, eq_points, coeffs, transcript):
    """SumcheckProof::prove_batch_eval (sumcheck.rs:251-351): instances of different sizes,
    e_i = sum_x P_i(x) eq(x_i, x), combined with `coeffs`."""
    k = len(claims)
    polys = [list(P) for P in polys]
    nmax = max(num_rounds)
    eqs = [EqSumCheckInstance(p, pts) for pts in eq_points]
    running = list(claims)
    e = sum(claims[i] * pow(2, nmax - num_rounds[i], p) * coeffs[i] for i in range(k)) % p
    rs, out = [], []
    for cur in range(nmax):
        rem = nmax - cur
        evals = []
        for i in range(k):
            if rem <= num_rounds[i]:
                e0, _, em1 = eqs[i].evaluation_points_quadratic_with_one_input(polys[i], running[i])
                evals.append((e0, 0, em1))
            else:
                sc = pow(2, rem - num_rounds[i] - 1, p) * claims[i] % p
                evals.append((sc, 0, sc))
        c0 = sum(evals[i][0] * coeffs[i] for i in range(k)) % p
        cm1 = sum(evals[i][2] * coeffs[i] for i in range(k)) % p
        c1 = (e - c0) % p
        quad = (c1 + cm1 - 2 * c0) * pow(2, -1, p) % p
        poly = UniPoly.from_evals_deg2(p, [c0, c1, quad])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        for i in range(k):
            if rem <= num_rounds[i]:
                running[i] = update_claim(p, running[i], evals[i], r)
                polys[i] = bind_top(p, polys[i], r)
                eqs[i].bound(r)
        e = poly.evaluate(r)
        out.append(poly.compressed())
    return out, rs, [P[0] for P in polys]


# ----------------------------------------------------------------------------------------
# Inner-product argument, literal restatement of provider/ipa_pc.rs:174-285 (key folded each round)
# ----------------------------------------------------------------------------------------
def commitment_transcript_bytes(P) -> bytes:
    """pedersen.rs:107-117: x || y || is_infinity (coordinates are (0, 0, 1) for the identity)."""
    if P is None:
        return to_repr(0) + to_repr(0) + b"\x01"
    return to_repr(P[0]) + to_repr(P[1]) + b"\x00"


def ipa_prove(curve: "Curve", ck, ck_c, comm_a, b_vec, c_claim, a_vec, transcript):
    """InnerProductArgument::prove.  ck: list of affine points (len >= n), ck_c: one affine point,
    U = (comm_a, b_vec, c).  Returns (L_vec, R_vec, a_hat)."""
    q = curve.q
    n = len(b_vec)
    assert len(a_vec) == n
    transcript.absorb_bytes(b"NoDS", b"IPA")  # dom_sep(protocol_name) (keccak.rs:162-167)
    ck = list(ck[:n])  # split_at(U.b_vec.len())
    transcript.absorb_bytes(b"U", commitment_transcript_bytes(comm_a) + to_repr(c_claim % q))  # :131-140
    r0 = transcript.squeeze(b"r")
    ck_c_s = curve.mul(r0, ck_c)  # ck_c.scale(&r)
    a, b = [x % q for x in a_vec], [x % q for x in b_vec]
    L_vec, R_vec = [], []
    while len(a) > 1:
        h = len(a) // 2
        c_L = sum(x * y for x, y in zip(a[:h], b[h:])) % q
        c_R = sum(x * y for x, y in zip(a[h:], b[:h])) % q
        L = curve.msm_naive(a[:h] + [c_L], ck[h:] + [ck_c_s])  # commit(ck_R.combine(ck_c), a_L || c_L, 0)
        R = curve.msm_naive(a[h:] + [c_R], ck[:h] + [ck_c_s])
        transcript.absorb_bytes(b"L", commitment_transcript_bytes(L))
        transcript.absorb_bytes(b"R", commitment_transcript_bytes(R))
        r = transcript.squeeze(b"r")
        ri = pow(r, -1, q)
        a = [(a[i] * r + ri * a[i + h]) % q for i in range(h)]
        b = [(b[i] * ri + r * b[i + h]) % q for i in range(h)]
        ck = [curve.add(curve.mul(ri, ck[i]), curve.mul(r, ck[i + h])) for i in range(h)]  # ck.fold(r_inv, r)
        L_vec.append(L)
        R_vec.append(R)
    return L_vec, R_vec, a[0]


def ipa_verify(curve: "Curve", ck, ck_c, comm_a, b_vec, c_claim, L_vec, R_vec, a_hat, transcript) -> bool:
    """InnerProductArgument::verify (provider/ipa_pc.rs:286-396), restated to pin `ipa_prove`."""
    q = curve.q
    n = len(b_vec)
    if n != 1 << len(L_vec) or len(L_vec) != len(R_vec) or len(L_vec) >= 32:
        return False
    ck = list(ck[:n])
    transcript.absorb_bytes(b"NoDS", b"IPA")
    transcript.absorb_bytes(b"U", commitment_transcript_bytes(comm_a) + to_repr(c_claim % q))
    r0 = transcript.squeeze(b"r")
    ck_c_s = curve.mul(r0, ck_c)
    P = curve.add(comm_a, curve.mul(c_claim % q, ck_c_s))
    rs = []
    for L, R in zip(L_vec, R_vec):
        transcript.absorb_bytes(b"L", commitment_transcript_bytes(L))
        transcript.absorb_bytes(b"R", commitment_transcript_bytes(R))
        rs.append(transcript.squeeze(b"r"))
    r_sq = [r * r % q for r in rs]
    r_inv = [pow(r, -1, q) for r in rs]
    r_inv_sq = [x * x % q for x in r_inv]
    s = [0] * n  # the vector with the tensor structure, :335-350
    s[0] = 1
    for x in r_inv:
        s[0] = s[0] * x % q
    for i in range(1, n):
        pos = i.bit_length() - 1
        s[i] = s[i - (1 << pos)] * r_sq[(len(L_vec) - 1) - pos] % q
    ck_hat = curve.msm_naive(s, ck)
    b_hat = sum(x * y for x, y in zip(b_vec, s)) % q
    P_hat = curve.msm_naive(r_sq + r_inv_sq + [1], list(L_vec) + list(R_vec) + [P])
    return P_hat == curve.msm_naive([a_hat % q, a_hat * b_hat % q], [ck_hat, ck_c_s])
