"""CPU restatement of the HyperKZG evaluation argument (src/provider/hyperkzg.rs:926-1245) on top of
the C oracle's field passes and MSM.  TEST INFRASTRUCTURE ONLY.

prove_core  = EvaluationEngine::prove with the transcript challenges r, q given (hyperkzg.rs:1076-1116)
verify_core = EvaluationEngine::verify (hyperkzg.rs:1119-1242) with the pairing check
              e(L, H) = e(R, tau*H) replaced by the equivalent group equation  L = [tau] R,
              which a TEST setup can evaluate because it knows tau (hyperkzg.rs:357-376 builds the
              test SRS [tau^i]G from a sampled tau in exactly this way).

The prover restatement is pinned by the verifier restatement: an honest proof must pass, a proof
with any message altered must fail (tests/test_oracle_hyperkzg.py).
"""
from . import coracle as co
from .pyref import CURVES, from_mont_bytes, mont_bytes


def setup_srs(cid: int, n: int, tau: int) -> bytes:
    """ck[i] = [tau^i] G as affine Montgomery bytes (hyperkzg.rs:357-376 `setup_from_tau`-style)."""
    c = CURVES[cid]
    g = c.affine_bytes(c.gen)
    out, t = [], 1
    for _ in range(n):
        out.append(co.scalar_mul(cid, g, t))
        t = t * tau % c.q
    return b"".join(out)


def _pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def _ints(p, b):
    return [from_mont_bytes(p, b[i:i + 32]) for i in range(0, len(b), 32)]


def prove_core(cid: int, ck: bytes, hat_P: bytes, x: list, r: int, q: int):
    """-> (com[ell-1] affine, v[ell][3] ints, w[3] affine)."""
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    ell = len(x)
    n = len(hat_P) // 32
    assert n == 1 << ell
    polys = [hat_P]
    for i in range(ell - 1):
        polys.append(co.kzg_fold(fid, polys[i], mont_bytes(p, x[ell - i - 1])))
    aff = c.affine_from_bytes
    com = [aff(co.msm(cid, f, ck[:2 * len(f)])) for f in polys[1:]]
    u = [r % p, (-r) % p, r * r % p]
    us = _pack(p, u)
    v = [_ints(p, co.poly_eval(fid, f, us)) for f in polys]
    B = co.rlc(fid, polys, _pack(p, [pow(q, k, p) for k in range(ell)]), n)
    w = []
    for ut in u:
        h = co.poly_div(fid, B, mont_bytes(p, ut))
        w.append(aff(co.msm(cid, h, ck[:2 * len(h)])))
    return com, v, w


def verify_core(cid: int, tau: int, C, x: list, y: int, com, v, w, r: int, q: int, d0: int) -> bool:
    """hyperkzg.rs:1119-1242 with challenges (r, q, d_0) given and L = [tau] R instead of the pairing."""
    c = CURVES[cid]
    p = c.q
    ell = len(x)
    if len(v) != ell or len(com) != ell - 1:
        return False
    for i in range(ell):  # consistency of (Y, ypos, yneg), :1143-1156
        ypos, yneg = v[i][0], v[i][1]
        Y = v[i + 1][2] if i + 1 < ell else y % p
        xi = x[ell - i - 1]
        if (2 * r * Y - (r * (1 - xi) * (ypos + yneg) + xi * (ypos - yneg))) % p != 0:
            return False
    d1 = d0 * d0 % p
    u = [r % p, (-r) % p, r * r % p]
    mult = (1 + d0 + d1) % p
    q_pows = [mult * pow(q, k, p) % p for k in range(ell)]
    B_u = []
    for i in range(3):
        acc = 0
        for vj in reversed(v):
            acc = (acc * q + vj[i]) % p
        B_u.append(acc)
    scalars = q_pows + [u[0], u[1] * d0 % p, u[2] * d1 % p, (-(B_u[0] + d0 * B_u[1] + d1 * B_u[2])) % p]
    points = [C] + list(com) + list(w) + [c.gen]
    L = c.msm_naive(scalars, points)
    R = c.add(c.add(w[0], c.mul(d0, w[1])), c.mul(d1, w[2]))
    return L == c.mul(tau, R)


# ---- the same with the transcript (hyperkzg.rs:861-898, 1099-1107, 1060-1070; verify :1119-1242) ---------
def _absorb_points(tr, label, pts):
    from .pyref import commitment_transcript_bytes
    tr.absorb_bytes(label, b"".join(commitment_transcript_bytes(P) for P in pts))


def prove(cid: int, ck: bytes, hat_P: bytes, x: list, tr):
    """EvaluationEngine::prove: r = H(com), q = H(v); squeezes the verifier's second challenge after W so that
    prover and verifier leave the transcript in the same state.  -> (com, w, v)."""
    from .pyref import to_repr
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    ell, n = len(x), len(hat_P) // 32
    assert n == 1 << ell
    polys = [hat_P]
    for i in range(ell - 1):
        polys.append(co.kzg_fold(fid, polys[i], mont_bytes(p, x[ell - i - 1])))
    aff = c.affine_from_bytes
    com = [aff(co.msm(cid, f, ck[:2 * len(f)])) for f in polys[1:]]
    _absorb_points(tr, b"c", com)
    r = tr.squeeze(b"c")
    u = [r % p, (-r) % p, r * r % p]
    v = [_ints(p, co.poly_eval(fid, f, _pack(p, u))) for f in polys]
    tr.absorb_bytes(b"v", b"".join(to_repr(e) for row in v for e in row))
    q = tr.squeeze(b"r")
    B = co.rlc(fid, polys, _pack(p, [pow(q, k, p) for k in range(ell)]), n)
    w = [aff(co.msm(cid, h, ck[:2 * len(h)])) for h in (co.poly_div(fid, B, mont_bytes(p, ut)) for ut in u)]
    _absorb_points(tr, b"W", w)
    tr.squeeze(b"d")
    return com, w, v


def verify(cid: int, tau: int, C, x: list, y: int, proof, tr) -> bool:
    """EvaluationEngine::verify with the challenges re-derived from the transcript."""
    from .pyref import to_repr
    com, w, v = proof
    _absorb_points(tr, b"c", com)
    r = tr.squeeze(b"c")
    tr.absorb_bytes(b"v", b"".join(to_repr(e) for row in v for e in row))
    q = tr.squeeze(b"r")
    _absorb_points(tr, b"W", w)
    d0 = tr.squeeze(b"d")
    if r == 0 or C is None:  # hyperkzg.rs:1138-1141
        return False
    return verify_core(cid, tau, C, x, y, com, v, w, r, q, d0)
