"""The Poseidon RO restatement (oracle/poseidon_ref.py) against what the reference pins and against the product's
independent host mirror (nova_b200/poseidon.py).  CPU only.

The reference's only literals for this path are the IO-pattern tag values of
src/frontend/gadgets/poseidon/sponge/api.rs:270-316; it holds no digest literal, so digests are pinned structurally:
two independently written derivations of the constants agree, the round numbers are the published ones, the matrix
is MDS-shaped (symmetric Cauchy, invertible), the permutation is a bijection on test points (inverse through M^-1 and
fifth roots), and the sponge / RO bookkeeping follows api.rs:205-243 and poseidon.rs:93-127 literally."""
import pytest

from nova_b200 import fields
from nova_b200 import poseidon as pp
from oracle import poseidon_ref as pr
from oracle.pyref import FIELD_MODULUS, SplitMix64


def test_io_pattern_tags_from_the_reference():
    T = pr.io_pattern_value
    assert T([], 0) == 0
    assert T([], 123) == 340282366920938463463374607431768191899
    assert T([("A", 2), ("S", 2)], 0) == 340282366920938463463374607090318361668
    assert T([("A", 2), ("S", 2)], 1) == 340282366920938463463374607090314341989
    assert T([("A", 1), ("A", 1), ("S", 2)], 0) == 340282366920938463463374607090318361668  # runs coalesce
    assert T([("A", 1), ("A", 1), ("S", 1), ("S", 1)], 0) == 340282366920938463463374607090318361668
    assert pp.io_pattern_tag(2, 2, 0) == T([("A", 2), ("S", 2)], 0) and pp.io_pattern_tag(0, 0, 123) == T([], 123)
    for n in (1, 9, 24, 25, 100):
        assert pp.io_pattern_tag(n) == T([("A", n), ("S", 1)], 0)


def test_round_numbers_are_the_published_ones():
    """neptune / Filecoin parameters for a ~256-bit field at 128-bit security: R_F = 8 and R_P = 55, 56, 57, 57, 59, 60
    for arities 2, 4, 8, 11, 24, 36; both derivations (f32 arithmetic, the loop-variable quirk) give them."""
    for arity, rp in ((2, 55), (4, 56), (5, 56), (8, 57), (11, 57), (24, 59), (36, 60)):
        assert pr.round_numbers(arity) == (8, rp) == pp.round_numbers(arity)


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("arity", [5, 24])
def test_two_derivations_of_the_constants_agree(fid, arity):
    p = FIELD_MODULUS[fid]
    assert p == fields.MODULUS[fid]
    r_f, r_p = pr.round_numbers(arity)
    t = arity + 1
    a, b = pr.round_constants(p, t, r_f, r_p), pp.grain_constants(p, t, r_f, r_p)
    assert a == b and len(a) == (r_f + r_p) * t and all(0 <= x < p for x in a) and len(set(a)) == len(a)
    m = pr.mds(p, t)
    assert m == pp.cauchy_mds(p, t)
    assert all(m[i][j] == m[j][i] for i in range(t) for j in range(t))  # product_mds relies on the symmetry (mds.rs:132)
    assert all(m[i][j] * (i + t + j) % p == 1 for i in range(t) for j in range(t))


def _mat_inv(p, m):
    n = len(m)
    a = [row[:] + [int(i == j) for j in range(n)] for i, row in enumerate(m)]
    for c in range(n):
        piv = next(r for r in range(c, n) if a[r][c])
        a[c], a[piv] = a[piv], a[c]
        inv = pow(a[c][c], -1, p)
        a[c] = [x * inv % p for x in a[c]]
        for r in range(n):
            if r != c and a[r][c]:
                f = a[r][c]
                a[r] = [(x - f * y) % p for x, y in zip(a[r], a[c])]
    return [row[n:] for row in a]


def test_permutation_is_invertible_round_by_round():
    """undo the plain schedule with M^-1 and fifth roots: the restated permutation is a bijection built from exactly the
    three layers the paper names (a wrong round order or S-box placement would not invert)"""
    p = FIELD_MODULUS[0]
    c = pr.cached_constants(p, 5)
    t = c.t
    minv = _mat_inv(p, c.m)
    e5 = pow(5, -1, p - 1)
    rng = SplitMix64(7)
    x = [rng.field(p) for _ in range(t)]
    y = c.permute(x)
    s = list(y)
    half = c.r_f // 2
    for r in reversed(range(c.r_f + c.r_p)):
        s = [sum(s[i] * minv[i][j] for i in range(t)) % p for j in range(t)]
        if r < half or r >= half + c.r_p:
            s = [pow(v, e5, p) for v in s]
        else:
            s[0] = pow(s[0], e5, p)
        s = [(v - k) % p for v, k in zip(s, c.rc[r * t:(r + 1) * t])]
    assert s == x and y != x


def test_sponge_and_ro_bookkeeping():
    p = FIELD_MODULUS[1]
    c = pr.cached_constants(p, 5)
    rng = SplitMix64(11)
    xs = [rng.field(p) for _ in range(13)]  # 13 > rate 5: three permutations while absorbing... (2 full blocks + rest)
    # manual sponge: tag in slot 0, add into the rate, permute when full, once more before the squeeze
    st = [pr.io_pattern_value([("A", 13), ("S", 1)], 0)] + [0] * 5
    for k, blk in enumerate((xs[0:5], xs[5:10], xs[10:13])):
        if k:
            st = c.permute(st)
        for i, e in enumerate(blk):
            st[1 + i] = (st[1 + i] + e) % p
    st = c.permute(st)
    assert pr.sponge_hash(c, xs) == st[1]
    ro = pr.PoseidonRO(p, 5)
    for e in xs:
        ro.absorb(e)
    r1 = ro.squeeze(128)
    assert r1 == st[1] & ((1 << 128) - 1) and ro.state == [st[1]]
    ro.absorb(5)
    r2 = ro.squeeze(250, start_with_one=True)  # the previous hash is part of the next input; top bit forced
    assert r2 >> 249 == 1 and r2 != r1
    assert pr.sponge_hash(c, [st[1], 5]) & ((1 << 249) - 1) == r2 & ((1 << 249) - 1)
