"""Parity of the streaming field kernels (cross-term T, folds, Z1+Z2, bind_poly_var_top)."""
import pytest

from oracle.pyref import FIELD_MODULUS, mont_bytes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [1, 2, 33, 4096, 100003])
def test_cross_term_and_folds(b200, oracle, fid, n):
    v = [oracle.gen_scalars(fid, 10 * k + n, n) for k in range(5)]
    u = oracle.gen_scalars(fid, 5, 1)
    assert b200.cross_term(fid, v[0], v[1], v[2], v[3], u) == oracle.cross_term(fid, v[0], v[1], v[2], v[3], None, u)
    assert b200.cross_term(fid, v[0], v[1], v[2], v[3], u, v[4]) == \
        oracle.cross_term(fid, v[0], v[1], v[2], v[3], v[4], u)
    assert b200.fold_witness(fid, v[0], v[1], u) == oracle.axpy(fid, v[0], v[1], u)
    assert b200.vec_add(fid, v[0], v[1]) == oracle.vec_add(fid, v[0], v[1])


@pytest.mark.parametrize("fid", [0, 3])
def test_edge_values(b200, oracle, fid):
    """0, 1, p-1 operands: results stay canonical (fully reduced) bit-for-bit."""
    p = FIELD_MODULUS[fid]
    vals = [0, 1, p - 1, p - 2, 2, (1 << 255) % p]
    a = b"".join(mont_bytes(p, x) for x in vals for _ in vals)
    b = b"".join(mont_bytes(p, y) for _ in vals for y in vals)
    for r in (0, 1, p - 1):
        rb = mont_bytes(p, r)
        assert b200.fold_witness(fid, a, b, rb) == oracle.axpy(fid, a, b, rb)
        assert b200.cross_term(fid, a, b, a, b, rb) == oracle.cross_term(fid, a, b, a, b, None, rb)
    assert b200.vec_add(fid, a, b) == oracle.vec_add(fid, a, b)


@pytest.mark.parametrize("fid", [0, 2])
def test_bind_poly_var_top(b200, oracle, fid):
    """evaluate == repeated bind (multilinear.rs:391-405) + KAT table [0,0,0,1,0,1,0,2] at (1,1,1) = 2."""
    p = FIELD_MODULUS[fid]
    one = mont_bytes(p, 1)
    z = oracle.field_from_u64(fid, [0, 0, 0, 1, 0, 1, 0, 2])
    for _ in range(3):
        z = b200.bind_poly_var_top(fid, z, one)
    assert z == mont_bytes(p, 2)
    n = 1 << 12
    z = oracle.gen_scalars(fid, 1, n)
    zo = z
    for k in range(12):
        r = oracle.gen_scalars(fid, 100 + k, 1)
        z = b200.bind_poly_var_top(fid, z, r)
        zo = oracle.bind_top(fid, zo, r)
        assert z == zo


def test_large_stream(b200, oracle):
    """2^20 rows (the T of a 2^20-constraint fold): spot-check against the oracle."""
    fid, n = 0, 1 << 20
    v = [oracle.gen_scalars(fid, 50 + k, n) for k in range(4)]
    u = oracle.gen_scalars(fid, 6, 1)
    assert b200.cross_term(fid, v[0], v[1], v[2], v[3], u) == oracle.cross_term(fid, v[0], v[1], v[2], v[3], None, u)
