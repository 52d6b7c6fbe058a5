"""The reference's folding fixture, replayed through the CPU oracle: the tiny cubic R1CS
x^3 + x + 5 = y (r1cs/mod.rs:1349-1413) is folded twice into the default running instance and the
result must satisfy the relaxed equation  AZ o BZ = u*CZ + E  (nova/nifs.rs:299-351 `execute_sequence`,
r1cs/mod.rs:490-560 `is_sat_relaxed`), plus one relaxed-relaxed fold (nifs.rs:120-167).

This pins the oracle's cross-term (commit_T / commit_T_relaxed, r1cs/mod.rs:578-664), the witness
and instance folds (r1cs/mod.rs:1044-1107, 1237-1292) and SpMV against an equation none of them is
defined by.  The GPU parity tests then compare the CUDA kernels with these same oracle functions."""
import pytest

from oracle.pyref import FIELD_MODULUS, SplitMix64, from_mont_bytes, mont_bytes

NUM_VARS, NUM_IO, NUM_CONS = 3, 2, 4


def tiny_r1cs():
    """(row, col, val) triplets exactly as pushed at r1cs/mod.rs:1371-1395; z = (vars, u, inputs)."""
    nv = NUM_VARS
    A = [(0, nv + 1, 1), (1, 0, 1), (2, 1, 1), (2, nv + 1, 1), (3, 2, 1), (3, nv, 5)]
    B = [(0, nv + 1, 1), (1, nv + 1, 1), (2, nv, 1), (3, nv, 1)]
    C = [(0, 0, 1), (1, 1, 1), (2, 2, 1), (3, nv + 2, 1)]
    return A, B, C


def csr(M, rows):
    data, idx, ptr = [], [], [0]
    for r in range(rows):
        for (rr, c, v) in M:
            if rr == r:
                data.append(v)
                idx.append(c)
        ptr.append(len(idx))
    return data, idx, ptr


def witness(p, x):
    """Satisfying assignment for input x: vars (x^2, x^3, x^3 + x), io (x, y)."""
    z0, z1 = x * x % p, x * x * x % p
    z2 = (z1 + x) % p
    return [z0, z1, z2], [x, (z2 + 5) % p]


class Oracle:
    def __init__(self, oracle, fid):
        self.o, self.fid, self.p = oracle, fid, FIELD_MODULUS[fid]
        self.mats = [csr(M, NUM_CONS) for M in tiny_r1cs()]

    def pack(self, xs):
        return b"".join(mont_bytes(self.p, x) for x in xs)

    def ints(self, b):
        return [from_mont_bytes(self.p, b[i:i + 32]) for i in range(0, len(b), 32)]

    def multiply_vec(self, z):
        zb = self.pack(z)
        return [self.o.spmv(self.fid, self.pack(d), i, pt, zb) for (d, i, pt) in self.mats]

    def is_sat_relaxed(self, W, E, u, X):
        az, bz, cz = (self.ints(v) for v in self.multiply_vec(W + [u] + X))
        return all((a * b - u * c - e) % self.p == 0 for a, b, c, e in zip(az, bz, cz, E))

    def fold(self, W1, E1, u1, X1, W2, X2, r, E2=None, u2=1):
        """NIFS::prove / NIFSRelaxed::prove on the vectors (commitments and the RO are O(1) host work)."""
        p = self.p
        Z = [(a + b) % p for a, b in zip(W1 + [u1] + X1, W2 + [u2] + X2)]
        u = (u1 + u2) % p
        az, bz, cz = self.multiply_vec(Z)
        T = self.o.cross_term(self.fid, az, bz, cz, self.pack(E1), self.pack(E2) if E2 is not None else None,
                              self.pack([u]))
        rb = self.pack([r])
        W = self.ints(self.o.axpy(self.fid, self.pack(W1), self.pack(W2), rb))
        E = self.ints(self.o.axpy(self.fid, self.pack(E1), T, rb))
        if E2 is not None:  # E1 + r T + r^2 E2 (r1cs/mod.rs:1075-1107)
            E = [(e + r * r % p * e2) % p for e, e2 in zip(E, E2)]
        return W, E, (u1 + r * u2) % p, [(a + r * b) % p for a, b in zip(X1, X2)]


@pytest.mark.parametrize("fid", [0, 1, 2, 3])
def test_fold_twice_then_relaxed_sat(oracle, fid):
    o = Oracle(oracle, fid)
    p = o.p
    rng = SplitMix64(50 + fid)
    W1, X1 = witness(p, rng.field(p))
    W2, X2 = witness(p, 3)
    assert o.is_sat_relaxed(W1, [0] * NUM_CONS, 1, X1) and o.is_sat_relaxed(W2, [0] * NUM_CONS, 1, X2)
    # default running instance: W = 0, E = 0, u = 0, X = 0 (r1cs/mod.rs:1016-1024, 1190-1200)
    run = ([0] * NUM_VARS, [0] * NUM_CONS, 0, [0] * NUM_IO)
    for (W, X) in ((W1, X1), (W2, X2)):
        run = o.fold(*run, W, X, rng.field(p))
        assert o.is_sat_relaxed(*run)
    # a wrong cross-term must break it: fold with E perturbed
    Wb, Eb, ub, Xb = run
    assert not o.is_sat_relaxed(Wb, [(Eb[0] + 1) % p] + Eb[1:], ub, Xb)
    # relaxed + relaxed (NIFSRelaxed, commit_T_relaxed): fold the running instance with another one
    other = o.fold([0] * NUM_VARS, [0] * NUM_CONS, 0, [0] * NUM_IO, *witness(p, 7), rng.field(p))
    other = o.fold(*other, *witness(p, 11), rng.field(p))
    W2r, E2r, u2r, X2r = other
    both = o.fold(*run, W2r, X2r, rng.field(p), E2=E2r, u2=u2r)
    assert o.is_sat_relaxed(*both)
