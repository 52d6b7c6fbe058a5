"""Device-resident mirror of the folding step -- the hot path of RecursiveSNARK::prove_step
(src/nova/mod.rs:456-564; §8a rows a19-a23):

  R1CSShape::commit_T / commit_T_relaxed            src/r1cs/mod.rs:578-664
  R1CSShape::is_sat_relaxed                         src/r1cs/mod.rs:474-535
  RelaxedR1CSWitness::fold / fold_relaxed           src/r1cs/mod.rs:1044-1107
  RelaxedR1CSInstance::fold / fold_relaxed          src/r1cs/mod.rs:1237-1292
  NIFS::prove / NIFSRelaxed::prove (orchestration)  src/nova/nifs.rs:36-73, 120-167

W, E and T never leave HBM: one step is  Z1 + Z2 -> 3 SpMVs -> T = AZ o BZ - u CZ - E1 (- E2) ->
commit(T) -> challenge -> W1 + r W2, E1 + r T (+ r^2 E2),  all `*_dev` calls on one stream.  The host
keeps what the Rust host keeps: the O(1) instance algebra (u, X, blinds), the random oracle that turns
comm_T into the challenge (Poseidon in the reference, passed in here as a callable) and three
commitment-sized group operations per fold.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import fields
from .native import check, lib
from .ppsnark import View, dev_copy, dev_scalar, dev_zeros
from .provider import CommitmentKey, Curve, DlogGroup, _cbuf, _jac_to_affine
from .spartan import DeviceVec, SparseMatrix


@dataclass
class R1CSInstance:          # r1cs/mod.rs:62-66
    comm_W: tuple | None
    X: list


@dataclass
class R1CSWitness:           # r1cs/mod.rs:55-59
    W: DeviceVec
    r_W: int = 0


@dataclass
class RelaxedR1CSInstance:   # r1cs/mod.rs:80-86
    comm_W: tuple | None
    comm_E: tuple | None
    X: list
    u: int

    @classmethod
    def default(cls, num_io: int):  # r1cs/mod.rs:1190-1200
        return cls(None, None, [0] * num_io, 0)


@dataclass
class RelaxedR1CSWitness:    # r1cs/mod.rs:69-76
    W: DeviceVec
    E: DeviceVec
    r_W: int = 0
    r_E: int = 0

    @classmethod
    def default(cls, num_vars: int, num_cons: int):  # r1cs/mod.rs:1016-1024
        return cls(dev_zeros(num_vars), dev_zeros(num_cons))


def _affine_bytes(curve: Curve, P) -> bytes:
    if P is None:
        return bytes(64)
    fid = curve.base_field
    return fields.to_mont_bytes(fid, P[0]) + fields.to_mont_bytes(fid, P[1])


def _lincomb(curve: Curve, terms):
    """sum_i k_i * P_i over commitments (the `comm_W_1 + comm_W_2 * r` of the instance folds): a 2-3
    term MSM over the points themselves."""
    fid = curve.scalar_field
    return DlogGroup(curve).vartime_multiscalar_mul(fields.pack(fid, [k for k, _ in terms]),
                                                    b"".join(_affine_bytes(curve, P) for _, P in terms))


@dataclass
class R1CSShape:
    """Device-resident `R1CSShape` (r1cs/mod.rs:30-47): the three matrices live behind spmv handles."""
    curve: Curve
    A: SparseMatrix
    B: SparseMatrix
    C: SparseMatrix
    num_cons: int
    num_vars: int
    num_io: int

    @property
    def fid(self) -> int:
        return Curve(self.curve).scalar_field

    def _z(self, W: DeviceVec, u: int, X: list) -> DeviceVec:
        """z = (W, u, X) on the device (r1cs/mod.rs:496, 593-594)."""
        z = DeviceVec(32 * (self.num_vars + 1 + self.num_io))
        check(lib().b200_memcpy_d2d(z.ptr, W.ptr, 32 * self.num_vars, None))
        tail = fields.pack(self.fid, [u] + list(X))
        check(lib().b200_memcpy_h2d(View(z, self.num_vars).ptr, _cbuf(tail), len(tail)))
        return z

    def multiply_vec_dev(self, z: DeviceVec):
        """(Az, Bz, Cz), r1cs/mod.rs:407-431."""
        out = tuple(DeviceVec(32 * self.num_cons) for _ in range(3))
        for M, o in zip((self.A, self.B, self.C), out):
            check(lib().b200_spmv_dev(M.handle, z.ptr, None, o.ptr, None, None))
        return out

    def _check_lengths(self, U, X2=None):
        if len(U.X) != self.num_io or (X2 is not None and len(X2) != self.num_io):
            raise ValueError("InvalidInputLength")  # r1cs/mod.rs:483-485

    def _commit(self, ck: CommitmentKey, v: DeviceVec, n: int, r: int):
        """CE::commit(ck, v, r) of a resident vector (pedersen.rs:263-270)."""
        out = DeviceVec(96)
        blind = dev_scalar(self.fid, r) if r else None
        check(lib().b200_commit_dev(ck.handle, v.ptr, n, blind.ptr if blind else None, out.ptr, None))
        return _jac_to_affine(Curve(self.curve), out.to_bytes(96))

    def _cross_term(self, ck, u1, X1, W1, E1, u2, X2, W2, E2, r_T):
        fid, p = self.fid, fields.MODULUS[self.fid]
        Z1, Z2 = self._z(W1, u1, X1), self._z(W2, u2, X2)
        zlen = self.num_vars + 1 + self.num_io
        Z = DeviceVec(32 * zlen)
        check(lib().b200_vec_add_dev(fid, Z1.ptr, Z2.ptr, zlen, Z.ptr, None))  # Mova §5.2: one multiply_vec of Z1 + Z2
        Az, Bz, Cz = self.multiply_vec_dev(Z)
        T = DeviceVec(32 * self.num_cons)
        u = dev_scalar(fid, (u1 + u2) % p)
        check(lib().b200_cross_term_dev(fid, Az.ptr, Bz.ptr, Cz.ptr, E1.ptr, E2.ptr if E2 is not None else None, u.ptr,
                                        self.num_cons, T.ptr, None))
        return T, self._commit(ck, T, self.num_cons, r_T)

    def commit_T(self, ck: CommitmentKey, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness, U2: R1CSInstance,
                 W2: R1CSWitness, r_T: int):
        """r1cs/mod.rs:578-627: T = AZ o BZ - (u1 + 1) CZ - E1 with Z = Z1 + Z2; -> (T, comm_T)."""
        self._check_lengths(U1, U2.X)
        return self._cross_term(ck, U1.u, U1.X, W1.W, W1.E, 1, U2.X, W2.W, None, r_T)

    def commit_T_relaxed(self, ck: CommitmentKey, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness,
                         U2: RelaxedR1CSInstance, W2: RelaxedR1CSWitness, r_T: int):
        """r1cs/mod.rs:631-664: ... - E1 - E2 with u = u1 + u2."""
        self._check_lengths(U1, U2.X)
        return self._cross_term(ck, U1.u, U1.X, W1.W, W1.E, U2.u, U2.X, W2.W, W2.E, r_T)

    def is_sat_relaxed(self, ck: CommitmentKey, U: RelaxedR1CSInstance, W: RelaxedR1CSWitness) -> bool:
        """r1cs/mod.rs:474-535: Az o Bz == u Cz + E row by row, and both commitments open."""
        self._check_lengths(U)
        z = self._z(W.W, U.u, U.X)             # named: the SpMVs below read it asynchronously
        Az, Bz, Cz = self.multiply_vec_dev(z)
        slack = DeviceVec(32 * self.num_cons)  # Az o Bz - u Cz - E must vanish
        u_dev = dev_scalar(self.fid, U.u)      # (kept in a variable: it must outlive the launch)
        check(lib().b200_cross_term_dev(self.fid, Az.ptr, Bz.ptr, Cz.ptr, W.E.ptr, None, u_dev.ptr,
                                        self.num_cons, slack.ptr, None))
        if any(slack.to_bytes(32 * self.num_cons)):
            return False
        return (U.comm_W == self._commit(ck, W.W, self.num_vars, W.r_W)
                and U.comm_E == self._commit(ck, W.E, self.num_cons, W.r_E))


def sample_random_instance_witness(ck: CommitmentKey, S: R1CSShape, Z: bytes, r_W: int, r_E: int):
    """R1CSShape::sample_random_instance_witness (r1cs/mod.rs:786-831) with the randomness supplied by the host
    (Z = (W, u, X): num_vars + 1 + num_io random scalars, Montgomery bytes): E = AZ o BZ - u CZ on the device,
    the two blinded commitments, -> (RelaxedR1CSInstance, RelaxedR1CSWitness)."""
    fid = S.fid
    zl = S.num_vars + 1 + S.num_io
    assert len(Z) == 32 * zl
    z = DeviceVec.from_bytes(Z)
    tail = fields.unpack(fid, Z[32 * S.num_vars:])
    u, X = tail[0], tail[1:]
    Az, Bz, Cz = S.multiply_vec_dev(z)
    E = DeviceVec(32 * S.num_cons)
    zero, u_dev = dev_zeros(S.num_cons), dev_scalar(fid, u)
    check(lib().b200_cross_term_dev(fid, Az.ptr, Bz.ptr, Cz.ptr, zero.ptr, None, u_dev.ptr, S.num_cons, E.ptr, None))
    W = dev_copy(z, S.num_vars)
    inst = RelaxedR1CSInstance(S._commit(ck, W, S.num_vars, r_W), S._commit(ck, E, S.num_cons, r_E), X, u)
    return inst, RelaxedR1CSWitness(W, E, r_W, r_E)  # the commits above synchronised: temporaries may go


def derandomize(ck: CommitmentKey, curve, U: RelaxedR1CSInstance, W: RelaxedR1CSWitness):
    """RelaxedR1CSWitness::derandomize + RelaxedR1CSInstance::derandomize (r1cs/mod.rs:1131-1142, 1294-1310;
    CE::derandomize, pedersen.rs:307-315): blinds set to zero, `h * r` subtracted from both commitments.
    -> (U', W', r_W, r_E).  Needs the host copy of h (a key built from bytes)."""
    curve = Curve(curve)
    p = fields.MODULUS[curve.scalar_field]
    assert ck.h is not None, "key has no blinding generator on the host"
    bf = curve.base_field
    h = (fields.from_mont_bytes(bf, ck.h[:32]), fields.from_mont_bytes(bf, ck.h[32:64]))
    comm_W = _lincomb(curve, [(1, U.comm_W), ((-W.r_W) % p, h)]) if W.r_W else U.comm_W
    comm_E = _lincomb(curve, [(1, U.comm_E), ((-W.r_E) % p, h)]) if W.r_E else U.comm_E
    return (RelaxedR1CSInstance(comm_W, comm_E, list(U.X), U.u), RelaxedR1CSWitness(W.W, W.E, 0, 0), W.r_W, W.r_E)


def fold_witness(fid: int, W1: RelaxedR1CSWitness, W2, T: DeviceVec, r_T: int, r: int, num_vars: int,
                 num_cons: int) -> RelaxedR1CSWitness:
    """RelaxedR1CSWitness::fold (W2: R1CSWitness) / fold_relaxed (W2: RelaxedR1CSWitness), r1cs/mod.rs:1044-1107."""
    p = fields.MODULUS[fid]
    rd = dev_scalar(fid, r)
    W = DeviceVec(32 * num_vars)
    check(lib().b200_axpy_dev(fid, W1.W.ptr, W2.W.ptr, rd.ptr, num_vars, W.ptr, None))       # W1 + r W2
    E = DeviceVec(32 * num_cons)
    check(lib().b200_axpy_dev(fid, W1.E.ptr, T.ptr, rd.ptr, num_cons, E.ptr, None))          # E1 + r T
    r_E = (W1.r_E + r * r_T) % p
    if isinstance(W2, RelaxedR1CSWitness):                                                   # + r^2 E2
        r2 = dev_scalar(fid, r * r % p)
        E2 = DeviceVec(32 * num_cons)
        check(lib().b200_axpy_dev(fid, E.ptr, W2.E.ptr, r2.ptr, num_cons, E2.ptr, None))
        E = E2
        r_E = (r_E + r * r % p * W2.r_E) % p
    check(lib().b200_sync())  # the challenge scalars above must outlive the launches
    return RelaxedR1CSWitness(W, E, (W1.r_W + r * W2.r_W) % p, r_E)


def fold_instance(curve: Curve, U1: RelaxedR1CSInstance, U2, comm_T, r: int) -> RelaxedR1CSInstance:
    """RelaxedR1CSInstance::fold (U2: R1CSInstance) / fold_relaxed, r1cs/mod.rs:1237-1292."""
    curve = Curve(curve)
    p = fields.MODULUS[curve.scalar_field]
    X = [(a + r * b) % p for a, b in zip(U1.X, U2.X)]
    comm_W = _lincomb(curve, [(1, U1.comm_W), (r, U2.comm_W)])
    if isinstance(U2, RelaxedR1CSInstance):
        comm_E = _lincomb(curve, [(1, U1.comm_E), (r, comm_T), (r * r % p, U2.comm_E)])
        u = (U1.u + r * U2.u) % p
    else:
        comm_E = _lincomb(curve, [(1, U1.comm_E), (r, comm_T)])
        u = (U1.u + r) % p
    return RelaxedR1CSInstance(comm_W, comm_E, X, u)


def nifs_prove(ck: CommitmentKey, S: R1CSShape, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness, U2, W2, r_T: int,
               challenge):
    """NIFS::prove (nifs.rs:36-73) / NIFSRelaxed::prove (nifs.rs:120-167) with the random oracle passed
    in: `challenge(comm_T) -> r` stands for `comm_T.absorb_in_ro(&mut ro); ro.squeeze(..)` (the RO's
    other inputs -- pp digest, U2 -- are the caller's).  -> (comm_T, (U, W))."""
    relaxed = isinstance(U2, RelaxedR1CSInstance)
    T, comm_T = (S.commit_T_relaxed if relaxed else S.commit_T)(ck, U1, W1, U2, W2, r_T)
    r = challenge(comm_T)
    U = fold_instance(S.curve, U1, U2, comm_T, r)
    W = fold_witness(S.fid, W1, W2, T, r_T, r, S.num_vars, S.num_cons)
    return comm_T, (U, W)


class ResidentView:
    """Non-owning handle on a device vector that something else keeps alive (e.g. the witness a
    WitnessStream hands back); quacks like DeviceVec for the calls above (`.ptr`)."""

    def __init__(self, ptr, owner):
        self.ptr, self.owner = ptr, owner


def fold_streamed_step(ck: CommitmentKey, S: R1CSShape, U1: RelaxedR1CSInstance, W1: RelaxedR1CSWitness, stream,
                       chunks, X2: list, r_W: int, r_T: int, challenge):
    """One folding step the way prove_step produces it (nova/mod.rs:456-564): the fresh witness arrives in
    chunks while the circuit is synthesised (`stream`: provider.WitnessStream over S.num_vars; `chunks`: an
    iterable of Montgomery byte strings, consumed as they are produced), its commitment comes out of
    `stream.finish`, and the resident copy goes straight into NIFS::prove without touching the host again.
    -> (U2, comm_T, (U, W)); call stream.reset() before the next step."""
    fid = S.fid
    for c in chunks:
        stream.append(c)
    comm_W2 = stream.finish(fields.to_mont_bytes(fid, r_W) if r_W else None)
    U2 = R1CSInstance(comm_W2, list(X2))
    W2 = R1CSWitness(ResidentView(stream.d_witness, stream), r_W)
    comm_T, (U, W) = nifs_prove(ck, S, U1, W1, U2, W2, r_T, challenge)
    return U2, comm_T, (U, W)
