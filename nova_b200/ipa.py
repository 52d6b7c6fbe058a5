"""Inner-product argument prover (src/provider/ipa_pc.rs:174-285) on the device.

The reference folds the commitment key every round (`ck.fold`, pedersen.rs:484-497) and commits
over the folded key.  Here the ORIGINAL key stays registered (window tables resident) and the fold
weights move into the scalars (include/nova_b200.h, "inner-product argument"): per round two
device inner products, one kernel building the two scalar vectors, two `b200_commit_dev` calls whose
blinding slot carries the `c * r0 * ck_c` term, and three element-wise kernels (fold a, fold b,
update weights).  L_vec, R_vec and a_hat are group / field elements, hence identical to the
reference's.  The host keeps the transcript and the O(1) algebra.
"""
from __future__ import annotations

import ctypes

from . import fields
from .native import check, lib
from .provider import CommitmentKey, Curve, _jac_to_affine
from .spartan import DeviceVec


def commitment_transcript_bytes(P) -> bytes:
    """pedersen.rs:107-117."""
    if P is None:
        return bytes(64) + b"\x01"
    return int(P[0]).to_bytes(32, "little") + int(P[1]).to_bytes(32, "little") + b"\x00"


class InnerProductArgument:
    @staticmethod
    def prove(curve, ck: CommitmentKey, comm_a, b_vec: bytes, c_claim: int, a_vec: bytes, transcript):
        """`ck` must be registered over the n bases with `h = ck_c` (the single generator the reference
        keeps in `ck_c`).  b_vec / a_vec are Montgomery field-element vectors of n = 2^l entries."""
        curve = Curve(curve)
        fid = curve.scalar_field
        q = fields.MODULUS[fid]
        size = lambda v: v.nbytes if isinstance(v, DeviceVec) else len(v)
        n = size(b_vec) // 32
        if size(a_vec) != size(b_vec):
            raise ValueError("InvalidInputLength")  # ipa_pc.rs:187-189
        assert n and n & (n - 1) == 0 and n <= len(ck) and ck.has_h
        L = lib()
        transcript.absorb_bytes(b"NoDS", b"IPA")
        transcript.absorb_bytes(b"U", commitment_transcript_bytes(comm_a) + int(c_claim % q).to_bytes(32, "little"))
        r0 = transcript.squeeze(b"r")
        def working(v):  # the folds overwrite both vectors: resident inputs are copied device-to-device
            if not isinstance(v, DeviceVec):
                return DeviceVec.from_bytes(v)
            c = DeviceVec(32 * n)
            check(L.b200_memcpy_d2d(c.ptr, v.ptr, 32 * n, None))
            return c
        a, b = working(a_vec), working(b_vec)
        w, sL, sR = DeviceVec(32 * n), DeviceVec(32 * n), DeviceVec(32 * n)
        a2, b2 = DeviceVec(16 * n), DeviceVec(16 * n)
        out = DeviceVec(96 * 2)
        ip = DeviceVec(64)
        check(L.b200_ipa_weights_dev(fid, w.ptr, n, 0, None, None, None))  # w := 1
        L_vec, R_vec = [], []
        nk = n
        off = lambda v, elems: ctypes.c_void_p(v.ptr.value + 32 * elems)
        while nk > 1:
            h = nk // 2
            # c_L = <a_L, b_R>, c_R = <a_R, b_L>
            check(L.b200_sc_eval_dev(fid, 11, a.ptr, off(b, h), None, h, None, None, 0, ip.ptr, None))
            check(L.b200_sc_eval_dev(fid, 11, off(a, h), b.ptr, None, h, None, None, 0, off(ip, 1), None))
            c_L, c_R = fields.unpack(fid, ip.to_bytes(64))
            check(L.b200_ipa_scalars_dev(fid, a.ptr, w.ptr, n, nk, sL.ptr, sR.ptr, None))
            blind = DeviceVec.from_bytes(fields.pack(fid, [c_L * r0 % q, c_R * r0 % q]))
            check(L.b200_commit_dev(ck.handle, sL.ptr, n, blind.ptr, out.ptr, None))
            check(L.b200_commit_dev(ck.handle, sR.ptr, n, off(blind, 1), ctypes.c_void_p(out.ptr.value + 96), None))
            raw = out.to_bytes(192)
            Lk, Rk = _jac_to_affine(curve, raw[:96]), _jac_to_affine(curve, raw[96:])
            transcript.absorb_bytes(b"L", commitment_transcript_bytes(Lk))
            transcript.absorb_bytes(b"R", commitment_transcript_bytes(Rk))
            r = transcript.squeeze(b"r")
            ri = pow(r, -1, q)
            rr = DeviceVec.from_bytes(fields.pack(fid, [r, ri]))
            check(L.b200_fold_halves_dev(fid, a.ptr, nk, rr.ptr, off(rr, 1), a2.ptr, None))  # a_L r + r^-1 a_R
            check(L.b200_fold_halves_dev(fid, b.ptr, nk, off(rr, 1), rr.ptr, b2.ptr, None))  # b_L r^-1 + r b_R
            check(L.b200_ipa_weights_dev(fid, w.ptr, n, nk, rr.ptr, off(rr, 1), None))
            check(L.b200_sync())
            a, a2 = a2, a
            b, b2 = b2, b
            L_vec.append(Lk)
            R_vec.append(Rk)
            nk = h
        a_hat = fields.unpack(fid, a.to_bytes(32))[0]
        return L_vec, R_vec, a_hat
