"""Host-side mirror of spartan::snark::RelaxedR1CSSNARK::prove up to the evaluation argument
(src/spartan/snark.rs:113-256), every O(N) step on the device:

  z = (W, u, X), (Az, Bz, Cz) = S.multiply_vec(z)                 snark.rs:131-147   b200_spmv_dev x3
  poly_uCz_E = u*Cz + E                                            snark.rs:148-150   b200_axpy_dev
  outer sum-check  prove_cubic_with_three_inputs                   snark.rs:159-166   b200_sumcheck_cubic3 *
  claim_Cz = Cz(r_x), eval_E = E(r_x)                              snark.rs:170-171   b200_mle_eval_dev
  evals_rx, compute_eval_table_sparse, A + r B + r^2 C             snark.rs:181-195   b200_eq_table_dev,
                                                                                      b200_spmv_t_dev x3, b200_rlc_dev
  inner sum-check  prove_quad_prod over (poly_ABC, poly_z)         snark.rs:202-208   b200_sumcheck_quad_prod *
  eval_W = W(r_y[1..])                                             snark.rs:218       b200_mle_eval_dev
  batch_eval_reduce -> (C, x, e) and the batched polynomial        spartan/mod.rs:377-432

  * `device_transcript=True` runs each of the two sum-check loops as ONE call with the Keccak transcript
    on the device (csrc/capi_sumcheck.inc); False keeps the per-round host transcript (the D2H ->
    algebra -> H2D path).  Both produce identical prover messages.

W, E and the batched opening polynomial stay resident; the caller hands the result to the evaluation
argument (spartan.hyperkzg_prove_resident).  The transcript object is passed in and advanced exactly
as E::TE is in the reference (needs `absorb_bytes`, `squeeze` and the serialisable fields `round`,
`state`, `buf`).
"""
from __future__ import annotations

import ctypes

from . import fields
from .native import check, lib
from .ppsnark import (View, _as_dev, _mle_eval, _prove_cubic3_resident, _rlc_dev, commitment_transcript_bytes,
                      dev_scalar, dev_zeros, to_repr)
from .provider import CommitmentKey, Curve, DlogGroup, _cbuf
from .spartan import DeviceVec, SumcheckProof


def _affine_bytes(curve: Curve, P) -> bytes:
    if P is None:
        return bytes(64)
    fid = curve.base_field
    return fields.to_mont_bytes(fid, P[0]) + fields.to_mont_bytes(fid, P[1])


def prove_core(curve, ck: CommitmentKey | None, S: dict, U: dict, W: dict, vk_digest: int, transcript,
               device_transcript: bool = True, timings: dict | None = None):
    """S: dict(num_cons, num_vars, A, B, C) with A/B/C `spartan.SparseMatrix` (regular shape: powers of
    two, num_io < num_vars).  U: dict(comm_W, comm_E (affine or None), u, X: ints).  W: dict(W, E) as
    Montgomery bytes or DeviceVec.  Returns the proof fields of RelaxedR1CSSNARK (without eval_arg), the
    joint opening claim (batched_c, batched_x, batched_e) and the batched polynomial (DeviceVec)."""
    import time
    curve = Curve(curve)
    t_last = [time.perf_counter()]

    def mark(name):
        if timings is not None:
            check(lib().b200_sync())
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + now - t_last[0]
            t_last[0] = now
    fid = curve.scalar_field
    p = fields.MODULUS[fid]
    num_cons, num_vars = S["num_cons"], S["num_vars"]
    nrx, nry = num_cons.bit_length() - 1, num_vars.bit_length()
    assert 1 << nrx == num_cons and 1 << (nry - 1) == num_vars and len(U["X"]) < num_vars
    tr = transcript
    tr.absorb_bytes(b"vk", to_repr(vk_digest % p))
    tr.absorb_bytes(b"U", commitment_transcript_bytes(U["comm_W"]) + commitment_transcript_bytes(U["comm_E"])
                    + to_repr(U["u"] % p) + b"".join(to_repr(x % p) for x in U["X"]))
    Wd, Ed = _as_dev(W["W"]), _as_dev(W["E"])
    # poly_z = (W, u, X) zero-extended to 2 * num_vars (snark.rs:197-200); also the SpMV input
    z = dev_zeros(2 * num_vars)
    check(lib().b200_memcpy_d2d(z.ptr, Wd.ptr, 32 * num_vars, None))
    check(lib().b200_memcpy_h2d(View(z, num_vars).ptr, _cbuf(fields.pack(fid, [U["u"]] + list(U["X"]))),
                                32 * (1 + len(U["X"]))))
    tau = [tr.squeeze(b"t") for _ in range(nrx)]
    Az, Bz, Cz = (DeviceVec(32 * num_cons) for _ in range(3))
    for M, out in ((S["A"], Az), (S["B"], Bz), (S["C"], Cz)):
        check(lib().b200_spmv_dev(M.handle, z.ptr, None, out.ptr, None, None))
    u_dev = dev_scalar(fid, U["u"])
    uCz_E = DeviceVec(32 * num_cons)
    check(lib().b200_axpy_dev(fid, Ed.ptr, Cz.ptr, u_dev.ptr, num_cons, uCz_E.ptr, None))  # E + u*Cz
    mark("spmv")
    if device_transcript:
        sc_outer, r_x, claims_outer = SumcheckProof.prove_cubic_with_three_inputs_device(fid, 0, tau, Az, Bz, uCz_E, tr)
    else:
        sc_outer, r_x, claims_outer = _prove_cubic3_resident(fid, 0, tau, Az, Bz, uCz_E, num_cons, tr)
    claim_Az, claim_Bz = claims_outer[0], claims_outer[1]
    rx_dev = DeviceVec.from_bytes(fields.pack(fid, r_x))
    claim_Cz = _mle_eval(fid, Cz, nrx, rx_dev)
    eval_E = _mle_eval(fid, Ed, nrx, rx_dev)
    tr.absorb_bytes(b"claims_outer", b"".join(to_repr(x) for x in (claim_Az, claim_Bz, claim_Cz, eval_E)))
    mark("outer_sumcheck")
    r = tr.squeeze(b"r")
    claim_inner_joint = (claim_Az + r * claim_Bz + r * r * claim_Cz) % p
    evals_rx = DeviceVec(32 * num_cons)
    check(lib().b200_eq_table_dev(fid, rx_dev.ptr, nrx, evals_rx.ptr, None))
    tabs = [DeviceVec(32 * 2 * num_vars) for _ in range(3)]  # compute_eval_table_sparse, spartan/mod.rs:497-534
    for M, out in zip((S["A"], S["B"], S["C"]), tabs):
        check(lib().b200_spmv_t_dev(M.handle, evals_rx.ptr, 2 * num_vars, out.ptr, None))
    poly_ABC = DeviceVec(32 * 2 * num_vars)
    _rlc_dev(fid, tabs, [1, r, r * r % p], 2 * num_vars, poly_ABC)
    mark("eval_tables")
    if device_transcript:
        sc_inner, r_y, _ = SumcheckProof.prove_quad_prod_device(fid, claim_inner_joint, nry, poly_ABC, z, tr)
    else:
        sc_inner, r_y, _ = _prove_quad_prod_resident(fid, claim_inner_joint, nry, poly_ABC, z, tr)
    eval_W = _mle_eval(fid, Wd, nry - 1, DeviceVec.from_bytes(fields.pack(fid, r_y[1:])))
    tr.absorb_bytes(b"w", to_repr(eval_W))
    mark("inner_sumcheck")
    # batch_eval_reduce (spartan/mod.rs:377-432)
    u_vec = [(U["comm_W"], r_y[1:], eval_W), (U["comm_E"], r_x, eval_E)]
    num_rounds = [len(x) for (_, x, _) in u_vec]
    rho = tr.squeeze(b"r")
    powers = [pow(rho, i, p) for i in range(len(u_vec))]
    sc_batch, r_b, evals_batch = SumcheckProof.prove_batch_eval(fid, [e for (_, _, e) in u_vec], num_rounds,
                                                                [Wd, Ed], [x for (_, x, _) in u_vec], powers, tr)
    tr.absorb_bytes(b"l", b"".join(to_repr(x) for x in evals_batch))
    c = tr.squeeze(b"c")
    nmax = len(r_b)
    batched_e, coeffs = 0, []
    for i, (ev, nv) in enumerate(zip(evals_batch, num_rounds)):  # PolyEvalInstance::batch_diff_size, mod.rs:304-344
        lag = 1
        for rr in r_b[:nmax - nv]:
            lag = lag * (1 - rr) % p
        g = pow(c, i, p)
        coeffs.append(g)
        batched_e = (batched_e + g * lag * ev) % p
    # C = sum_i c^i C_i : a two-term MSM over the commitments themselves
    batched_c = DlogGroup(curve).vartime_multiscalar_mul(
        fields.pack(fid, coeffs), b"".join(_affine_bytes(curve, cm) for (cm, _, _) in u_vec))
    size_max = max(num_vars, num_cons)
    batched_poly = DeviceVec(32 * size_max)  # PolyEvalWitness::batch_diff_size, mod.rs:165-222
    _rlc_diff(fid, [Wd, Ed], [num_vars, num_cons], coeffs, size_max, batched_poly)
    mark("batch_eval_reduce")
    return dict(sc_proof_outer=sc_outer, claims_outer=(claim_Az, claim_Bz, claim_Cz), eval_E=eval_E,
                sc_proof_inner=sc_inner, eval_W=eval_W, sc_proof_batch=sc_batch, evals_batch=evals_batch,
                r_x=r_x, r_y=r_y, batched_c=batched_c, batched_x=r_b, batched_e=batched_e, batched_poly=batched_poly)


def _rlc_diff(fid, polys, lens, coeffs, n, out):
    k = len(polys)
    ptrs = (ctypes.c_void_p * k)(*[v.ptr.value for v in polys])
    ls = (ctypes.c_size_t * k)(*lens)
    cd = DeviceVec.from_bytes(fields.pack(fid, coeffs))
    check(lib().b200_rlc_dev(fid, ptrs, ls, k, cd.ptr, n, out.ptr, None))
    check(lib().b200_sync())  # `cd` and the pointer table must outlive the launch


def _prove_quad_prod_resident(fid, claim, num_rounds, A: DeviceVec, B: DeviceVec, transcript):
    """SumcheckProof::prove_quad_prod (sumcheck.rs:199-242) on resident vectors, host transcript."""
    from .spartan import SC_QUAD_PROD, UniPoly, _bind_dev, _sc_eval_dev
    p = fields.MODULUS[fid]
    length = 1 << num_rounds
    rs, polys = [], []
    for _ in range(num_rounds):
        e0, bc = _sc_eval_dev(fid, SC_QUAD_PROD, A, B, None, length, None, None, 0)
        poly = UniPoly.from_evals_deg2(p, [e0, (claim - e0) % p, bc])
        transcript.absorb_bytes(b"p", poly.to_transcript_bytes())
        r = transcript.squeeze(b"c")
        rs.append(r)
        polys.append(poly.compress())
        claim = poly.evaluate(r)
        _bind_dev(fid, A, length, r)
        _bind_dev(fid, B, length, r)
        length //= 2
    return polys, rs, fields.unpack(fid, A.to_bytes(32)) + fields.unpack(fid, B.to_bytes(32))


def prove(curve, ck: CommitmentKey, S: dict, U: dict, W: dict, vk_digest: int, transcript, device_transcript: bool = True,
          timings: dict | None = None, ee: str = "hyperkzg"):
    """The whole RelaxedR1CSSNARK::prove (snark.rs:113-256): prove_core, then EE::prove on the batched claim with
    the same transcript; `ck` must be the key the commitments in U were made with.
      ee = "hyperkzg": hyperkzg.rs:926-1116 (primary curve, S1)       -> eval_arg = (com, w, v)
      ee = "ipa":      ipa_pc.rs:64-77, 174-285 (secondary curve, S2) -> eval_arg = (L_vec, R_vec, a_hat);
                       `ck` must carry the generator ck_c as its blinding base."""
    proof = prove_core(curve, ck, S, U, W, vk_digest, transcript, device_transcript, timings)
    if ee == "hyperkzg":
        from .spartan import hyperkzg_prove
        proof["eval_arg"] = hyperkzg_prove(curve, ck, proof["batched_poly"], proof["batched_x"], transcript, timings)
    elif ee == "ipa":
        from .ipa import InnerProductArgument
        fid = Curve(curve).scalar_field
        x = proof["batched_x"]
        b_vec = DeviceVec(32 << len(x))  # EqPolynomial::new(point).evals(), ipa_pc.rs:73
        x_dev = DeviceVec.from_bytes(fields.pack(fid, x))
        check(lib().b200_eq_table_dev(fid, x_dev.ptr, len(x), b_vec.ptr, None))
        proof["eval_arg"] = InnerProductArgument.prove(curve, ck, proof["batched_c"], b_vec, proof["batched_e"],
                                                       proof["batched_poly"], transcript)
    else:
        raise ValueError(f"unknown evaluation engine {ee!r}")
    return proof
