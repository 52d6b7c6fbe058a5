"""Shared body of the folding-step parity check (GPU: tests/test_zz_new_paths_gpu.py; CPU with the emulated
device: tests/test_r1cs_mirror_cpu.py).  The reference's own fixture -- the tiny cubic R1CS
x^3 + x + 5 = y folded twice into the default running instance, then `is_sat_relaxed`, then a relaxed +
relaxed fold (src/nova/nifs.rs:299-351, 429-501; src/r1cs/mod.rs:1349-1413) -- runs through
nova_b200.r1cs (commit_T, NIFS orchestration, witness / instance folds, is_sat_relaxed) with blinded
commitments, and every folded vector is compared with the oracle's fold of the same inputs."""
from oracle.pyref import CURVES, SplitMix64, mont_bytes
from test_oracle_nifs_fold import NUM_CONS, NUM_IO, NUM_VARS, Oracle, csr, tiny_r1cs, witness


def pack(p, xs):
    return b"".join(mont_bytes(p, x) for x in xs)


def run_tiny_fixture(nb, oracle, cid):
    from nova_b200 import r1cs, spartan as sp
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    o = Oracle(oracle, fid)
    rng = SplitMix64(70 + cid)
    ncols = NUM_VARS + 1 + NUM_IO
    mats = [sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols) for (d, idx, ptr) in (csr(M, NUM_CONS) for M in tiny_r1cs())]
    S = r1cs.R1CSShape(nb.Curve(cid), *mats, NUM_CONS, NUM_VARS, NUM_IO)
    bases = oracle.gen_bases(cid, max(NUM_CONS, NUM_VARS) + 1)
    n_key = max(NUM_CONS, NUM_VARS)
    ck = nb.CommitmentKey(nb.Curve(cid), bases[:64 * n_key], bases[64 * n_key:])

    def commit_ref(v, r):
        return c.affine_from_bytes(oracle.msm(cid, pack(p, list(v) + [r]), bases[:64 * len(v)] + bases[64 * n_key:]))

    def fresh(x):  # a satisfying (R1CSInstance, R1CSWitness) with a blinded commitment
        Wv, Xv = witness(p, x)
        r_W = rng.field(p)
        return (r1cs.R1CSInstance(commit_ref(Wv, r_W), Xv), r1cs.R1CSWitness(sp.DeviceVec.from_bytes(pack(p, Wv)), r_W),
                Wv, Xv)

    def vec(dv, n):
        return o.ints(dv.to_bytes(32 * n))

    U = r1cs.RelaxedR1CSInstance.default(NUM_IO)
    W = r1cs.RelaxedR1CSWitness.default(NUM_VARS, NUM_CONS)
    run = ([0] * NUM_VARS, [0] * NUM_CONS, 0, [0] * NUM_IO)  # the oracle's copy of the running pair
    assert S.is_sat_relaxed(ck, U, W)
    for x in (rng.field(p), 3):
        U2, W2, Wv, Xv = fresh(x)
        r_T, r = rng.field(p), rng.field(p)
        comm_T, (U, W) = r1cs.nifs_prove(ck, S, U, W, U2, W2, r_T, lambda cT: r)
        run = o.fold(*run, Wv, Xv, r)
        assert (vec(W.W, NUM_VARS), vec(W.E, NUM_CONS), U.u, U.X) == (run[0], run[1], run[2], run[3])
        assert S.is_sat_relaxed(ck, U, W)
    # a wrong error vector, a wrong blind and a wrong commitment are all rejected
    bad_E = list(run[1])
    bad_E[0] = (bad_E[0] + 1) % p
    assert not S.is_sat_relaxed(ck, U, r1cs.RelaxedR1CSWitness(W.W, sp.DeviceVec.from_bytes(pack(p, bad_E)), W.r_W, W.r_E))
    assert not S.is_sat_relaxed(ck, U, r1cs.RelaxedR1CSWitness(W.W, W.E, (W.r_W + 1) % p, W.r_E))
    assert not S.is_sat_relaxed(ck, r1cs.RelaxedR1CSInstance(U.comm_E, U.comm_E, U.X, U.u), W)
    # relaxed + relaxed (NIFSRelaxed::prove, commit_T_relaxed): a second running pair folded into the first
    U_b = r1cs.RelaxedR1CSInstance.default(NUM_IO)
    W_b = r1cs.RelaxedR1CSWitness.default(NUM_VARS, NUM_CONS)
    other = ([0] * NUM_VARS, [0] * NUM_CONS, 0, [0] * NUM_IO)
    for x in (7, 11):
        U2, W2, Wv, Xv = fresh(x)
        r = rng.field(p)
        _, (U_b, W_b) = r1cs.nifs_prove(ck, S, U_b, W_b, U2, W2, rng.field(p), lambda cT: r)
        other = o.fold(*other, Wv, Xv, r)
    r = rng.field(p)
    _, (U_c, W_c) = r1cs.nifs_prove(ck, S, U, W, U_b, W_b, rng.field(p), lambda cT: r)
    both = o.fold(*run, other[0], other[3], r, E2=other[1], u2=other[2])
    assert (vec(W_c.W, NUM_VARS), vec(W_c.E, NUM_CONS), U_c.u, U_c.X) == (both[0], both[1], both[2], both[3])
    assert S.is_sat_relaxed(ck, U_c, W_c)
    ck.release()


def run_streamed_steps(nb, oracle, cid):
    """Three prove_step-like folds where the fresh witness reaches the device through ONE WitnessStream (chunks of
    1 and 2 scalars, re-armed with reset() every step) and is folded from its resident copy."""
    from nova_b200 import r1cs, spartan as sp
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    o = Oracle(oracle, fid)
    rng = SplitMix64(170 + cid)
    ncols = NUM_VARS + 1 + NUM_IO
    mats = [sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols) for (d, idx, ptr) in (csr(M, NUM_CONS) for M in tiny_r1cs())]
    S = r1cs.R1CSShape(nb.Curve(cid), *mats, NUM_CONS, NUM_VARS, NUM_IO)
    n_key = max(NUM_CONS, NUM_VARS)
    bases = oracle.gen_bases(cid, n_key + 1)
    ck = nb.CommitmentKey(nb.Curve(cid), bases[:64 * n_key], bases[64 * n_key:])
    U = r1cs.RelaxedR1CSInstance.default(NUM_IO)
    W = r1cs.RelaxedR1CSWitness.default(NUM_VARS, NUM_CONS)
    run = ([0] * NUM_VARS, [0] * NUM_CONS, 0, [0] * NUM_IO)
    ws = nb.WitnessStream(ck, NUM_VARS)
    for step, x in enumerate((rng.field(p), 3, 11)):
        Wv, Xv = witness(p, x)
        r_W, r_T, r = (0 if step == 1 else rng.field(p)), rng.field(p), rng.field(p)
        chunks = [pack(p, Wv[:1]), pack(p, Wv[1:])]
        U2, comm_T, (U, W) = r1cs.fold_streamed_step(ck, S, U, W, ws, chunks, Xv, r_W, r_T, lambda cT: r)
        exp_comm = c.affine_from_bytes(oracle.msm(cid, pack(p, Wv + [r_W]), bases[:64 * NUM_VARS] + bases[64 * n_key:]))
        assert U2.comm_W == exp_comm
        run = o.fold(*run, Wv, Xv, r)
        assert (o.ints(W.W.to_bytes(32 * NUM_VARS)), o.ints(W.E.to_bytes(32 * NUM_CONS)), U.u, U.X) == \
               (run[0], run[1], run[2], run[3])
        assert S.is_sat_relaxed(ck, U, W)
        ws.reset()
    ws.release()
    ck.release()


def run_compressed_half(nb, oracle, cid, num_cons=16, num_vars=8, num_io=2, device_transcript=False):
    """The per-curve half of CompressedSNARK::prove (nova/mod.rs:813-881) at toy size: a satisfied running pair
    is folded with a freshly sampled random pair (sample_random_instance_witness + NIFSRelaxed::prove),
    derandomized, and handed to spartan::snark::prove; the restated verifier must accept the proof for the
    derandomized instance."""
    from nova_b200 import r1cs, snark as ds, spartan as sp
    from oracle import snark_ref as sr
    from oracle.ppsnark_ref import random_instance
    from oracle.pyref import Keccak256Transcript
    from snark_parity import csr as csr_rows
    c = CURVES[cid]
    fid, p = c.scalar_field, c.q
    rng = SplitMix64(900 + cid + num_cons)
    Sd, Wd, u, X = random_instance(p, rng, num_cons, num_vars, num_io)
    ncols = num_vars + 1 + num_io
    mats = []
    for name in "ABC":
        d, idx, ptr = csr_rows(Sd[name], num_cons)
        mats.append(sp.SparseMatrix(fid, pack(p, d), idx, ptr, ncols))
    S = r1cs.R1CSShape(nb.Curve(cid), *mats, num_cons, num_vars, num_io)
    n_key = max(num_cons, num_vars)
    bases = oracle.gen_bases(cid, n_key + 1)
    ck = nb.CommitmentKey(nb.Curve(cid), bases[:64 * n_key], bases[64 * n_key:])
    r_W, r_E = rng.field(p), rng.field(p)
    W_run = r1cs.RelaxedR1CSWitness(sp.DeviceVec.from_bytes(pack(p, Wd["W"])), sp.DeviceVec.from_bytes(pack(p, Wd["E"])), r_W, r_E)
    U_run = r1cs.RelaxedR1CSInstance(S._commit(ck, W_run.W, num_vars, r_W), S._commit(ck, W_run.E, num_cons, r_E), X, u)
    assert S.is_sat_relaxed(ck, U_run, W_run)
    Z = pack(p, [rng.field(p) for _ in range(ncols)])
    U_rand, W_rand = r1cs.sample_random_instance_witness(ck, S, Z, rng.field(p), rng.field(p))
    assert S.is_sat_relaxed(ck, U_rand, W_rand)
    r = rng.field(p)
    _, (U_n, W_n) = r1cs.nifs_prove(ck, S, U_run, W_run, U_rand, W_rand, rng.field(p), lambda cT: r)
    assert S.is_sat_relaxed(ck, U_n, W_n)
    U_d, W_d, _, _ = r1cs.derandomize(ck, nb.Curve(cid), U_n, W_n)
    assert S.is_sat_relaxed(ck, U_d, W_d)  # commitments now open with zero blinds
    tr = Keccak256Transcript(p, b"RelaxedR1CSSNARK")
    Ssn = dict(num_cons=num_cons, num_vars=num_vars, A=mats[0], B=mats[1], C=mats[2])
    Uc = dict(comm_W=U_d.comm_W, comm_E=U_d.comm_E, u=U_d.u, X=U_d.X)
    proof = ds.prove_core(nb.Curve(cid), None, Ssn, Uc, dict(W=W_d.W, E=W_d.E), 99, tr, device_transcript=device_transcript)
    joint = sr.verify_core(p, c, Sd, Uc, 99, proof)  # raises if either sum-check or the batch claim is wrong
    assert joint == (proof["batched_c"], proof["batched_x"], proof["batched_e"])
    ck.release()
