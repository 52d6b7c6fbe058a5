#!/usr/bin/env python3
"""Kernel-level replay of RecursiveSNARK::prove_step (src/nova/mod.rs:456-564) on a
synthetic shape with the MinRoot circuit's size and sparsity (BASELINE.json configs[2]):

  NIFS::prove(secondary, Grumpkin ~10.5k)  Z1+Z2 -> 3 SpMV -> T -> commit(T) -> fold W, E
  commit(W_primary)                        BN254 MSM over num_vars
  NIFS::prove(primary, BN254 ~2.07e5)      Z1+Z2 -> 3 SpMV -> T -> commit(T) -> fold W, E
  commit(W_secondary)                      Grumpkin MSM

= 4 MSMs, 6 SpMVs, 2 cross-terms, 4 folds, 2 vector adds per step.  Circuit synthesis, Poseidon
RO and control flow stay on the host in the real prover and are NOT part of this number.
The GPU side keeps shapes, keys and the running (W, E) resident; the fresh step witness W2 is
uploaded from pinned host memory every step and the 4 commitments are read back.
`gpu_replay()` is called by bench.py; `cpu_replay()` is its cpu_baseline leg (the only part that
touches oracle/).
"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

PRIMARY_CONS = 3 * 65536 + 9986   # minroot 65536 iterations/step + augmented circuit
SECONDARY_CONS = 10538            # src/nova/circuit/mod.rs:452-456
NUM_IO = 2
K0 = 0x5EED


def synth_vec(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 60) - 1)
    return a


def synth_shape(rows, cols, seed):
    """CSR with 1-2 unit coefficients per row (MinRoot-like: x*x = y style constraints)."""
    rng = np.random.default_rng(seed)
    per = rng.integers(1, 3, size=rows)
    indptr = np.zeros(rows + 1, dtype=np.uint64)
    np.cumsum(per, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = rng.integers(0, cols, size=nnz, dtype=np.uint64)
    return nnz, indices, indptr


def one_mont(fid):
    R = 1 << 256
    p = {0: 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001,
         1: 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47}[fid]
    return (R % p).to_bytes(32, "little")


class Side:
    """One curve's half of the step: shape (A,B,C), key, resident running instance."""

    def __init__(self, curve, fid, cons, seed):
        self.curve, self.fid, self.cons = curve, fid, cons
        self.vars = cons
        self.zlen = self.vars + 1 + NUM_IO
        self.shapes = [synth_shape(cons, self.zlen, seed + k) for k in range(3)]
        self.W1 = synth_vec(self.vars, seed + 10)
        self.E1 = synth_vec(cons, seed + 11)
        self.W2 = synth_vec(self.vars, seed + 12)
        self.X1 = synth_vec(1 + NUM_IO, seed + 13)
        self.X2 = synth_vec(1 + NUM_IO, seed + 14)
        self.u = synth_vec(1, seed + 15)
        self.r = synth_vec(1, seed + 16)


def make_sides():
    return Side(1, 1, SECONDARY_CONS, 100), Side(0, 0, PRIMARY_CONS, 200)  # (Grumpkin/Fq, BN254/Fr)


def gpu_replay(steps=5, warmup=2, return_outputs=False, overlap=True):
    """overlap: commit(W2) runs on a second stream beside the SpMV / cross-term chain of the same side (both only need the
    uploaded witness; `b200_msm_dev` is asynchronous on the caller's stream) -- what a host that calls the `_dev` entry
    points from two streams gets.  False = everything on one stream (the round-1 replay)."""
    import torch

    import nova_b200 as nb
    from nova_b200.native import check, lib
    L = lib()
    check(L.b200_init(0))
    stream, stream2 = torch.cuda.Stream(), torch.cuda.Stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    sp2 = ctypes.c_void_p(stream2.cuda_stream) if overlap else sp
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    sides = make_sides()
    st = []
    for s in sides:
        d = {}
        d["ck"] = nb.CommitmentKey.setup_synthetic(nb.Curve(s.curve), max(s.cons, s.vars), k0=K0)
        hs = []
        for nnz, idx, ptr in s.shapes:
            data = one_mont(s.fid) * nnz
            h = ctypes.c_uint64(0)
            ia = idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
            ip = ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
            check(L.b200_spmv_register(s.fid, ctypes.create_string_buffer(data, len(data)), ia, ip, s.cons, s.zlen,
                                       ctypes.byref(h)))
            hs.append(h.value)
        d["mats"] = hs
        d["Z1"] = dev(np.concatenate([s.W1, s.X1]))
        d["E1"] = dev(s.E1)
        d["Z2"] = torch.zeros(s.zlen * 32, dtype=torch.uint8, device="cuda")
        d["Z2"][s.vars * 32:] = dev(s.X2)
        d["W2_host"] = torch.from_numpy(s.W2.view(np.uint8).reshape(-1)).pin_memory()
        d["Z"] = torch.zeros(s.zlen * 32, dtype=torch.uint8, device="cuda")
        d["az"], d["bz"], d["cz"], d["T"] = (torch.zeros(s.cons * 32, dtype=torch.uint8, device="cuda") for _ in range(4))
        d["Wf"] = torch.zeros(s.vars * 32, dtype=torch.uint8, device="cuda")
        d["Ef"] = torch.zeros(s.cons * 32, dtype=torch.uint8, device="cuda")
        d["u"], d["r"] = dev(s.u), dev(s.r)
        d["comm"] = torch.zeros(96 * 2, dtype=torch.uint8, device="cuda")
        d["comm_host"] = torch.empty(96 * 2, dtype=torch.uint8).pin_memory()
        st.append(d)

    def nifs_and_commit(s, d):
        # fresh witness of this step arrives from the host (frontend/r1cs.rs:40-50)
        d["Z2"][:s.vars * 32].copy_(d["W2_host"], non_blocking=True)
        if overlap:
            stream2.wait_stream(stream)                                                          # the upload
        check(L.b200_msm_dev(d["ck"].handle, 0, P(d["Z2"]), s.vars, P(d["comm"]), sp2))         # commit(W2)
        check(L.b200_vec_add_dev(s.fid, P(d["Z1"]), P(d["Z2"]), s.zlen, P(d["Z"]), sp))         # Z1 + Z2
        for h, o in zip(d["mats"], (d["az"], d["bz"], d["cz"])):
            check(L.b200_spmv_dev(h, P(d["Z"]), None, P(o), None, sp))                           # 3 SpMV
        check(L.b200_cross_term_dev(s.fid, P(d["az"]), P(d["bz"]), P(d["cz"]), P(d["E1"]), None, P(d["u"]),
                                    s.cons, P(d["T"]), sp))                                      # T
        check(L.b200_msm_dev(d["ck"].handle, 0, P(d["T"]), s.cons, ctypes.c_void_p(d["comm"].data_ptr() + 96), sp))
        if overlap:
            stream.wait_stream(stream2)                                                          # comm_W is written there
        d["comm_host"].copy_(d["comm"], non_blocking=True)                                       # comm_W, comm_T -> host
        stream.synchronize()                       # the RO challenge r is derived from comm_T on the host
        check(L.b200_axpy_dev(s.fid, P(d["Z1"]), P(d["Z2"]), P(d["r"]), s.vars, P(d["Wf"]), sp)) # W fold
        check(L.b200_axpy_dev(s.fid, P(d["E1"]), P(d["T"]), P(d["r"]), s.cons, P(d["Ef"]), sp))  # E fold

    def step():
        for s, d in zip(sides, st):
            nifs_and_commit(s, d)
        stream.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
    res = {"ms_per_step": ms, "primary_constraints": PRIMARY_CONS, "secondary_constraints": SECONDARY_CONS,
           "ops_per_step": "4 MSM + 6 SpMV + 2 cross-term + 4 fold + 2 vec-add, fresh W uploaded, 4 commitments read back",
           "excluded": "circuit synthesis, Poseidon RO, control flow (host side of the real prover)",
           "h2d_bytes_per_step": sum(32 * s.vars for s in sides), "d2h_bytes_per_step": 2 * 192}
    if return_outputs:  # what the step produced, for the parity check against cpu_replay (tools/workloads.py)
        import hashlib

        from nova_b200.provider import Curve, _jac_to_affine
        outs = {}
        for s, d in zip(sides, st):
            raw = bytes(d["comm"].cpu().numpy().tobytes())
            outs[f"curve{s.curve}"] = {
                "comm_W2": str(_jac_to_affine(Curve(s.curve), raw[:96])), "comm_T": str(_jac_to_affine(Curve(s.curve), raw[96:])),
                "W_fold_sha256": hashlib.sha256(d["Wf"].cpu().numpy().tobytes()).hexdigest(),
                "E_fold_sha256": hashlib.sha256(d["Ef"].cpu().numpy().tobytes()).hexdigest()}
        res["outputs"] = outs
    return res


def cpu_replay(steps=1, return_outputs=False):
    from oracle import coracle as co
    cores = os.cpu_count() or 1
    try:  # honour a cgroup CPU quota (the GPU boxes expose 128 CPUs with a 16-CPU quota)
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            cores = max(1, min(cores, int(int(quota) / int(period))))
    except Exception:
        pass
    L = co.lib()
    sides = make_sides()
    prep = []
    for s in sides:
        bases = co.gen_bases(s.curve, max(s.cons, s.vars), K0)
        mats = []
        for nnz, idx, ptr in s.shapes:
            mats.append((one_mont(s.fid) * nnz, idx, ptr))
        prep.append((bases, mats))
    B = lambda b: ctypes.create_string_buffer(bytes(b), len(b))

    results = {}

    def step():
        for s, (bases, mats) in zip(sides, prep):
            W2, W1, E1 = s.W2.tobytes(), s.W1.tobytes(), s.E1.tobytes()
            Z1 = W1 + s.X1.tobytes()
            Z2 = W2 + s.X2.tobytes()
            cw = co.msm(s.curve, W2, bases[:64 * s.vars], cores)
            Z = ctypes.create_string_buffer(len(Z1))
            L.orc_vec_par(s.fid, 2, B(Z1), B(Z2), None, None, None, None, ctypes.c_size_t(s.zlen), Z, cores)
            outs = []
            for data, idx, ptr in mats:
                o = ctypes.create_string_buffer(32 * s.cons)
                L.orc_spmv_par(s.fid, B(data), idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                               ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), ctypes.c_size_t(s.cons), Z, o, cores)
                outs.append(o)
            T = ctypes.create_string_buffer(32 * s.cons)
            L.orc_vec_par(s.fid, 0, outs[0], outs[1], outs[2], B(E1), None, B(s.u.tobytes()), ctypes.c_size_t(s.cons), T, cores)
            ct = co.msm(s.curve, T.raw, bases[:64 * s.cons], cores)
            Wf = ctypes.create_string_buffer(32 * s.vars)
            L.orc_vec_par(s.fid, 1, B(W1), B(W2), None, None, None, B(s.r.tobytes()), ctypes.c_size_t(s.vars), Wf, cores)
            Ef = ctypes.create_string_buffer(32 * s.cons)
            L.orc_vec_par(s.fid, 1, B(E1), T, None, None, None, B(s.r.tobytes()), ctypes.c_size_t(s.cons), Ef, cores)
            if return_outputs:
                import hashlib

                from oracle.pyref import CURVES
                aff = CURVES[s.curve].affine_from_bytes
                results[f"curve{s.curve}"] = {"comm_W2": str(aff(cw)), "comm_T": str(aff(ct)),
                                           "W_fold_sha256": hashlib.sha256(Wf.raw).hexdigest(),
                                           "E_fold_sha256": hashlib.sha256(Ef.raw).hexdigest()}

    step()  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return {"ms_per_step": (time.perf_counter() - t0) * 1e3 / steps, "cores": cores, "kind": "port",
            **({"outputs": results} if return_outputs else {}),
            "note": "same op sequence through the C restatement (threaded MSM / SpMV / vector kernels); includes "
                    "Python buffer copies of ~60 MB per step"}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "cpu":
        print(cpu_replay())
    elif len(sys.argv) > 1 and sys.argv[1] == "ab":  # one stream vs commit(W2) beside the SpMV chain
        for ov in (False, True, False, True):
            print("overlap", ov, round(gpu_replay(steps=10, warmup=3, overlap=ov)["ms_per_step"], 4))
    else:
        print(gpu_replay())
