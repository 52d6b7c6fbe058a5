"""Shared body: the device Poseidon RO (nova_b200.poseidon, csrc/poseidon.cuh) against oracle/poseidon_ref.py -- GPU:
tests/test_zz_new_paths_gpu.py; CPU with the emulated device (host mirror only): tests/test_poseidon_mirror_cpu.py."""
from oracle import poseidon_ref as pr
from oracle.pyref import CURVES, FIELD_MODULUS, SplitMix64


def run_ro(nb, fid, arity):
    """squeeze twice (the second hashes the first hash + new elements, poseidon.rs:104-106) for inputs shorter than,
    equal to and longer than the rate (several permutations while absorbing), and the empty input"""
    from nova_b200 import poseidon as dp
    p = FIELD_MODULUS[fid]
    rng = SplitMix64(900 + 10 * fid + arity)
    for n in (0, 1, 9, arity, arity + 1, 2 * arity + 3):
        xs = [rng.field(p) for _ in range(n)]
        dev, ref = dp.PoseidonRO(fid, arity), pr.PoseidonRO(p, arity)
        for e in xs:
            dev.absorb(e)
            ref.absorb(e)
        assert dev.squeeze(128) == ref.squeeze(128), (fid, arity, n)
        assert dev.state == ref.state
        dev.absorb(7)
        ref.absorb(7)
        assert dev.squeeze(250, True) == ref.squeeze(250, True), (fid, arity, n, "second")


def run_nifs_challenge(nb, oracle, cid):
    """The folding challenge of NIFS::prove (nifs.rs:47-63): RO over the curve's BASE field absorbs pp_digest, U2
    (comm_W as (x, y, is_infinity), X) and comm_T, squeezes 128 bits; the scalar-field element the folds use is the same
    integer (base_as_scalar).  Host-driven RO == oracle RO, and the resident form (elements and challenge stay on the
    device, challenge converted into the scalar field there) gives the same Montgomery bytes."""
    from nova_b200 import fields
    from nova_b200 import poseidon as dp
    from nova_b200.spartan import DeviceVec
    c = CURVES[cid]
    base_fid, scalar_fid = c.base_field, c.scalar_field
    pb = FIELD_MODULUS[base_fid]
    rng = SplitMix64(40 + cid)
    G = c.gen
    comm_W, comm_T = c.mul(rng.field(c.q), G), c.mul(rng.field(c.q), G)
    elems = [rng.field(pb), comm_W[0], comm_W[1], 0, rng.field(pb), rng.field(pb), comm_T[0], comm_T[1], 0]
    ref = pr.PoseidonRO(pb)
    dev = dp.PoseidonRO(base_fid)
    for e in elems:
        ref.absorb(e)
        dev.absorb(e)
    r = ref.squeeze(128)
    assert dev.squeeze(128) == r and r < (1 << 128)
    d = DeviceVec.from_bytes(fields.pack(base_fid, elems))
    out = dp.squeeze_dev(base_fid, d, len(elems), 128, out_field=scalar_fid)
    raw = out.to_bytes(96)
    assert raw[32:64] == fields.to_mont_bytes(scalar_fid, r)
    assert int.from_bytes(raw[64:96], "little") == r
    assert fields.unpack(base_fid, raw[:32])[0] == ref.state[0]
