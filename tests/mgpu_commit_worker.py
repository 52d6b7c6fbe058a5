"""Worker for the one-process multi-GPU commit (b200_mgpu_*): argv[1] = comma-separated device ids (repeats allowed:
virtual devices on one GPU).  Checks commits of several lengths (ragged against the 4096-point blocks, shorter than
one block round, with and without the blinding term, the empty vector) against the C oracle, on two curves."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    devices = [int(x) for x in sys.argv[1].split(",")]
    import nova_b200 as nb
    from nova_b200.native import check, lib
    from oracle import coracle as co
    from oracle.pyref import CURVES
    check(lib().b200_init(devices[0]))
    ok = True
    for cid, n in ((0, 70001), (2, 20000), (0, 5000)):
        c = CURVES[cid]
        bases = co.gen_bases(cid, n + 1)
        key = nb.MultiGpuCommitmentKey(nb.Curve(cid), bases[:64 * n], bases[64 * n:], devices=devices)
        for m in (n, 4096 * len(devices) + 17, 4097, 1, 0, n - 5):
            if m > n:
                continue
            v = co.gen_scalars(c.scalar_field, 31 + m % 97, m)
            r = co.gen_scalars(c.scalar_field, 7, 1)
            exp = c.affine_from_bytes(co.msm(cid, v, bases[:64 * m]))
            got = key.commit(v)
            ok &= got == exp
            exp_b = c.affine_from_bytes(co.msm_naive(cid, v + r, bases[:64 * m] + bases[64 * n:]))
            ok &= key.commit(v, r) == exp_b
            if not ok:
                print("MISMATCH", cid, n, m)
                break
        key.release()
    print("OK" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
