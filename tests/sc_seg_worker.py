"""Subprocess body of test_segmented_eq_reduction_on_device (the switch NOVA_B200_SC_SEG is read once per process):
every eq-weighted sum-check form at a size where the segmented kernel is taken (shift >= 10), through
b200_sc_eval, against the C oracle; with the switch off the same script checks the default kernel."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes

    from nova_b200.native import check, lib
    from oracle import coracle as co
    check(lib().b200_init(0))
    ok = True
    for fid in (0, 3):
        for form in (4, 5, 6, 7, 8, 9, 10):
            for count, shift, nleft in (((1 << 14) + 777, 10, 17), (1 << 15, 12, 8)):
                length = count if form == 10 else 2 * count
                A, B, C = (co.gen_scalars(fid, 10 * form + k, length) for k in range(3))
                left, right = co.gen_scalars(fid, 91, nleft), co.gen_scalars(fid, 92, 1 << shift)
                out = ctypes.create_string_buffer(96)
                buf = lambda b: ctypes.create_string_buffer(b, len(b))
                check(lib().b200_sc_eval(fid, form, buf(A), buf(B), buf(C), length, buf(left), nleft, buf(right),
                                         1 << shift, shift, out))
                exp = co.sc_eval(fid, form, A, B, C, left, right, shift)
                ok &= out.raw[:len(exp)] == exp
    print("SEG" if os.environ.get("NOVA_B200_SC_SEG") == "1" else "FLAT", "OK" if ok else "MISMATCH")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
