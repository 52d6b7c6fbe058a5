"""The C++ host mirror (include/nova_b200.hpp): compiles and links against the C ABI on the CPU
box; on the GPU box runs commits concurrently from 8 threads and checks every result against the
oracle."""
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    hdrs = [os.path.join(ROOT, "include", f) for f in ("nova_b200.hpp", "nova_b200.h")]
    lib = os.path.join(ROOT, "nova_b200", "libnova_b200.so")
    if not os.path.exists(EXE) or any(os.path.getmtime(p) > os.path.getmtime(EXE) for p in [src, lib] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", src, "-o", EXE, "-L" + os.path.dirname(lib),
                               "-lnova_b200", "-Wl,-rpath," + os.path.dirname(lib)])


def test_cpp_mirror_compiles_and_links():
    build()
    out = subprocess.check_output([EXE, "--compile-check"], text=True)
    assert "nova_b200" in out


@pytest.mark.gpu
def test_cpp_mirror_concurrent_commits(oracle, tmp_path):
    from oracle.pyref import CURVES
    build()
    cid, c = 0, CURVES[0]
    n = 6000
    bases = oracle.gen_bases(cid, n + 1)
    sc = oracle.gen_scalars(c.scalar_field, 11, n)
    r = oracle.gen_scalars(c.scalar_field, 12, 1)
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        for blob, sz in ((bases[:64 * n], 64), (bases[64 * n:], 64), (sc, 32), (r, 32)):
            f.write(struct.pack("<Q", len(blob) // sz))
            f.write(blob)
    out = subprocess.run([EXE, str(case)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = open(str(case) + ".out", "rb").read()
    (k,) = struct.unpack_from("<Q", raw, 0)
    pts = [raw[8 + 96 * i:8 + 96 * i + 96] for i in range(k)]
    off = 8 + 96 * k
    aff = lambda b: c.affine_from_bytes(oracle.jacobian_to_affine(cid, b))
    assert aff(pts[0]) == c.affine_from_bytes(oracle.msm_naive(cid, sc + r, bases))
    for j in range(32):  # 8 threads x 4 commits, prefix lengths n / (1 + j % 5)
        ln = n // (1 + j % 5)
        assert aff(pts[1 + j]) == c.affine_from_bytes(oracle.msm(cid, sc[:32 * ln], bases[:64 * ln])), j
    (nf,) = struct.unpack_from("<Q", raw, off)
    folded = raw[off + 8:off + 8 + 32 * nf]
    assert folded == oracle.axpy(c.scalar_field, sc, sc, r)
    off += 8 + 32 * nf
    (nz,) = struct.unpack_from("<Q", raw, off)
    assert raw[off + 8:off + 8 + 32 * nz] == oracle.bind_top(c.scalar_field, sc[:32 * (n & ~1)], r)
