"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares;
no compute call is made here.  Also: the product package never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = []
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        if f.endswith(".h"):
            txt = open(os.path.join(inc, f)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            syms += re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(syms))


def test_library_exports_every_declared_symbol():
    from nova_b200.native import SIGNATURES, STRING_FUNCS, library_path
    assert os.path.exists(library_path()), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(library_path())
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"
        assert s in SIGNATURES or s in STRING_FUNCS, f"{s} has no ctypes signature in nova_b200/native.py"


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nova_b200.native import lib
    L = lib()
    assert L.b200_init(0) != 0
    assert b"no CPU fallback" in L.b200_last_error() or b"CUDA" in L.b200_last_error()
    import nova_b200
    with pytest.raises(nova_b200.B200Error):
        nova_b200.vec_add(0, bytes(32), bytes(32))


def test_product_never_imports_oracle():
    """Nothing under nova_b200/ (sources, headers, Python) may import, include, link or path into
    oracle/.  The reference's own identifiers that merely contain the word (`compute_oracles`,
    `evaluation_oracles`, `mem_oracles`, ppsnark.rs:220-253,365) are allowed in comments/names."""
    import re
    bad = re.compile(r"(from|import)\s+oracle|oracle\s*[/.]|[\"'<]oracle|liboracle|coracle|pyref|_ref/")
    pkg = os.path.join(ROOT, "nova_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read().replace("`oracle/`", "")
                m = bad.search(txt)
                assert m is None or f == "__init__.py", f"{f} reaches into the oracle: {m.group(0)!r}"
