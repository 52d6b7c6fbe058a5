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


def test_variant_libraries_are_not_stale():
    """A/B builds of the library (nova_b200/libnova_b200_<variant>.so, selected with NOVA_B200_LIB) that are lying
    around must export the current ABI: a variant built before an entry point was added would fail at load time
    on the GPU box."""
    import glob
    for path in glob.glob(os.path.join(ROOT, "nova_b200", "libnova_b200_*.so")):
        L = ctypes.CDLL(path)
        missing = [s for s in declared_symbols() if not hasattr(L, s)]
        assert not missing, f"{os.path.basename(path)} is stale (rebuild: make -C nova_b200/csrc variant ...): {missing}"


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


def test_struct_layouts_match_the_python_binding(tmp_path):
    """The structs that cross the ABI by pointer (b200_transcript, b200_sc_state, b200_scb_desc, b200_scb_state):
    sizes and field offsets as a C compiler sees include/nova_b200.h == what nova_b200's ctypes code assumes."""
    import ctypes
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nova_b200.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b200_transcript), '
                   'offsetof(b200_transcript, state), sizeof(b200_sc_state), offsetof(b200_sc_state, round), '
                   'offsetof(b200_sc_state, tstate), sizeof(b200_scb_desc), offsetof(b200_scb_desc, tau), '
                   'sizeof(b200_scb_state), offsetof(b200_scb_state, coeff), offsetof(b200_scb_state, q));return 0;}\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    from nova_b200.ppsnark import ScbDesc
    assert got[0] == 72 and got[1] == 8            # spartan._device_loop packs round (8) + state (64)
    assert got[2] == 144 and got[3] == 64 and got[4] == 72   # head of the state the mirrors read back
    assert got[5] == ctypes.sizeof(ScbDesc) and got[6] == ScbDesc.tau.offset
    assert got[7] == 1296 and got[8] == 144 and got[9] == 144 + 512 + 512
