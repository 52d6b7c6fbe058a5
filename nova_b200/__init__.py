"""nova_b200 -- B200-native (sm_100a) implementation of Nova's prover hot path.

The product is the C-ABI shared library `libnova_b200.so` (include/nova_b200.h).  This Python
package is the host-side mirror of the reference's provider interface for that path
(`DlogGroupExt`, `CommitmentEngineTrait`, `R1CSShape`, `MultilinearPolynomial` ...), used by
the tests, the benchmark and as executable documentation of the binding a Rust
`provider/b200.rs` would contain.  It never imports `oracle/` and has no
CPU fallback: if the CUDA library is missing or no GPU is present, calls raise.
"""
from .native import B200Error, lib, library_path  # noqa: F401
from . import fields  # noqa: F401
from .provider import (  # noqa: F401
    CommitmentEngine,
    CommitmentKey,
    Curve,
    DlogGroup,
    MultiGpuCommitmentKey,
    WitnessStream,
    bind_poly_var_top,
    cross_term,
    fold_witness,
    vec_add,
)
