"""nova_b200/ptau.py (PTAU file -> device-resident key, src/provider/ptau.rs) on the CPU: the file logic and the
G2 checks are host code; the G1 validation calls go to the emulated device.  GPU twin:
tests/test_zz_new_paths_gpu.py."""
import gc

import pytest

import emulated_device
import ptau_parity


@pytest.fixture()
def ptau():
    from nova_b200 import ptau as mod
    emulated_device.install()
    yield mod
    gc.collect()
    emulated_device.uninstall()


def test_g2_host_arithmetic(ptau):
    """the constants and the group law the G2 checks rest on: generator on the curve and of order r, [r+1]G = G,
    (a + b)G = aG + bG, raw round trip, identity"""
    G = ptau.G2_GENERATOR
    r = ptau._R_ORDER
    assert ptau.g2_on_curve(G) and ptau.g2_is_torsion_free(G)
    assert ptau.g2_mul(G, r + 1) == G and ptau.g2_mul(G, r - 1) == (G[0], tuple((-c) % ptau._Q for c in G[1]))
    a, b = 0x1234567890ABCDEF, 0xFEDCBA987654321
    assert ptau._g2_add(ptau.g2_mul(G, a), ptau.g2_mul(G, b)) == ptau.g2_mul(G, a + b)
    assert ptau.g2_from_raw(ptau.g2_to_raw(G)) == G and ptau.g2_from_raw(bytes(128)) is None
    assert ptau.g2_on_curve(None) and ptau.g2_is_torsion_free(None)
    # b' = 3 / (9 + u)
    assert ptau._f2_mul(ptau._B2, (9, 1)) == (3, 0)
    P = ptau_parity.non_subgroup_g2(ptau)
    assert ptau.g2_on_curve(P) and not ptau.g2_is_torsion_free(P)


def test_reference_read_ptau_cases(ptau):
    ptau_parity.run_reference_cases(ptau)


def test_format_errors(ptau):
    ptau_parity.run_format_errors(ptau)


def test_save_load_setup_round_trip(ptau, oracle, tmp_path):
    ptau_parity.run_setup_round_trip(ptau, oracle, tmp_path)


def test_checked_registration(ptau, oracle):
    ptau_parity.run_checked_registration(oracle, 1 << 9, 300, 450)


@pytest.mark.parametrize("cid", [1, 2, 3])
def test_pedersen_key_file(ptau, oracle, cid):
    ptau_parity.run_pedersen_key_file(ptau, oracle, cid)
