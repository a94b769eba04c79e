"""Pin the oracle's Fr/Fq Montgomery arithmetic (ark-ff restatement) against Python big-ints."""
import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import L_FR, Q_FQ, P, f_op, from_mont, int_to_limbs, lib, limbs_to_int, to_mont


@pytest.mark.parametrize("which,p", [(0, L_FR), (1, Q_FQ)])
def test_mont_ops_random(which, p):
    rng = np.random.default_rng(7 + which)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, 2**64 - 1, 2**128 + 5, 2**252, p >> 1]
    vals = edge + [int.from_bytes(rng.bytes(40), "little") % p for _ in range(200)]
    for i in range(len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        A, B = to_mont(a, p), to_mont(b, p)
        assert from_mont(f_op(which, 0, A, B), p) == (a + b) % p
        assert from_mont(f_op(which, 1, A, B), p) == (a - b) % p
        assert from_mont(f_op(which, 2, A, B), p) == (a * b) % p
        assert from_mont(f_op(which, 4, A), p) == (-a) % p
        # outputs are canonical residues (< p) in Montgomery form
        assert limbs_to_int(f_op(which, 2, A, B)) < p
        if a:
            assert from_mont(f_op(which, 3, A), p) == pow(a, -1, p)


@pytest.mark.parametrize("which,p", [(0, L_FR), (1, Q_FQ)])
def test_conversions(which, p):
    out = np.zeros(4, dtype=np.uint64)
    for v in [0, 1, 28, 2**63, 2**64 - 1]:
        lib().orc_f_from_u64(which, ol.C.c_uint64(v), P(out))
        assert from_mont(out, p) == v
        assert limbs_to_int(out) == v * 2**256 % p  # ark-ff layout: a * R mod p
        can = np.zeros(4, dtype=np.uint64)
        lib().orc_f_to_canonical(which, P(out), P(can))
        assert limbs_to_int(can) == v
        back = np.zeros(4, dtype=np.uint64)
        lib().orc_f_from_canonical(which, P(can), P(back))
        assert (back == out).all()


def test_from_le_bytes_mod_order_64():
    rng = np.random.default_rng(3)
    cases = [bytes(64), b"\xff" * 64, (1).to_bytes(64, "little")] + [rng.bytes(64) for _ in range(50)]
    for b in cases:
        out = np.zeros(4, dtype=np.uint64)
        buf = np.frombuffer(b, dtype=np.uint8).copy()
        lib().orc_fr_from_le_bytes_mod_order_64(P(buf), P(out))
        assert from_mont(out) == int.from_bytes(b, "little") % L_FR


def test_constants():
    # SURVEY Appendix C constants, recomputed
    assert int_to_limbs(2**256 % L_FR).tolist() == [0xD6EC31748D98951D, 0xC6EF5BF4737DCF70, 0xFFFFFFFFFFFFFFFE, 0x0FFFFFFFFFFFFFFF]
    assert (-pow(L_FR, -1, 2**64)) % 2**64 == 0xD2B51DA312547E1B
    assert 2**256 % Q_FQ == 38 and (-pow(Q_FQ, -1, 2**64)) % 2**64 == 0x86BCA1AF286BCA1B
