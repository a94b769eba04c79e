"""Pin the oracle's Keccak / Shake256 / ChaCha20 / Merlin restatements."""
import hashlib

import numpy as np

import oracle_lib as ol
from oracle_lib import P, lib, sz


def sha3_256_via_oracle_keccak(msg):
    rate = 136
    st = np.zeros(25, dtype=np.uint64)
    m = bytearray(msg) + b"\x06" + bytes((-len(msg) - 2) % rate) + b"\x80" if (len(msg) + 1) % rate else bytearray(msg) + b"\x86"
    for off in range(0, len(m), rate):
        blk = np.frombuffer(bytes(m[off:off + rate]) + bytes(200 - rate), dtype=np.uint64)
        st ^= blk
        lib().orc_keccak_f1600(P(st))
    return st.tobytes()[:32]


def test_keccak_f_via_sha3():
    for msg in [b"", b"abc", b"a" * 135, b"b" * 136, b"c" * 300]:
        assert sha3_256_via_oracle_keccak(msg) == hashlib.sha3_256(msg).digest()


def test_shake256():
    for msg in [b"", b"gens_sparse_poly", b"x" * 200]:
        out = np.zeros(300, dtype=np.uint8)
        buf = np.frombuffer(msg, dtype=np.uint8).copy() if msg else np.zeros(1, dtype=np.uint8)
        lib().orc_shake256(P(buf), sz(len(msg)), P(out), sz(300))
        assert bytes(out) == hashlib.shake_256(msg).digest(300)


def test_chacha20_zero_key_block0():
    # well-known ChaCha20 keystream for the all-zero key / nonce, block counter 0
    ks = bytes.fromhex("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                       "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
    out = np.zeros(16, dtype=np.uint32)
    lib().orc_chacha20_words(P(np.zeros(32, dtype=np.uint8)), P(out), sz(16))
    assert out.tobytes() == ks


MERLIN_KAT = "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"


def test_merlin_published_vector():
    # merlin's own "equivalence_simple" test: new("test protocol"); append("some label","some data");
    # challenge_bytes("challenge", 32)   
    t = lib().orc_transcript_new(b"test protocol")
    t = ol.C.c_void_p(t)
    data = np.frombuffer(b"some data", dtype=np.uint8).copy()
    lib().orc_transcript_append_message(t, b"some label", P(data), sz(9))
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_transcript_challenge_bytes(t, b"challenge", P(out), sz(32))
    lib().orc_transcript_free(t)
    assert bytes(out).hex() == MERLIN_KAT
