"""Key / ciphertext wire format (include/fhe_b200.h fhe_b200_wire_*; the reference declares none).
Host-only: runs without a GPU."""
import struct

import numpy as np
import pytest

import fhe_b200
import oracle


def _sample(polys=4, limbs=3, n=64, seed=7):
    chain = oracle.prime_chain(limbs)
    rng = np.random.default_rng(seed)
    w = np.stack([np.stack([rng.integers(0, m, n, dtype=np.uint64) for m in chain]) for _ in range(polys)])
    return chain, w


def test_round_trip_and_header():
    chain, w = _sample()
    blob = fhe_b200.wire_pack("switch_key", w, chain, ntt_form=True, galois_elt=127)
    assert len(blob) == 64 + w.size * 8 and blob[:8] == b"FHEB200\0"
    ver, kind, n, limbs, polys, flags, gal, rsv = struct.unpack_from("<8I", blob, 8)
    assert (ver, kind, n, limbs, polys, flags, gal, rsv) == (1, 4, 64, 3, 4, 1, 127, 0)
    assert struct.unpack_from("<Q", blob, 48)[0] == w.size
    assert blob[64:] == w.astype("<u8").tobytes()
    hdr, got = fhe_b200.wire_unpack(blob, chain)
    assert hdr == {"kind": "switch_key", "n": 64, "limbs": 3, "polys": 4, "ntt_form": True, "galois_elt": 127}
    assert np.array_equal(got, w)
    hdr2, got2 = fhe_b200.wire_unpack(blob)          # without the chain check
    assert hdr2 == hdr and np.array_equal(got2, w)


def test_ciphertext_batch_shape():
    chain, w = _sample(polys=6)
    blob = fhe_b200.wire_pack("ciphertext", w.reshape(3, 2, 3, 64), chain)
    hdr, got = fhe_b200.wire_unpack(blob, chain)
    assert hdr["kind"] == "ciphertext" and hdr["polys"] == 6 and not hdr["ntt_form"]
    assert np.array_equal(got.reshape(3, 2, 3, 64), w.reshape(3, 2, 3, 64))


@pytest.mark.parametrize("damage", ["magic", "version", "short", "long", "payload", "header", "chain", "wrap", "more_limbs", "unreduced"])
def test_rejects_damaged_objects(damage):
    chain, w = _sample()
    blob = bytearray(fhe_b200.wire_pack("public_key", w, chain))
    use_chain = chain
    if damage == "magic":
        blob[0] ^= 1
    elif damage == "version":
        struct.pack_into("<I", blob, 8, 2)
    elif damage == "short":
        blob = blob[:-8]
    elif damage == "long":
        blob += b"\0" * 8
    elif damage == "payload":
        blob[64 + 100] ^= 0x40
    elif damage == "header":
        struct.pack_into("<I", blob, 20, 4)            # limbs no longer matches payload_words
    elif damage == "wrap":                             # n * limbs * polys wraps to the stored word count modulo 2^64
        struct.pack_into("<3I", blob, 16, 1 << 31, 1 << 31, 4 * 64 * 3 * 4)
        struct.pack_into("<Q", blob, 48, ((1 << 31) * (1 << 31) * (4 * 64 * 3 * 4)) % (1 << 64))
    elif damage == "chain":
        use_chain = oracle.prime_chain(4)[1:]
    elif damage == "more_limbs":                       # a forged header with more limbs than the caller's chain must be
        big = np.zeros((1, 8, 64), dtype=np.uint64)    # refused before the chain is read (it used to be hashed out of bounds)
        blob = bytearray(fhe_b200.wire_pack("public_key", big, oracle.prime_chain(8)))
        use_chain = chain                              # 3 moduli
    elif damage == "unreduced":                        # a payload word that is not below its limb's modulus
        w2 = w.copy(); w2.reshape(-1)[5] = np.uint64(2**64 - 1)
        blob = bytearray(fhe_b200.wire_pack("public_key", w2, chain))
    with pytest.raises(fhe_b200.FheB200Error):
        fhe_b200.wire_unpack(bytes(blob), use_chain)


def test_pack_rejects_bad_arguments():
    lib = fhe_b200.load_library()
    assert lib.fhe_b200_wire_pack(9, 64, 1, 1, 0, 0, None, None, None) != 0
    assert lib.fhe_b200_wire_unpack(None, 0, None, 0, None, None, None, None, None, None, None) != 0
    assert b"wire" in lib.fhe_b200_last_error()
