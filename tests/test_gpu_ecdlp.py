"""Decoding of decrypted amounts on the device: ECDLPInstance::decode (reference src/elgamal.rs:67-92) and
ElGamalSecretKey::decrypt (src/elgamal.rs:140-145).  The amount is unique, so parity is exact: the decoded value must be the
plaintext the oracle encrypted, for edge values of the baby-step / giant-step split, out-of-range and invalid inputs."""
import random

import pytest

import oracle

pytestmark = pytest.mark.gpu
L = 2**252 + 27742317777372353535851937790883648493
G = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")      # ristretto255 basepoint encoding (RFC 9496)


@pytest.fixture(scope="module")
def ctx():
    import xelis_he_b200 as xhe
    c = xhe.Ctx(0, party_capacity=2)
    yield c
    c.close()


def _vg(v):
    return oracle.scalarmult((v % L).to_bytes(32, "little"), G)


def test_decode_edge_values_and_ranges(ctx):
    import xelis_he_b200 as xhe
    l1 = 12
    t = xhe.Ecdlp(ctx, l1_bits=l1)
    try:
        assert t.table_bytes == 8 << (l1 + 1)
        rng = random.Random(5)
        stride = 1 << (l1 + 1)
        vals = [0, 1, 2, (1 << l1) - 1, 1 << l1, (1 << l1) + 1, stride - 1, stride, stride + 1, 31 * stride, 32 * stride - 1, 32 * stride, 33 * stride + 7,
                (1 << 24) - 1, (1 << 24) - stride, 123456, 7 * stride + (1 << l1), 7 * stride - (1 << l1)] + [rng.randrange(1 << 24) for _ in range(200)]
        pts = b"".join(_vg(v) for v in vals)
        got, st = t.decode(pts, range_bits=24)
        assert got == vals and st == bytes([1]) * len(vals)
        # out of range: v >= 2^range_bits and "negative" amounts are not found; a narrower range finds only what is inside
        outside = [1 << 24, (1 << 24) + 5, (1 << 30) + 3]
        got, st = t.decode(b"".join(_vg(v) for v in outside) + _vg(L - 1) + _vg(77), range_bits=24)
        assert got == [-1, -1, -1, -1, 77] and st == bytes([0, 0, 0, 0, 1])
        got, st = t.decode(_vg(5000) + _vg(3), range_bits=12)
        assert got == [-1, 3] and st == bytes([0, 1])
        # invalid encodings are reported, not searched
        bad = bytes([1]) + bytes(31)                      # negative s
        got, st = t.decode(bad + _vg(9) + bytes([0xFF]) * 32, range_bits=24)
        assert got == [-1, 9, -1] and st == bytes([2, 1, 2])
        assert t.decode(b"", range_bits=24) == ([], b"")
    finally:
        t.close()


def test_decrypt_and_decode_matches_the_plaintexts(ctx):
    """amounts encrypted by the oracle (pubkey.encrypt, src/elgamal.rs:109-114) come back through C - s * D and the search;
    a ciphertext for another key decrypts to a point that is not a small multiple of G"""
    import xelis_he_b200 as xhe
    t = xhe.Ecdlp(ctx, l1_bits=16)
    try:
        kp = oracle.Keypair.derive(b"ecdlp-wallet"); other = oracle.Keypair.derive(b"ecdlp-other")
        rng = oracle.Rng(b"ecdlp"); r = random.Random(11)
        amounts = [0, 1, 2**32 - 1, 2**31, 10**9 + 7] + [r.randrange(2**32) for _ in range(300)]
        cts = b"".join(kp.encrypt(a, rng) for a in amounts)
        got, st = t.decrypt_decode(kp.sk, cts, range_bits=32)
        assert got == amounts and st == bytes([1]) * len(amounts)
        got, st = t.decrypt_decode(kp.sk, other.encrypt(5, rng) + kp.encrypt(5, rng), range_bits=32)
        assert got == [-1, 5] and st == bytes([0, 1])
        # homomorphic sums decode too (what a wallet does with its balance): ct(a) + ct(b) -> a + b
        a, b = 123456789, 987654321
        s, ok = oracle.ct_update(kp.encrypt(a, rng), kp.encrypt(b, rng), bytes([0]))
        assert ok == bytes([1])
        assert t.decrypt_decode(kp.sk, s, range_bits=32)[0] == [a + b]
    finally:
        t.close()
