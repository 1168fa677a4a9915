"""The device arithmetic headers (fe25519 / sc25519 / ge25519 .cuh) compiled for the HOST (tests/hostemu, carry chains as
portable C) against Python big integers and the oracle.  A test harness for the library logic, not a product path."""
import ctypes as C
import hashlib
import os
import random
import subprocess

import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "hostemu", "libfe_emu.so")
    src = os.path.join(HERE, "hostemu", "fe_emu.cpp")
    subprocess.check_call(["g++", "-O2", "-DXHE_BOUND_CHECKS", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def _fe(emu, op, a, b=0):
    A = (C.c_uint32 * 8)(*[(a >> (32 * i)) & 0xFFFFFFFF for i in range(8)]); B = (C.c_uint32 * 8)(*[(b >> (32 * i)) & 0xFFFFFFFF for i in range(8)]); O = (C.c_uint32 * 8)()
    emu.emu_fe_op(op, A, B, O)
    return sum(O[i] << (32 * i) for i in range(8))


def test_field_ops(emu):
    rng = random.Random(1)
    edge = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**256 - 1, 2**256 - 38, 2**255, 2**255 - 1, 2**32 - 1, (2**256 - 1) ^ (2**128)]
    vals = edge + [rng.getrandbits(256) for _ in range(150)]
    for a in vals:
        for b in rng.sample(vals, 8) + edge[:6]:
            assert _fe(emu, 0, a, b) % P == (a + b) % P and _fe(emu, 1, a, b) % P == (a - b) % P and _fe(emu, 2, a, b) % P == a * b % P
        assert _fe(emu, 3, a) % P == a * a % P and _fe(emu, 7, a) % P == -a % P and _fe(emu, 4, a) == a % P
    for a in vals[:30]:
        if a % P:
            assert _fe(emu, 5, a) * a % P == 1 and _fe(emu, 6, a) % P == pow(a, (P - 5) // 8, P)


def test_scalar_and_point_ops(emu):
    rnd = hashlib.shake_256(b"emu").digest(64 * 120)

    def scop(op, a, b=bytes(32)):
        o = (C.c_uint8 * 32)(); emu.emu_sc_op(op, a, b, o); return bytes(o)
    for i in range(60):
        a = oracle.sc_reduce_wide(rnd[64 * i:64 * i + 64]); b = oracle.sc_reduce_wide(rnd[64 * (i + 60):64 * (i + 60) + 64])
        for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "inv"), (4, "neg")):
            assert scop(op, a, b) == oracle.sc_op(name, a, b)
        assert scop(5, rnd[64 * i:64 * i + 32], rnd[64 * i + 32:64 * i + 64]) == a
    # edge values against Python integers: extremes of every carry path of the Montgomery reduction
    L = 2**252 + 27742317777372353535851937790883648493
    le = lambda v: v.to_bytes(32, "little")
    edge = [0, 1, 2, L - 1, L - 2, 2**252, 2**252 - 1, 2**128 - 1, (L - 1) ^ (2**125), 2**32 - 1, L >> 1]
    for a in edge:
        for b in edge:
            assert scop(2, le(a), le(b)) == le(a * b % L), (a, b)
        if a:
            assert scop(3, le(a)) == le(pow(a, L - 2, L))
    for lo, hi in ((2**256 - 1, 2**256 - 1), (0, 2**256 - 1), (2**256 - 1, 0), (L, L), (2**255, 2**255)):
        assert scop(5, le(lo), le(hi)) == le((lo + (hi << 256)) % L)
        assert scop(6, le(lo)) == le(lo % L)
    pts = [oracle.from_uniform(rnd[64 * i:64 * i + 64]) for i in range(24)]
    for i in range(0, 24, 2):
        a, b = pts[i], pts[i + 1]; o = (C.c_uint8 * 32)()
        emu.emu_from_uniform(rnd[64 * i:64 * i + 64], o); assert bytes(o) == a
        emu.emu_point_op(0, a, b, o); assert bytes(o) == oracle.point_add(a, b)
        emu.emu_point_op(1, a, b, o); assert bytes(o) == oracle.point_add(a, b, sub=True)
        emu.emu_point_op(2, a, b, o); assert bytes(o) == oracle.point_add(a, a)
        emu.emu_point_op(3, a, b, o); assert bytes(o) == oracle.point_add(oracle.point_add(a, a), b)
        emu.emu_point_op(4, a, b, o); assert bytes(o) == oracle.point_add(oracle.point_add(a, a), b)
    enc = hashlib.shake_256(b"random-encodings").digest(32 * 200) + bytes(32)
    ok = oracle.decode_batch(enc)
    for i in range(201):
        xy = (C.c_uint8 * 64)()
        assert emu.emu_decode(enc[32 * i:32 * i + 32], xy) == ok[i]
