"""oracle -- CPU restatement of the reference's batch-verification path.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this package;
the product (xelis_he_b200) never does.  The C sources beside this file cite the reference file:line they follow.
PARITY STATUS: encodings / group law / scalar field are pinned to libsodium and RFC 9496 vectors, Merlin/STROBE and
BLAKE3 to public KATs (tests/test_oracle_*.py); the Bulletproofs boundary is "parity unpinned" (see bp.h).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


def _load():
    if not os.path.exists(_LIB):
        build()
    lib = C.CDLL(_LIB)
    lib.xo_batch_verify_timed.restype = C.c_double
    lib.xo_msm_timed.restype = C.c_double
    lib.xo_mint_transfers.restype = C.c_void_p
    lib.xo_mint_chain.restype = C.c_void_p
    lib.xo_ledger_new.restype = C.c_void_p
    lib.xo_ledger_clone.restype = C.c_void_p
    lib.xo_ledger_dump.restype = C.c_size_t
    lib.xo_tx_build.restype = C.c_size_t
    return lib


lib = _load()


class _Batch(C.Structure):
    _fields_ = [("blobs", C.POINTER(C.c_uint8)), ("offsets", C.POINTER(C.c_size_t)), ("n", C.c_size_t), ("ledger", C.c_void_p)]


ERR = {0: "Ok", 1: "Signature", 2: "Decompression", 3: "CommitmentEqProof", 4: "CiphertextValidityProof", 5: "GenericProof",
       6: "RangeProof", 7: "Transcript", 8: "Format", 9: "InvalidNonce", 10: "State", 11: "Parse"}


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(b)


def decode_batch(enc: bytes, want_xy=False):
    n = len(enc) // 32
    ok = (C.c_uint8 * n)()
    xy = (C.c_uint8 * (64 * n))() if want_xy else None
    lib.xo_decode_batch(_buf(enc), C.c_size_t(n), ok, xy)
    return (bytes(ok), bytes(xy)) if want_xy else bytes(ok)


def point_add(a: bytes, b: bytes, sub=False):
    out = (C.c_uint8 * 32)()
    return bytes(out) if lib.xo_point_add(_buf(a), _buf(b), int(sub), out) else None


def scalarmult(s: bytes, p: bytes):
    out = (C.c_uint8 * 32)()
    return bytes(out) if lib.xo_scalarmult(_buf(s), _buf(p), out) else None


def from_uniform(u: bytes):
    out = (C.c_uint8 * 32)()
    lib.xo_from_uniform(_buf(u), out)
    return bytes(out)


def sc_reduce_wide(b: bytes):
    out = (C.c_uint8 * 32)()
    lib.xo_sc_reduce_wide(_buf(b), out)
    return bytes(out)


def sc_op(op: str, a: bytes, b: bytes = bytes(32)):
    out = (C.c_uint8 * 32)()
    lib.xo_sc_op({"add": 0, "sub": 1, "mul": 2, "inv": 3, "neg": 4}[op], _buf(a), _buf(b), out)
    return bytes(out)


def msm(scalars: bytes, points: bytes, mode="dalek"):
    n = len(scalars) // 32
    out = (C.c_uint8 * 32)()
    ok = lib.xo_msm(_buf(scalars), _buf(points), C.c_size_t(n), {"dalek": 0, "straus": 1, "pippenger": 2, "naive": 3}[mode], out)
    return bytes(out) if ok else None


def msm_timed(scalars: bytes, points: bytes):
    n = len(scalars) // 32
    out = (C.c_uint8 * 32)()
    t = lib.xo_msm_timed(_buf(scalars), _buf(points), C.c_size_t(n), out)
    return t, bytes(out)


def ct_update(bal: bytes, delta: bytes, sub: bytes):
    n = len(sub)
    out = (C.c_uint8 * (64 * n))()
    ok = (C.c_uint8 * n)()
    lib.xo_ct_update(_buf(bal), _buf(delta), _buf(sub), C.c_size_t(n), out, ok)
    return bytes(out), bytes(ok)


def gen_msm_inputs(seed: int, n: int, threads=8, points=True):
    s = (C.c_uint8 * (32 * n))()
    p = (C.c_uint8 * (32 * n))() if points else None
    lib.xo_gen_msm_inputs(C.c_uint64(seed), C.c_size_t(n), s, p, threads)
    return bytes(s), (bytes(p) if points else None)


def msm_expected_known_bases(scalars: bytes, idx, base_scalars: bytes):
    import numpy as np
    idx = np.ascontiguousarray(idx, dtype=np.uint32)
    out = (C.c_uint8 * 32)()
    lib.xo_msm_expected_known_bases(_buf(scalars), idx.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(len(idx)), _buf(base_scalars), out)
    return bytes(out)


class Ledger:
    """mock::Ledger of the reference (src/lib.rs:106-201)."""

    def __init__(self, ptr=None):
        self.ptr = C.c_void_p(ptr if ptr is not None else lib.xo_ledger_new())

    def clone(self):
        return Ledger(lib.xo_ledger_clone(self.ptr))

    def set_balance(self, pk, asset, ct):
        lib.xo_ledger_set_balance(self.ptr, _buf(pk), _buf(asset), _buf(ct))

    def get_balance(self, pk, asset):
        out = (C.c_uint8 * 64)()
        return bytes(out) if lib.xo_ledger_get_balance(self.ptr, _buf(pk), _buf(asset), out) else None

    def set_nonce(self, pk, nonce):
        lib.xo_ledger_set_nonce(self.ptr, _buf(pk), C.c_uint64(nonce))

    def set_multisig(self, pk, signers, threshold):
        lib.xo_ledger_set_multisig(self.ptr, _buf(pk), _buf(b"".join(signers)) if signers else None, len(signers), C.c_uint8(threshold))

    def dump(self):
        n = lib.xo_ledger_dump(self.ptr, None, C.c_size_t(0))
        out = (C.c_uint8 * (128 * max(n, 1)))()
        lib.xo_ledger_dump(self.ptr, out, C.c_size_t(128 * n))
        raw = bytes(out)[: 128 * n]
        return [(raw[i:i + 32], raw[i + 32:i + 64], raw[i + 64:i + 128]) for i in range(0, len(raw), 128)]

    def __del__(self):
        try:
            lib.xo_ledger_free(self.ptr)
        except Exception:
            pass


class Batch:
    """A minted batch: list of xtx1 blobs + the initial ledger."""

    def __init__(self, ptr):
        if not ptr:
            raise RuntimeError("minting failed")
        self.ptr = C.c_void_p(ptr)
        b = C.cast(self.ptr, C.POINTER(_Batch)).contents
        self.n = b.n
        offs = [b.offsets[i] for i in range(b.n + 1)]
        raw = C.string_at(b.blobs, offs[-1])
        self.blobs = [raw[offs[i]:offs[i + 1]] for i in range(b.n)]
        self._ledger_ptr = b.ledger

    def ledger(self):
        return Ledger(lib.xo_ledger_clone(C.c_void_p(self._ledger_ptr)))

    def verify_timed(self, threads=1):
        rc = C.c_int(0)
        t = lib.xo_batch_verify_timed(self.ptr, threads, C.byref(rc))
        return t, rc.value

    def __del__(self):
        try:
            lib.xo_batch_free(self.ptr)
        except Exception:
            pass


def mint_transfers(seed, T, a=1, k=1, threads=8):
    return Batch(lib.xo_mint_transfers(C.c_uint64(seed), C.c_size_t(T), a, k, threads))


def mint_chain(seed, T, k=1):
    return Batch(lib.xo_mint_chain(C.c_uint64(seed), C.c_size_t(T), k))


def verify_batch(blobs, ledger: Ledger, rng_seed=1):
    """Transaction::verify_batch (src/tx/verify.rs:487-517). Returns (code, first_failing_index)."""
    n = len(blobs)
    arr = (C.c_char_p * max(n, 1))(*blobs) if n else (C.c_char_p * 1)()
    lens = (C.c_size_t * max(n, 1))(*[len(b) for b in blobs])
    rng = (C.c_uint8 * 512)()
    seed = (C.c_uint64 * 1)(rng_seed)
    lib.xo_rng_init(rng, seed, C.c_size_t(8))
    fi = C.c_long(-1)
    rc = lib.xo_verify_batch(arr, lens, C.c_size_t(n), ledger.ptr, rng, C.byref(fi))
    return rc, fi.value


def apply_without_verify(blob, ledger: Ledger):
    return lib.xo_apply_without_verify(_buf(blob), C.c_size_t(len(blob)), ledger.ptr)
