import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libxhe_cuda.so")
_lib = None

ERR_NAMES = {0: "Ok", 1: "Signature", 2: "Decompression", 3: "CommitmentEqProof", 4: "CiphertextValidityProof", 5: "GenericProof",
             6: "RangeProof", 7: "Transcript", 8: "Format", 9: "InvalidNonce", 10: "State", 11: "Parse",
             -1: "E_ARG", -2: "E_CUDA", -3: "E_NOMEM", -4: "E_NCCL", -5: "E_CAPACITY"}


class XheError(RuntimeError):
    def __init__(self, code, msg=""):
        super().__init__(f"xhe error {code} ({ERR_NAMES.get(code, '?')}) {msg}")
        self.code = code


def lib_path():
    return _LIB_PATH


def load_library():
    """dlopen libxhe_cuda.so; raises loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise XheError(-2, f"{_LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(_LIB_PATH)
    vp, sz, i32, u8p = C.c_void_p, C.c_size_t, C.c_int32, C.c_char_p
    sigs = {
        "xhe_ctx_create": (i32, [C.c_int, C.c_uint32, C.POINTER(vp)]),
        "xhe_ctx_destroy": (None, [vp]),
        "xhe_ctx_party_capacity": (C.c_uint32, [vp]),
        "xhe_last_error": (C.c_char_p, [vp]),
        "xhe_ctx_set_stream": (i32, [vp, vp]),
        "xhe_ctx_sync": (i32, [vp]),
        "xhe_ctx_launch_count": (C.c_uint64, [vp]),
        "xhe_ristretto_decompress": (i32, [vp, u8p, sz, vp, vp]),
        "xhe_ristretto_compress": (i32, [vp, u8p, sz, vp]),
        "xhe_ristretto_from_uniform": (i32, [vp, u8p, sz, vp]),
        "xhe_ct_update": (i32, [vp, u8p, u8p, u8p, sz, vp, vp]),
        "xhe_decompress_dev": (i32, [vp, vp, sz, vp, vp, vp]),
        "xhe_compress_dev": (i32, [vp, vp, sz, vp]),
        "xhe_from_uniform_dev": (i32, [vp, vp, sz, vp]),
        "xhe_ct_update_resident_dev": (i32, [vp, vp, vp, vp, sz]),
        "xhe_ct_update_dev": (i32, [vp, vp, vp, vp, sz, vp, vp]),
        "xhe_msm_workspace_bytes": (sz, [vp, sz]),
        "xhe_msm_dev": (i32, [vp, vp, vp, sz, vp, sz, vp, vp]),
        "xhe_msm_vartime": (i32, [vp, u8p, u8p, sz, vp, C.POINTER(C.c_int32)]),
        "xhe_sum_encodings": (i32, [vp, u8p, sz, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        "xhe_copy_small": (i32, [vp, vp, vp, sz]),
        "xhe_ledger_create": (i32, [vp, sz, C.POINTER(vp)]),
        "xhe_ledger_destroy": (None, [vp]),
        "xhe_ledger_size": (sz, [vp]),
        "xhe_ledger_load": (i32, [vp, u8p, u8p, sz, vp]),
        "xhe_ledger_update": (i32, [vp, u8p, u8p, u8p, sz, vp]),
        "xhe_ledger_update_dense_dev": (i32, [vp, vp, vp]),
        "xhe_ledger_export": (i32, [vp, u8p, sz, vp, vp]),
        "xhe_ledger_device_table": (vp, [vp, C.POINTER(sz)]),
        "xhe_batch_record_dev": (i32, [vp, vp]),
        "xhe_shard_decide_dev": (i32, [vp, vp, C.c_uint32, vp]),
        "xhe_msm_plan": (i32, [sz, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "xhe_measure_int_peak": (i32, [vp, C.c_int, C.POINTER(C.c_double)]),
        "xhe_selftest_fe": (i32, [vp, C.c_int, vp, vp, sz, vp]),
        "xhe_selftest_oct": (i32, [vp, C.c_int, vp, vp, sz, vp]),
        "xhe_ecdlp_create": (i32, [vp, C.c_uint32, C.POINTER(vp)]),
        "xhe_ecdlp_destroy": (None, [vp]),
        "xhe_ecdlp_table_bytes": (sz, [vp]),
        "xhe_ecdlp_decode": (i32, [vp, u8p, sz, C.c_uint32, vp, vp]),
        "xhe_decrypt_decode": (i32, [vp, u8p, u8p, sz, C.c_uint32, vp, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class Ctx:
    """One verifier context per device (xhe_ctx)."""

    def __init__(self, device=0, party_capacity=8):
        self.lib = load_library()
        p = C.c_void_p()
        rc = self.lib.xhe_ctx_create(device, party_capacity, C.byref(p))
        if rc != 0:
            raise XheError(rc, "xhe_ctx_create failed (a CUDA device is required; there is no CPU fallback)")
        self.p = p
        self.stream_ptr = None

    def close(self):
        if getattr(self, "p", None):
            self.lib.xhe_ctx_destroy(self.p)
            self.p = None

    __del__ = close

    def _chk(self, rc):
        if rc < 0:
            raise XheError(rc, self.lib.xhe_last_error(self.p).decode())
        return rc

    def set_stream(self, stream_ptr):
        self._chk(self.lib.xhe_ctx_set_stream(self.p, C.c_void_p(stream_ptr)))
        self.stream_ptr = stream_ptr

    def sync(self):
        self._chk(self.lib.xhe_ctx_sync(self.p))

    @property
    def launches(self):
        return int(self.lib.xhe_ctx_launch_count(self.p))

    # ---- host-buffer API
    def decompress(self, enc: bytes, want_xy=False):
        n = len(enc) // 32
        ok = C.create_string_buffer(max(n, 1))
        xy = C.create_string_buffer(64 * max(n, 1)) if want_xy else None
        self._chk(self.lib.xhe_ristretto_decompress(self.p, enc, n, xy, ok))
        return (ok.raw[:n], xy.raw[:64 * n]) if want_xy else ok.raw[:n]

    def compress(self, xy: bytes):
        n = len(xy) // 64
        enc = C.create_string_buffer(32 * max(n, 1))
        self._chk(self.lib.xhe_ristretto_compress(self.p, xy, n, enc))
        return enc.raw[:32 * n]

    def from_uniform(self, u: bytes):
        n = len(u) // 64
        enc = C.create_string_buffer(32 * max(n, 1))
        self._chk(self.lib.xhe_ristretto_from_uniform(self.p, u, n, enc))
        return enc.raw[:32 * n]

    def ct_update(self, bal: bytes, delta: bytes, sub: bytes):
        n = len(sub)
        out = C.create_string_buffer(64 * max(n, 1))
        ok = C.create_string_buffer(max(n, 1))
        self._chk(self.lib.xhe_ct_update(self.p, bal, delta, sub, n, out, ok))
        return out.raw[:64 * n], ok.raw[:n]

    def msm(self, scalars: bytes, points: bytes):
        """RistrettoPoint::vartime_multiscalar_mul + is_identity (src/proofs.rs:49-67): returns (encoding, is_identity)."""
        n = len(scalars) // 32
        out = C.create_string_buffer(32)
        ident = C.c_int32(0)
        rc = self.lib.xhe_msm_vartime(self.p, scalars, points, n, out, C.byref(ident))
        if rc != 0:
            raise XheError(rc, self.lib.xhe_last_error(self.p).decode())
        return out.raw, bool(ident.value)

    def sum_encodings(self, encodings: bytes):
        """Sum of canonical Ristretto encodings (per-rank partial MSM results): returns (encoding of the sum, is_identity and
        all inputs valid).  One small kernel, no allocation -- safe to call between batches in flight."""
        n = len(encodings) // 32
        out = C.create_string_buffer(32)
        ident, valid = C.c_int32(0), C.c_int32(0)
        self._chk(self.lib.xhe_sum_encodings(self.p, encodings, n, out, C.byref(ident), C.byref(valid)))
        return out.raw, bool(ident.value) and bool(valid.value)

    def copy_small(self, dst_ptr, src_ptr, nbytes):
        """kernel-based copy on this context's stream between device-accessible addresses (device or pinned host memory)"""
        self._chk(self.lib.xhe_copy_small(self.p, dst_ptr, src_ptr, nbytes))

    def msm_plan(self, n):
        c, w = C.c_int(), C.c_int()
        self.lib.xhe_msm_plan(n, C.byref(c), C.byref(w))
        return c.value, w.value

    def int_peak(self, which):
        r = C.c_double()
        self._chk(self.lib.xhe_measure_int_peak(self.p, which, C.byref(r)))
        return r.value

    def selftest_fe(self, op, a_words, b_words):
        import numpy as np
        a = np.ascontiguousarray(a_words, dtype=np.uint32)
        b = np.ascontiguousarray(b_words, dtype=np.uint32)
        n = a.size // 8
        out = np.zeros_like(a)
        self._chk(self.lib.xhe_selftest_fe(self.p, op, a.ctypes.data, b.ctypes.data, n, out.ctypes.data))
        return out


    def selftest_oct(self, op, a_words, b_words):
        """csrc/oct.cuh on the device: op 0 mul, 1 add, 2 sub (rows of 8 words); 3 point doubling, 4 point addition (rows of 32 words)"""
        import numpy as np
        a = np.ascontiguousarray(a_words, dtype=np.uint32)
        b = np.ascontiguousarray(b_words, dtype=np.uint32)
        n = a.size // (8 if op < 3 else 32)
        out = np.zeros_like(a)
        self._chk(self.lib.xhe_selftest_oct(self.p, op, a.ctypes.data, b.ctypes.data, n, out.ctypes.data))
        return out


class Ecdlp:
    """Decoding of decrypted amounts on the device (include/xhe.h, xhe_ecdlp_*; reference ECDLPInstance::decode,
    src/elgamal.rs:67-92): a baby-step table of 2^l1_bits entries, built once."""

    def __init__(self, ctx, l1_bits=22):
        self.ctx, self.lib = ctx, ctx.lib
        p = C.c_void_p()
        ctx._chk(self.lib.xhe_ecdlp_create(ctx.p, l1_bits, C.byref(p)))
        self.p, self.l1_bits = p, l1_bits

    @property
    def table_bytes(self):
        return int(self.lib.xhe_ecdlp_table_bytes(self.p))

    def decode(self, points: bytes, range_bits=32):
        """points: n x 32-byte encodings of M = v * G -> (values: list of int, -1 where not found; status bytes: 1 found, 0 not in range, 2 invalid encoding)"""
        n = len(points) // 32
        vals = (C.c_int64 * max(n, 1))(); st = C.create_string_buffer(max(n, 1))
        self.ctx._chk(self.lib.xhe_ecdlp_decode(self.p, points, n, range_bits, vals, st))
        return list(vals[:n]), st.raw[:n]

    def decrypt_decode(self, secret_key: bytes, ciphertexts: bytes, range_bits=32):
        """ElGamalSecretKey::decrypt (M = C - s * D, src/elgamal.rs:140-145) followed by decode; ciphertexts: n x 64 bytes"""
        n = len(ciphertexts) // 64
        vals = (C.c_int64 * max(n, 1))(); st = C.create_string_buffer(max(n, 1))
        self.ctx._chk(self.lib.xhe_decrypt_decode(self.p, secret_key, ciphertexts, n, range_bits, vals, st))
        return list(vals[:n]), st.raw[:n]

    def close(self):
        if getattr(self, "p", None):
            self.lib.xhe_ecdlp_destroy(self.p); self.p = None

    __del__ = close


class DeviceLedger:
    """Device-resident ledger (include/xhe.h, xhe_ledger_*): balances stay decompressed in HBM; keys are account || asset."""

    def __init__(self, ctx, capacity):
        self.ctx, self.lib = ctx, ctx.lib
        p = C.c_void_p()
        ctx._chk(self.lib.xhe_ledger_create(ctx.p, capacity, C.byref(p)))
        self.p = p

    def close(self):
        if getattr(self, "p", None):
            self.lib.xhe_ledger_destroy(self.p); self.p = None

    __del__ = close

    def __len__(self):
        return int(self.lib.xhe_ledger_size(self.p))

    def load(self, records):
        """records: iterable of (pk, asset, ct64); returns per-record ok flags"""
        records = list(records)
        keys = b"".join(pk + asset for pk, asset, _ in records); cts = b"".join(ct for _, _, ct in records)
        ok = C.create_string_buffer(max(len(records), 1))
        self.ctx._chk(self.lib.xhe_ledger_load(self.p, keys, cts, len(records), ok))
        return ok.raw[:len(records)]

    def update(self, keys: bytes, deltas: bytes, sub: bytes):
        n = len(sub)
        st = C.create_string_buffer(max(n, 1))
        self.ctx._chk(self.lib.xhe_ledger_update(self.p, keys, deltas, sub, n, st))
        return st.raw[:n]

    def update_dense_dev(self, d_delta_niels_planar, d_sub):
        self.ctx._chk(self.lib.xhe_ledger_update_dense_dev(self.p, C.c_void_p(d_delta_niels_planar), C.c_void_p(d_sub)))

    def export(self, keys: bytes):
        n = len(keys) // 64
        out = C.create_string_buffer(64 * max(n, 1)); found = C.create_string_buffer(max(n, 1))
        self.ctx._chk(self.lib.xhe_ledger_export(self.p, keys, n, out, found))
        return out.raw[:64 * n], found.raw[:n]
