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
    lib.xo_mint_chain_par.restype = C.c_void_p
    lib.xo_mint_mixed.restype = C.c_void_p
    lib.xo_ledger_dump_multisig.restype = C.c_size_t
    lib.xo_batch_slice.restype = C.c_void_p
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

    def dump_multisig(self):
        """[(pk, [signers], threshold)] of every account with a multisig setting"""
        n = lib.xo_ledger_dump_multisig(self.ptr, None, C.c_size_t(0))
        out = (C.c_uint8 * max(n, 1))()
        lib.xo_ledger_dump_multisig(self.ptr, out, C.c_size_t(n))
        raw, o, res = bytes(out)[:n], 0, []
        while o < len(raw):
            k, th = raw[o + 32], raw[o + 33]
            res.append((raw[o:o + 32], [raw[o + 34 + 32 * i:o + 66 + 32 * i] for i in range(k)], th)); o += 34 + 32 * k
        return res

    def record_outputs(self, on=True):
        """keep (compressed) what set_output_ciphertext receives; off by default like the reference mock, which drops it"""
        lib.xo_ledger_record_outputs(self.ptr, 1 if on else 0)
        return self

    def dump_outputs(self):
        """(pk, asset, compressed output ciphertext) of the last set_output_ciphertext call per key (src/tx/verify.rs:339-340)."""
        lib.xo_ledger_dump_outputs.restype = C.c_size_t
        n = lib.xo_ledger_dump_outputs(self.ptr, None, C.c_size_t(0))
        out = (C.c_uint8 * (128 * max(n, 1)))()
        lib.xo_ledger_dump_outputs(self.ptr, out, C.c_size_t(128 * n))
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

    def slice(self, count):
        return Batch(lib.xo_batch_slice(self.ptr, C.c_size_t(count)))

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


def mint_chain(seed, T, k=1, threads=1):
    """one sender, T transactions chained through its balance (benches/tx.rs:153-186); threads > 1 mints in two parallel passes"""
    if threads > 1:
        return Batch(lib.xo_mint_chain_par(C.c_uint64(seed), C.c_size_t(T), k, threads))
    return Batch(lib.xo_mint_chain(C.c_uint64(seed), C.c_size_t(T), k))


def minted_keypair(seed, i, which="s"):
    """Keypair of sender ("s") / receiver ("r") i of a batch minted with `seed` (mint_transfers, mint_mixed)"""
    out = (C.c_uint8 * 32)()
    lib.xo_mint_secret(C.c_uint64(seed), ord(which), C.c_uint64(i), out)
    return Keypair(bytes(out))


def mint_mixed(seed, T, threads=8):
    """config 5: transfers (k 1..4, a 1..2), burns, contract calls, multisig set-ups and transfers from threshold-2 multisig accounts"""
    return Batch(lib.xo_mint_mixed(C.c_uint64(seed), C.c_size_t(T), threads))


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


def verify_batch_partial(blobs, ledger: Ledger, rng_seed=1):
    """one shard of a multi-GPU batch: local verdict plus the partial sigma / range MSM encodings (no identity decision)."""
    n = len(blobs)
    arr = (C.c_char_p * max(n, 1))(*blobs) if n else (C.c_char_p * 1)()
    lens = (C.c_size_t * max(n, 1))(*[len(b) for b in blobs])
    rng = (C.c_uint8 * 512)()
    seed = (C.c_uint64 * 1)(rng_seed)
    lib.xo_rng_init(rng, seed, C.c_size_t(8))
    fi = C.c_long(-1)
    part = (C.c_uint8 * 64)()
    rc = lib.xo_verify_batch_ex(arr, lens, C.c_size_t(n), ledger.ptr, rng, C.byref(fi), part)
    return rc, fi.value, bytes(part[:32]), bytes(part[32:])


def apply_without_verify(blob, ledger: Ledger):
    return lib.xo_apply_without_verify(_buf(blob), C.c_size_t(len(blob)), ledger.ptr)


# ---------------------------------------------------------------------------------------------------------------------
# general transaction builder binding (xo_tx_build) -- used to replay the reference's test scenarios (src/lib.rs:254-1093)
# ---------------------------------------------------------------------------------------------------------------------
class _TransferSpec(C.Structure):
    _fields_ = [("asset", C.c_uint8 * 32), ("dest", C.c_uint8 * 32), ("amount", C.c_uint64), ("extra", C.c_char_p), ("extra_len", C.c_uint32), ("has_extra", C.c_int)]


class _TxSpec(C.Structure):
    _fields_ = [("version", C.c_uint8), ("type", C.c_uint8), ("source", C.c_uint8 * 32), ("fee", C.c_uint64), ("nonce", C.c_uint64),
                ("transfers", C.POINTER(_TransferSpec)), ("n_transfers", C.c_uint32),
                ("burn_asset", C.c_uint8 * 32), ("burn_amount", C.c_uint64),
                ("contract", C.c_uint8 * 32), ("call_assets", C.c_char_p), ("call_amounts", C.POINTER(C.c_uint64)), ("n_call_assets", C.c_uint32),
                ("raw_tail", C.c_char_p), ("raw_tail_len", C.c_uint32), ("n_params", C.c_uint32),
                ("signers", C.c_char_p), ("n_signers", C.c_uint32), ("threshold", C.c_uint8),
                ("assets", C.c_char_p), ("balances", C.POINTER(C.c_uint64)), ("n_assets", C.c_uint32)]


NATIVE = bytes(32)
TRANSFERS, BURN, CALL, DEPLOY, MULTISIG = 0, 1, 2, 3, 4


class Rng:
    def __init__(self, seed: bytes):
        self.buf = (C.c_uint8 * 512)()
        lib.xo_rng_init(self.buf, seed, C.c_size_t(len(seed)))

    def scalar(self):
        out = (C.c_uint8 * 32)()
        lib.xo_rng_scalar(self.buf, out)
        return bytes(out)


class Keypair:
    """ElGamalKeypair (src/elgamal.rs:155-215): sk scalar bytes, pk compressed."""

    def __init__(self, sk: bytes):
        self.sk = sk
        pk = (C.c_uint8 * 32)()
        lib.xo_pubkey_from_secret(_buf(sk), pk, None)
        self.pk = bytes(pk)

    @staticmethod
    def derive(tag: bytes):
        import hashlib
        return Keypair(sc_reduce_wide(hashlib.shake_256(b"xhe-test-key" + tag).digest(64)))

    def encrypt(self, amount: int, rng: "Rng"):
        """pubkey.encrypt(amount) -> compressed ciphertext (src/elgamal.rs:109-114)."""
        ge = (C.c_uint8 * 160)()   # struct ge: 4 x fe (5 x u64)
        pk = (C.c_uint8 * 32)()
        lib.xo_pubkey_from_secret(_buf(self.sk), pk, ge)
        ct = (C.c_uint8 * 64)()
        lib.xo_encrypt(ct, ge, C.c_uint64(amount), _buf(rng.scalar()))
        return bytes(ct)


def build_tx(kp: Keypair, ledger: "Ledger", rng: Rng, *, fee=0, nonce=0, version=1, transfers=None, burn=None, call=None, deploy=None,
             multisig_setup=None, balances, cosigners=None):
    """TransactionBuilder::build (src/tx/builder.rs:547-554).  `balances` = [(asset, plaintext balance)] in commitment order
    (native first).  transfers = [(asset, dest_pk, amount[, extra bytes])]; burn = (asset, amount); call = (contract, [(asset, amount)],
    [(key, value)]); deploy = code bytes; multisig_setup = ([signer pks], threshold); cosigners = [(index, Keypair)]."""
    sp = _TxSpec()
    sp.version, sp.fee, sp.nonce = version, fee, nonce
    keep = []
    if transfers is not None:
        sp.type = TRANSFERS
        arr = (_TransferSpec * max(len(transfers), 1))()
        for i, t in enumerate(transfers):
            arr[i].asset[:] = t[0]; arr[i].dest[:] = t[1]; arr[i].amount = t[2]
            if len(t) > 3 and t[3] is not None:
                arr[i].extra = t[3]; arr[i].extra_len = len(t[3]); arr[i].has_extra = 1
        sp.transfers = arr; sp.n_transfers = len(transfers); keep.append(arr)
    elif burn is not None:
        sp.type = BURN; sp.burn_asset[:] = burn[0]; sp.burn_amount = burn[1]
    elif call is not None:
        sp.type = CALL; sp.contract[:] = call[0]
        assets = b"".join(a for a, _ in call[1]); amts = (C.c_uint64 * max(len(call[1]), 1))(*[v for _, v in call[1]])
        tail = b"".join(len(k).to_bytes(4, "little") + k + len(v).to_bytes(4, "little") + v for k, v in call[2])
        sp.call_assets = assets; sp.call_amounts = amts; sp.n_call_assets = len(call[1]); sp.raw_tail = tail; sp.raw_tail_len = len(tail); sp.n_params = len(call[2])
        keep += [assets, amts, tail]
    elif deploy is not None:
        sp.type = DEPLOY; sp.raw_tail = deploy; sp.raw_tail_len = len(deploy)
    elif multisig_setup is not None:
        sp.type = MULTISIG; s = b"".join(multisig_setup[0]); sp.signers = s; sp.n_signers = len(multisig_setup[0]); sp.threshold = multisig_setup[1]; keep.append(s)
    a = b"".join(x for x, _ in balances); bal = (C.c_uint64 * len(balances))(*[v for _, v in balances])
    sp.assets = a; sp.balances = bal; sp.n_assets = len(balances)
    out = C.POINTER(C.c_uint8)()
    n_ms = len(cosigners) if cosigners else 0
    ms_idx = bytes(i for i, _ in cosigners) if cosigners else None
    ms_sk = b"".join(k.sk for _, k in cosigners) if cosigners else None
    n = lib.xo_tx_build(C.byref(out), C.byref(sp), _buf(kp.sk), ledger.ptr, rng.buf, ms_idx, ms_sk, n_ms)
    if n == 0:
        raise RuntimeError("xo_tx_build failed (insufficient funds / bad input)")
    return C.string_at(out, n)


def tx_to_bytes(blob: bytes):
    """Transaction::to_bytes (src/tx/verify.rs:623-688) -> (bytes, multisig_index)."""
    view = (C.c_uint8 * 256)()
    keep = _buf(blob)   # the parsed view points into this buffer
    if lib.xo_tx_parse(view, keep, C.c_size_t(len(blob))) != 0:
        return None
    out = C.POINTER(C.c_uint8)()
    msi = C.c_size_t(0)
    lib.xo_tx_to_bytes.restype = C.c_size_t
    n = lib.xo_tx_to_bytes(view, C.byref(out), C.byref(msi))
    lib.xo_tx_free(view)
    return C.string_at(out, n), msi.value


def resign(blob: bytes, kp: Keypair, rng: Rng):
    """Sign again after a mutation, so the tampering survives the signature check and reaches the proof checks."""
    b = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    if not lib.xo_resign(b, C.c_size_t(len(blob)), _buf(kp.sk), rng.buf):
        return None
    return bytes(b)
