"""Python face of the host layer (xelis_he_b200/host/verifier.*): the reference's verification API over the C library.

    reference                                     here
    mock::Ledger              src/lib.rs:106-201     Ledger
    Transaction::verify_batch src/tx/verify.rs:487   verify_batch(ctx, txs, ledger)
    Transaction::verify       src/tx/verify.rs:520   verify(ctx, tx, ledger)   (same verdicts, appendix D.8)
    apply_without_verify      src/tx/verify.rs:545   apply_without_verify(ctx, txs, ledger)
"""
import ctypes as C

from ._lib import ERR_NAMES, XheError, load_library


def _lib():
    lib = load_library()
    if getattr(lib, "_xheh_ready", False):
        return lib
    vp, sz = C.c_void_p, C.c_size_t
    lib.xheh_ledger_new.restype = vp
    lib.xheh_ledger_clone.restype = vp; lib.xheh_ledger_clone.argtypes = [vp]
    lib.xheh_ledger_free.argtypes = [vp]
    lib.xheh_ledger_set_balance.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_char_p]
    lib.xheh_ledger_get_balance.argtypes = [vp, C.c_char_p, C.c_char_p, vp]
    lib.xheh_ledger_set_nonce.argtypes = [vp, C.c_char_p, C.c_uint64]
    lib.xheh_ledger_set_multisig.argtypes = [vp, C.c_char_p, C.c_char_p, sz, C.c_uint8]
    lib.xheh_ledger_has_multisig.argtypes = [vp, C.c_char_p]
    lib.xheh_ledger_size.restype = sz; lib.xheh_ledger_size.argtypes = [vp]
    lib.xheh_ledger_import.argtypes = [vp, C.c_char_p, sz]
    lib.xheh_ledger_export.restype = sz; lib.xheh_ledger_export.argtypes = [vp, vp, sz]
    lib.xheh_ledger_record_outputs.argtypes = [vp, C.c_int]
    lib.xheh_ledger_outputs_size.restype = sz; lib.xheh_ledger_outputs_size.argtypes = [vp]
    lib.xheh_ledger_export_outputs.restype = sz; lib.xheh_ledger_export_outputs.argtypes = [vp, vp, sz]
    lib.xheh_verify_batch.restype = C.c_int32
    lib.xheh_verify_batch.argtypes = [vp, vp, vp, vp, sz, C.c_char_p, sz, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    lib.xheh_verify_batch_partial.restype = C.c_int32
    lib.xheh_verify_batch_partial.argtypes = [vp, vp, vp, vp, sz, C.c_char_p, sz, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_double), vp]
    lib.xheh_verify_batch_ex.restype = C.c_int32
    lib.xheh_verify_batch_ex.argtypes = [vp, vp, vp, vp, sz, C.c_char_p, sz, C.c_int, C.c_uint32, C.POINTER(C.c_long), C.POINTER(C.c_double), vp]
    lib.xheh_verify_batch_shard.restype = C.c_int32
    lib.xheh_verify_batch_shard.argtypes = [vp, vp, vp, vp, sz, sz, sz, C.c_char_p, sz, C.c_int, C.c_uint32, C.POINTER(C.c_long), C.POINTER(C.c_double), vp]
    lib.xheh_export_taken.restype = sz; lib.xheh_export_taken.argtypes = [vp, vp, sz]
    lib.xheh_ledger_apply_records.restype = C.c_int32; lib.xheh_ledger_apply_records.argtypes = [vp, C.c_char_p, sz]
    lib.xheh_commit_pending.restype = C.c_int32; lib.xheh_commit_pending.argtypes = [vp, vp]
    lib.xheh_apply_without_verify.restype = C.c_int32
    lib.xheh_apply_without_verify.argtypes = [vp, vp, vp, vp, sz]
    lib._xheh_ready = True
    return lib


_MODE_FLAGS = {"host": 0, "device": 1, "fast": 5}
_DETERMINISTIC = 8      # tests only: batch factors from the seed alone (a run can be replayed); otherwise OS entropy is always folded in


class Ledger:
    def __init__(self, ptr=None):
        self.lib = _lib()
        self.ptr = C.c_void_p(ptr if ptr is not None else self.lib.xheh_ledger_new())

    def clone(self):
        return Ledger(self.lib.xheh_ledger_clone(self.ptr))

    def set_balance(self, pk, asset, ct):
        self.lib.xheh_ledger_set_balance(self.ptr, pk, asset, ct)

    def get_balance(self, pk, asset):
        out = C.create_string_buffer(64)
        return out.raw if self.lib.xheh_ledger_get_balance(self.ptr, pk, asset, out) else None

    def set_nonce(self, pk, nonce):
        self.lib.xheh_ledger_set_nonce(self.ptr, pk, nonce)

    def set_multisig(self, pk, signers, threshold):
        self.lib.xheh_ledger_set_multisig(self.ptr, pk, b"".join(signers), len(signers), threshold)

    def has_multisig(self, pk):
        return bool(self.lib.xheh_ledger_has_multisig(self.ptr, pk))

    def import_records(self, records):
        """records: iterable of (pk, asset, ct64); accounts get nonce 0."""
        blob = b"".join(pk + asset + ct for pk, asset, ct in records)
        self.lib.xheh_ledger_import(self.ptr, blob, len(blob) // 128)

    def apply_records(self, records: bytes):
        """update_account_balance for each 128-byte record (account, asset, new ciphertext), in order -- the balance updates a
        peer rank exported with export_taken()"""
        return self.lib.xheh_ledger_apply_records(self.ptr, records, len(records) // 128)

    def dump(self):
        n = self.lib.xheh_ledger_size(self.ptr)
        buf = C.create_string_buffer(128 * max(n, 1))
        self.lib.xheh_ledger_export(self.ptr, buf, 128 * n)
        raw = buf.raw[:128 * n]
        return sorted((raw[i:i + 32], raw[i + 32:i + 64], raw[i + 64:i + 128]) for i in range(0, len(raw), 128))

    def record_outputs(self, on=True):
        """keep the ciphertext of every set_output_ciphertext call (src/tx/verify.rs:339-340, 582); the reference mock drops
        them, and a state that does not ask for them saves the device two encodings per (tx, asset)"""
        self.lib.xheh_ledger_record_outputs(self.ptr, 1 if on else 0)

    def dump_outputs(self):
        n = self.lib.xheh_ledger_outputs_size(self.ptr)
        buf = C.create_string_buffer(128 * max(n, 1))
        self.lib.xheh_ledger_export_outputs(self.ptr, buf, 128 * n)
        raw = buf.raw[:128 * n]
        return sorted((raw[i:i + 32], raw[i + 32:i + 64], raw[i + 64:i + 128]) for i in range(0, len(raw), 128))

    def __del__(self):
        try:
            self.lib.xheh_ledger_free(self.ptr)
        except Exception:
            pass


class DeviceLedgerState:
    """BlockchainVerificationState over a device-resident ledger: pass it wherever a Ledger is accepted.  The
    fast path reads balances from the device table and commits accepted updates there; dump() compresses on demand."""

    def __init__(self, ctx, capacity):
        self.lib = _lib()
        lib = self.lib
        lib.xheh_dledger_new.restype = C.c_void_p; lib.xheh_dledger_new.argtypes = [C.c_void_p, C.c_size_t]
        lib.xheh_dledger_free.argtypes = [C.c_void_p]
        lib.xheh_dledger_import.restype = C.c_int32; lib.xheh_dledger_import.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        lib.xheh_dledger_export.restype = C.c_size_t; lib.xheh_dledger_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        lib.xheh_dledger_set_multisig.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_uint8]
        lib.xheh_dledger_snapshot.restype = C.c_int32; lib.xheh_dledger_snapshot.argtypes = [C.c_void_p]
        lib.xheh_dledger_restore.restype = C.c_int32; lib.xheh_dledger_restore.argtypes = [C.c_void_p]
        p = lib.xheh_dledger_new(ctx.p, capacity)
        if not p:
            raise XheError(-3, "device ledger allocation failed")
        self.ptr = C.c_void_p(p)
        self.ctx = ctx

    def import_records(self, records):
        blob = b"".join(pk + asset + ct for pk, asset, ct in records)
        rc = self.lib.xheh_dledger_import(self.ptr, blob, len(blob) // 128)
        if rc:
            raise XheError(rc, "device ledger import failed")

    def set_multisig(self, pk, signers, threshold):
        self.lib.xheh_dledger_set_multisig(self.ptr, pk, b"".join(signers), len(signers), threshold)

    def dump(self):
        n = self.lib.xheh_dledger_export(self.ptr, None, 0)
        buf = C.create_string_buffer(128 * max(n, 1))
        self.lib.xheh_dledger_export(self.ptr, buf, 128 * n)
        raw = buf.raw[:128 * n]
        return sorted((raw[i:i + 32], raw[i + 32:i + 64], raw[i + 64:i + 128]) for i in range(0, len(raw), 128))

    def snapshot(self):
        return self.lib.xheh_dledger_snapshot(self.ptr)

    def restore(self):
        return self.lib.xheh_dledger_restore(self.ptr)

    def close(self):
        if getattr(self, "ptr", None):
            self.lib.xheh_dledger_free(self.ptr); self.ptr = None

    __del__ = close


def _attach_index(bl, lib, threads=0):
    """key-digest index of the batch (a few 8-byte words per transaction), built where the transactions are framed: a
    shard-mode call finds the earlier transactions its shard depends on without reading the other shards' bytes"""
    import time
    lib.xheh_batch_index_build.restype = C.c_void_p; lib.xheh_batch_index_build.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.xheh_batch_index_free.argtypes = [C.c_void_p]; lib.xheh_batch_index_free.restype = None
    lib.xheh_batch_index_bytes.argtypes = [C.c_void_p]; lib.xheh_batch_index_bytes.restype = C.c_size_t
    t0 = time.perf_counter()
    bl.index = lib.xheh_batch_index_build(bl.ptrs, bl.lens, bl.n, threads)
    bl.index_build_ms = 1e3 * (time.perf_counter() - t0)
    bl.index_bytes = int(lib.xheh_batch_index_bytes(bl.index))


class _Blobs:
    def __init__(self, blobs, index=False, index_threads=0):
        n = len(blobs)
        self.keep = [C.create_string_buffer(b, len(b)) for b in blobs]
        self.ptrs = (C.c_void_p * max(n, 1))(*[C.addressof(k) for k in self.keep])
        self.lens = (C.c_size_t * max(n, 1))(*[len(b) for b in blobs])
        self.n = n
        self.index = None
        if index and n:
            self.lib = _lib()
            _attach_index(self, self.lib, index_threads)

    def __del__(self):
        try:
            if getattr(self, "index", None):
                self.lib.xheh_batch_index_free(self.index); self.index = None
        except Exception:
            pass


class _PinnedBlobs:
    """the batch in ONE page-locked buffer, laid out as the device reads it (blobs back to back, padded to 16 bytes): the fast
    path uploads it in place -- zero-copy input.  Same interface as _Blobs."""

    def __init__(self, blobs, index=True, index_threads=0):
        lib = _lib()
        lib.xheh_blob_arena_alloc.restype = C.c_void_p; lib.xheh_blob_arena_alloc.argtypes = [C.c_size_t]
        lib.xheh_blob_arena_free.argtypes = [C.c_void_p]
        n = len(blobs)
        offs, off = [], 0
        for b in blobs:
            offs.append(off); off += (len(b) + 15) & ~15
        self.lib, self.base = lib, lib.xheh_blob_arena_alloc(max(off, 16))
        if not self.base:
            raise XheError(-3, "page-locked allocation failed")
        buf = (C.c_uint8 * max(off, 16)).from_address(self.base)
        mv = memoryview(buf).cast("B")
        for b, o in zip(blobs, offs):
            mv[o:o + len(b)] = b
        self.ptrs = (C.c_void_p * max(n, 1))(*[self.base + o for o in offs])
        self.lens = (C.c_size_t * max(n, 1))(*[len(b) for b in blobs])
        self.n = n
        self.index = None
        if index and n:
            _attach_index(self, lib, index_threads)

    def __del__(self):
        try:
            if getattr(self, "index", None):
                self.lib.xheh_batch_index_free(self.index); self.index = None
            if self.base:
                self.lib.xheh_blob_arena_free(self.base); self.base = None
        except Exception:
            pass


def verify_batch(ctx, blobs, ledger, seed=None, threads=0, prepared=None, fiat_shamir="host", deterministic=False):
    """Transaction::verify_batch.  Returns (code, first_failing_tx, timings dict); code 0 = Ok, >0 = verdicts (ERR_NAMES).
    fiat_shamir = "host" (Merlin transcripts on host threads, north_star's split), "device"  or "fast"
    (device transcripts + device-side batch layout, optimistic; any failure is re-decided by the exact path)."""
    lib = _lib()
    bl = prepared or _Blobs(blobs)
    fi = C.c_long(-1)
    tm = (C.c_double * 7)()
    rc = lib.xheh_verify_batch_ex(ctx.p, ledger.ptr, bl.ptrs, bl.lens, bl.n, seed, len(seed) if seed else 0, threads, _MODE_FLAGS[fiat_shamir] | (_DETERMINISTIC if deterministic else 0), C.byref(fi), tm, None)
    if rc == -5:      # XHE_E_CAPACITY: nothing was verified or applied; the index names the transaction
        e = XheError(rc, f"transaction {fi.value} needs more range-proof parties than this context's party_capacity"); e.index = fi.value
        raise e
    if rc < 0:
        raise XheError(rc, lib.xhe_last_error(ctx.p).decode())
    keys = ("parse_ms", "resolve_ms", "transcript_ms", "device_ms", "finish_ms", "total_ms", "keccak_f")
    d = dict(zip(keys, tm)); d["fast_path"] = d["keccak_f"] < 0
    return rc, fi.value, d


def verify_batch_partial(ctx, blobs, ledger, seed=None, threads=0, prepared=None, fiat_shamir="host", deterministic=False):
    """Shard mode for multi-GPU batches: (local code, first failing local tx, sigma partial enc, range partial enc, timings).
    The sigma / range identity decisions are left to the caller (xelis_he_b200.distributed); balance updates are held back
    until commit_pending()."""
    lib = _lib()
    bl = prepared or _Blobs(blobs)
    fi = C.c_long(-1)
    tm = (C.c_double * 7)()
    part = C.create_string_buffer(64)
    rc = lib.xheh_verify_batch_ex(ctx.p, ledger.ptr, bl.ptrs, bl.lens, bl.n, seed, len(seed) if seed else 0, threads, 2 | _MODE_FLAGS[fiat_shamir] | (_DETERMINISTIC if deterministic else 0), C.byref(fi), tm, part)
    if rc < 0:
        raise XheError(rc, lib.xhe_last_error(ctx.p).decode())
    keys = ("parse_ms", "resolve_ms", "transcript_ms", "device_ms", "finish_ms", "total_ms", "keccak_f")
    d = dict(zip(keys, tm)); d["fast_path"] = d["keccak_f"] < 0
    return rc, fi.value, part.raw[:32], part.raw[32:], d


def verify_batch_shard(ctx, blobs, ledger, lo, hi, seed=None, threads=0, prepared=None, fiat_shamir="host", deterministic=False, use_index=True):
    """One rank's share of a sharded batch.  `blobs` is the WHOLE batch (every rank holds it); this call
    verifies transactions [lo, hi) and reads the earlier ones only to follow the (account, asset) balance chains -- and the
    multisig settings -- its own transactions depend on, so the verdicts equal the reference's sequential walk
    (src/tx/verify.rs:301-374) however the batch is cut.  When `prepared` carries the batch's key-digest index
    (prepare_blobs_pinned builds it), the earlier shards are searched through it instead of through their bytes.  Returns (local code, first failing tx as an index into the whole
    batch, sigma partial enc, range partial enc, timings); state updates are held back until commit_pending()."""
    lib = _lib()
    bl = prepared or _Blobs(blobs)
    fi = C.c_long(-1)
    tm = (C.c_double * 7)()
    part = C.create_string_buffer(64)
    flags = _MODE_FLAGS[fiat_shamir] | (_DETERMINISTIC if deterministic else 0)
    index = getattr(bl, "index", None) if use_index else None
    if index:
        lib.xheh_verify_batch_shard_ix.restype = C.c_int32
        lib.xheh_verify_batch_shard_ix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.c_uint32,
                                                   C.POINTER(C.c_long), C.c_void_p, C.c_void_p, C.c_void_p]
        rc = lib.xheh_verify_batch_shard_ix(ctx.p, ledger.ptr, bl.ptrs, bl.lens, bl.n, lo, hi, seed, len(seed) if seed else 0, threads, flags, C.byref(fi), tm, part, index)
    else:
        rc = lib.xheh_verify_batch_shard(ctx.p, ledger.ptr, bl.ptrs, bl.lens, bl.n, lo, hi, seed, len(seed) if seed else 0, threads, flags, C.byref(fi), tm, part)
    if rc < 0:
        raise XheError(rc, lib.xhe_last_error(ctx.p).decode())
    keys = ("parse_ms", "resolve_ms", "transcript_ms", "device_ms", "finish_ms", "total_ms", "keccak_f")
    d = dict(zip(keys, tm)); d["fast_path"] = d["keccak_f"] < 0
    return rc, fi.value, part.raw[:32], part.raw[32:], d


def commit_pending(ctx, ledger):
    return _lib().xheh_commit_pending(ctx.p, ledger.ptr)


def export_taken(handle):
    """the balance updates a detached shard-mode batch holds (take_pending), as 128-byte records (account, asset, new
    ciphertext) in update order -- what a rank sends to its peers when every replica of the state must stay complete"""
    lib = _lib()
    n = lib.xheh_export_taken(handle, None, 0)
    buf = C.create_string_buffer(128 * max(n, 1))
    lib.xheh_export_taken(handle, buf, 128 * n)
    return buf.raw[:128 * n]


def take_pending(ctx):
    """Detach the balance updates the last shard-mode batch on `ctx` held back, so the context can take its next batch
    before the cross-rank decision is known.  Commit them with commit_taken or release them with drop_taken."""
    lib = _lib()
    lib.xheh_take_pending.restype = C.c_void_p; lib.xheh_take_pending.argtypes = [C.c_void_p]
    return lib.xheh_take_pending(ctx.p)


def commit_taken(handle, ledger):
    lib = _lib()
    lib.xheh_commit_taken.restype = C.c_int32; lib.xheh_commit_taken.argtypes = [C.c_void_p, C.c_void_p]
    return lib.xheh_commit_taken(handle, ledger.ptr)


def drop_taken(handle):
    lib = _lib()
    lib.xheh_drop_taken.restype = None; lib.xheh_drop_taken.argtypes = [C.c_void_p]
    if handle:
        lib.xheh_drop_taken(handle)


def verify(ctx, blob, ledger, seed=None):
    return verify_batch(ctx, [blob], ledger, seed)[:2]


def apply_without_verify(ctx, blobs, ledger):
    lib = _lib()
    bl = _Blobs(blobs)
    rc = lib.xheh_apply_without_verify(ctx.p, ledger.ptr, bl.ptrs, bl.lens, bl.n)
    if rc < 0:
        raise XheError(rc, lib.xhe_last_error(ctx.p).decode())
    return rc


def shard_dependencies(prepared, lo, hi, use_index=True, threads=1):
    """the transactions of [0, lo) that shard [lo, hi) of the prepared batch depends on (host-only; no device work)"""
    lib = _lib()
    lib.xheh_shard_dependencies.restype = C.c_long
    lib.xheh_shard_dependencies.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    cap = max(lo, 1)
    out = (C.c_size_t * cap)()
    k = lib.xheh_shard_dependencies(prepared.ptrs, prepared.lens, prepared.n, lo, hi, getattr(prepared, "index", None) if use_index else None, threads, out, cap)
    if k < 0:
        raise XheError(4, "a transaction of the shard does not parse")
    return list(out[:k])


prepare_blobs = _Blobs
prepare_blobs_pinned = _PinnedBlobs
