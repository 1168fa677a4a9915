"""Multi-GPU batch verification: one process per GPU, transactions sharded across ranks.

Every rank verifies its contiguous shard with its own GPU (`verifier.verify_batch_partial`): per-TX checks are local, and
the two big multiscalar multiplications are computed as PARTIAL sums (each rank also folds its own share of the static
Bulletproofs generator scalars into its partial, so no scalars are exchanged).  The only exchange is one small
all-gather per batch: (local verdict, first failing tx, sigma partial encoding, range partial encoding) = 72 bytes per
rank over NCCL (or gloo on CPU in the tests).  Ristretto encodings are canonical, so summing the decoded partials and
testing the identity is exactly the reference's `mega_check.is_identity()` on the whole batch (src/proofs.rs:49-67).
NCCL has no user-defined reduction, hence all-gather + local add rather than an all-reduce.
"""
import os
import struct

OK, GENERIC_PROOF, RANGE_PROOF = 0, 5, 6


def pack_local(code, fail_idx, shard_offset, sigma_enc, range_enc):
    gidx = shard_offset + fail_idx if fail_idx >= 0 else -1
    return struct.pack("<iq", code, gidx) + sigma_enc + range_enc + bytes(4)      # 80 bytes


def decide(records, sum_is_identity):
    """records: per-rank 80-byte records in rank (= shard) order; sum_is_identity(list of 32-byte encodings) -> bool.
    Mirrors the reference's order: first failing tx in batch order, then the sigma check, then the range check."""
    parsed = [struct.unpack("<iq", r[:12]) + (r[12:44], r[44:76]) for r in records]
    for code, gidx, _, _ in parsed:          # shards are contiguous and in order: the lowest rank with a per-tx error wins
        if code != OK and code not in (RANGE_PROOF,) or (code == RANGE_PROOF and gidx >= 0):
            return code, gidx
    if not sum_is_identity([p[2] for p in parsed]):
        return GENERIC_PROOF, -1
    if any(code == RANGE_PROOF for code, _, _, _ in parsed):   # structural range-proof failure inside a shard
        return RANGE_PROOF, -1
    if not sum_is_identity([p[3] for p in parsed]):
        return RANGE_PROOF, -1
    return OK, -1


_pinned = {}


def sum_is_identity(ctx, encodings):
    """is the sum of the ranks' partial results the identity?  Honest shards each report the identity encoding (32 zero
    bytes) -- the common case needs no device work; anything else is decoded and added on the device (xhe_sum_encodings)."""
    zero = bytes(32)
    if all(e == zero for e in encodings):
        return True
    return ctx.sum_encodings(b"".join(encodings))[1]


def all_gather_records(local_record, group=None, device=None, ctx=None):
    """One all-gather of the fixed-size per-rank records.  On CUDA the record travels through pinned host buffers and the
    host waits on a blocking-sync event: a synchronous copy to pageable memory would sit inside the driver for as long as
    the slowest rank takes to arrive, and other threads' kernel launches (further batches in flight) queue up behind it.
    With `ctx` (a context that owns a stream) the two tiny host<->device hops are kernel copies on that stream instead of
    copy-engine transfers, which would wait behind the megabyte uploads of the batches in flight."""
    import threading
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = len(local_record)
    if device is None or device.type != "cuda":
        t = torch.frombuffer(bytearray(local_record), dtype=torch.uint8)
        out = torch.empty(world * n, dtype=torch.uint8)
        dist.all_gather_into_tensor(out, t, group=group)
        raw = bytes(out.numpy())
        return [raw[i * n:(i + 1) * n] for i in range(world)]
    own_stream = ctx is not None and getattr(ctx, "stream_ptr", None)
    key = (threading.get_ident(), n, world, device.index, bool(own_stream))
    buf = _pinned.get(key)
    if buf is None:
        stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=device) if own_stream else torch.cuda.Stream(device=device)
        buf = _pinned[key] = (torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(world * n, dtype=torch.uint8, pin_memory=True),
                              torch.empty(n, dtype=torch.uint8, device=device), torch.empty(world * n, dtype=torch.uint8, device=device),
                              torch.cuda.Event(blocking=True), stream)
    h_in, h_out, d_in, d_out, ev, stream = buf
    h_in.copy_(torch.frombuffer(bytearray(local_record), dtype=torch.uint8))
    with torch.cuda.stream(stream):
        if own_stream:
            ctx.copy_small(d_in.data_ptr(), h_in.data_ptr(), n)
        else:
            d_in.copy_(h_in, non_blocking=True)
        dist.all_gather_into_tensor(d_out, d_in, group=group)
        if own_stream:
            ctx.copy_small(h_out.data_ptr(), d_out.data_ptr(), world * n)
        else:
            h_out.copy_(d_out, non_blocking=True)
        ev.record(stream)
    ev.synchronize()
    raw = bytes(h_out.numpy())
    return [raw[i * n:(i + 1) * n] for i in range(world)]


class OrderedGatherer:
    """All-gathers for several batches in flight.  NCCL collectives must be issued in the same order on every rank, so the
    worker threads never call NCCL themselves: they submit (sequence number, record) and one thread per process performs the
    all-gathers strictly in sequence order on a single communicator."""

    def __init__(self, group=None, device=None):
        import threading
        self.group, self.device = group, device
        self.cv = threading.Condition()
        self.pending, self.results, self.next_seq, self.stop = {}, {}, 0, False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        import torch
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        while True:
            with self.cv:
                while self.next_seq not in self.pending and not self.stop:
                    self.cv.wait(0.05)
                if self.stop and self.next_seq not in self.pending:
                    return
                rec = self.pending.pop(self.next_seq)
            out = all_gather_records(rec, self.group, self.device)
            with self.cv:
                self.results[self.next_seq] = out
                self.next_seq += 1
                self.cv.notify_all()

    def gather(self, seq, record):
        with self.cv:
            self.pending[seq] = record
            self.cv.notify_all()
            while seq not in self.results:
                self.cv.wait(0.05)
            return self.results.pop(seq)

    def close(self):
        with self.cv:
            self.stop = True
            self.cv.notify_all()
        self.thread.join(5)


class AsyncDecider:
    """Cross-rank decisions off the verification threads.  A worker submits (sequence number, its shard's record, the
    detached balance updates, the ledger they belong to) and moves on to its next batch; one thread per process all-gathers
    the records strictly in sequence order (NCCL's ordering rule), sums the partial encodings on its own small context,
    commits or drops the updates, and keeps the verdicts.  drain() waits for everything submitted so far."""

    def __init__(self, ctx, group=None, device=None):
        import threading
        self.ctx, self.group, self.device = ctx, group, device
        self.cv = threading.Condition()
        self.pending, self.verdicts, self.next_seq, self.submitted, self.stop, self.error = {}, {}, 0, 0, False, None
        self.stats = {"n": 0, "exchanges": 0, "exchange_ms": 0.0, "decide_ms": 0.0, "commit_ms": 0.0}      # time spent by the decision thread
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    SLOTS = 8          # decisions per exchange at most: a backlog is cleared with one all-gather instead of one each

    def _run(self):
        import time
        import torch
        from . import verifier
        try:
            if self.device is not None and self.device.type == "cuda":
                torch.cuda.set_device(self.device)
            while True:
                with self.cv:
                    while self.next_seq not in self.pending and not self.stop:
                        self.cv.wait(0.05)
                    if self.next_seq not in self.pending:
                        return
                    mine = []
                    while len(mine) < self.SLOTS and self.next_seq + len(mine) in self.pending:
                        mine.append(self.pending[self.next_seq + len(mine)])
                # message: how many consecutive decisions this rank is ready for, then that many records (fixed size)
                rec_len = len(mine[0][0])
                msg = struct.pack("<I", len(mine)) + b"".join(m[0] for m in mine) + bytes(rec_len * (self.SLOTS - len(mine)))
                t0 = time.perf_counter()
                if os.environ.get("XHE_DIAG_LOCAL_DECISION"):      # diagnostics: skip the exchange, decide on this rank's records alone
                    msgs = [msg]
                else:
                    msgs = all_gather_records(msg, self.group, self.device, self.ctx)
                t1 = time.perf_counter()
                ready = min(struct.unpack("<I", m[:4])[0] for m in msgs)      # every rank computes the same number (>= 1)
                for j in range(ready):
                    t2 = time.perf_counter()
                    records = [m[4 + rec_len * j:4 + rec_len * (j + 1)] for m in msgs]
                    verdict = decide(records, lambda encs: sum_is_identity(self.ctx, encs))
                    t3 = time.perf_counter()
                    _, handle, ledger = mine[j]
                    if handle:
                        if verdict[0] == OK and ledger is not None:
                            verifier.commit_taken(handle, ledger)
                        else:
                            verifier.drop_taken(handle)
                    self.stats["decide_ms"] += 1e3 * (t3 - t2); self.stats["commit_ms"] += 1e3 * (time.perf_counter() - t3)
                    with self.cv:
                        del self.pending[self.next_seq]
                        self.verdicts[self.next_seq] = verdict
                        self.next_seq += 1
                        self.cv.notify_all()
                self.stats["n"] += ready; self.stats["exchanges"] += 1; self.stats["exchange_ms"] += 1e3 * (t1 - t0)
        except Exception as e:          # surface the failure to drain() instead of hanging the workers
            with self.cv:
                self.error = e
                self.cv.notify_all()

    def submit(self, seq, record, handle=None, ledger=None):
        with self.cv:
            self.pending[seq] = (record, handle, ledger)
            self.submitted += 1
            self.cv.notify_all()

    def drain(self, upto, timeout=120.0):
        """wait until every sequence number below `upto` is decided; returns {seq: (code, first failing tx)}"""
        import time
        deadline = time.time() + timeout
        with self.cv:
            while self.next_seq < upto and self.error is None:
                if time.time() > deadline:
                    raise TimeoutError("cross-rank decisions did not complete")
                self.cv.wait(0.05)
            if self.error is not None:
                raise self.error
            return dict(self.verdicts)

    def close(self):
        with self.cv:
            self.stop = True
            self.cv.notify_all()
        self.thread.join(5)


def verify_batch_distributed(ctx, shard_blobs, ledger, shard_offset, group=None, seed=None, threads=0, prepared=None, commit=True, fiat_shamir="host", gather=None):
    """Transaction::verify_batch over a batch sharded across the ranks of `group`.  Returns (code, global first failing tx,
    timings).  On accept every rank commits its own shard's balance updates to its ledger.  `gather` (record -> list of
    records) replaces the direct all-gather when several batches are in flight (OrderedGatherer)."""
    import time
    import torch
    from . import verifier
    code, idx, s_enc, r_enc, tm = verifier.verify_batch_partial(ctx, shard_blobs, ledger, seed=seed, threads=threads, prepared=prepared, fiat_shamir=fiat_shamir)
    t0 = time.perf_counter()
    rec = pack_local(code, idx, shard_offset, s_enc, r_enc)
    records = gather(rec) if gather else all_gather_records(rec, group, torch.device("cuda", torch.cuda.current_device()))
    t1 = time.perf_counter()
    verdict = decide(records, lambda encs: sum_is_identity(ctx, encs))
    t2 = time.perf_counter()
    if verdict[0] == OK and commit:
        verifier.commit_pending(ctx, ledger)
    tm = dict(tm); tm["exchange_ms"] = 1e3 * (t1 - t0); tm["decide_ms"] = 1e3 * (t2 - t1); tm["commit_ms"] = 1e3 * (time.perf_counter() - t2)
    return verdict[0], verdict[1], tm
