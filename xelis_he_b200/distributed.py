"""Multi-GPU batch verification: one process per GPU, transactions sharded across ranks.

Every rank holds the WHOLE batch (as every node of a network receives the whole block) and verifies one contiguous shard of
it with its own GPU (`verifier.verify_batch_shard`): per-TX checks are local, and the two big multiscalar multiplications are
computed as PARTIAL sums (each rank also folds its own share of the static Bulletproofs generator scalars into its partial,
so no scalars are exchanged).  The only exchange on the verification path is one small all-gather per batch: (local verdict,
first failing tx, sigma partial encoding, range partial encoding) = 80 bytes per rank over NCCL (or gloo on CPU in the
tests).  Ristretto encodings are canonical, so summing the decoded partials and testing the identity is exactly the
reference's `mega_check.is_identity()` on the whole batch (src/proofs.rs:49-67).  NCCL has no user-defined reduction, hence
all-gather + local add rather than an all-reduce.

Dependent transactions.  The reference threads `state` through the batch in order (src/tx/verify.rs:301-336,354-374): a
transaction's source ciphertext is the initial balance moved by EVERY earlier transaction on that (account, asset) -- its own
earlier spends and what it received (src/lib.rs:908-921).  A shard therefore follows the balance chains of its keys back
through the earlier shards: the host layer replays the group operations of the earlier transactions that touch those keys
(no proofs -- their own rank verifies them; no communication -- the bytes are in the batch), so verdicts and balances equal
the single-process result however the batch is cut.  After an accepted batch a rank commits its own shard's updates; with
`sync_state=True` the ranks also exchange their updates (one all-gather of 128 bytes per touched balance) and apply them in
shard order, so every replica of the state ends up identical to the reference's.
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
    per_tx = [(gidx, code) for code, gidx, _, _ in parsed if code != OK and gidx >= 0]
    if per_tx:                               # the first failing transaction of the whole batch (src/tx/verify.rs:492-498)
        gidx, code = min(per_tx)
        return code, gidx
    for code, gidx, _, _ in parsed:          # a shard-level error without a transaction (state backend failure)
        if code != OK and code not in (GENERIC_PROOF, RANGE_PROOF):
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
    the records strictly in sequence order (NCCL's ordering rule) and sums the partial encodings on its own small context;
    a second thread commits or drops the held-back updates in the same order, so that the exchange cadence never waits for
    a ledger update (tens of thousands of hash-table writes per batch).  drain() waits for everything submitted so far."""

    def __init__(self, ctx, group=None, device=None):
        import collections
        import threading
        self.ctx, self.group, self.device = ctx, group, device
        self.cv = threading.Condition()
        self.pending, self.verdicts, self.next_seq, self.done_seq, self.submitted, self.stop, self.error = {}, {}, 0, 0, 0, False, None
        self.commit_q = collections.deque()
        self.stats = {"n": 0, "exchanges": 0, "exchange_ms": 0.0, "decide_ms": 0.0, "commit_ms": 0.0}      # time spent by the two threads
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.committer = threading.Thread(target=self._commit_loop, daemon=True)
        self.thread.start(); self.committer.start()

    SLOTS = 8          # decisions per exchange at most: a backlog is cleared with one all-gather instead of one each

    def _commit_loop(self):
        import time
        from . import verifier
        try:
            while True:
                with self.cv:
                    while not self.commit_q and not self.stop:
                        self.cv.wait(0.05)
                    if not self.commit_q:
                        return
                    seq, verdict, handle, ledger = self.commit_q.popleft()
                t0 = time.perf_counter()
                if handle:
                    if verdict[0] == OK and ledger is not None:
                        verifier.commit_taken(handle, ledger)
                    else:
                        verifier.drop_taken(handle)
                with self.cv:
                    self.stats["commit_ms"] += 1e3 * (time.perf_counter() - t0)
                    self.done_seq = seq + 1
                    self.cv.notify_all()
        except Exception as e:
            with self.cv:
                self.error = e
                self.cv.notify_all()

    def _run(self):
        import time
        import torch
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
                    records = [m[4 + rec_len * j:4 + rec_len * (j + 1)] for m in msgs]
                    verdict = decide(records, lambda encs: sum_is_identity(self.ctx, encs))
                    _, handle, ledger = mine[j]
                    with self.cv:
                        del self.pending[self.next_seq]
                        self.verdicts[self.next_seq] = verdict
                        self.commit_q.append((self.next_seq, verdict, handle, ledger))
                        self.next_seq += 1
                        self.cv.notify_all()
                self.stats["n"] += ready; self.stats["exchanges"] += 1; self.stats["exchange_ms"] += 1e3 * (t1 - t0); self.stats["decide_ms"] += 1e3 * (time.perf_counter() - t1)
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
        """wait until every sequence number below `upto` is decided and its updates committed or dropped; returns {seq: (code, first failing tx)}"""
        import time
        deadline = time.time() + timeout
        with self.cv:
            while self.done_seq < upto and self.error is None:
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
        self.thread.join(5); self.committer.join(5)


def bind_rank_to_local_cores(local_rank, local_world, verbose=False):
    """Give this rank its own share of the host cores, on the NUMA node of its GPU.  With one process per GPU and a host
    phase per batch (parsing, the state walk, launches), ranks that float over all cores of a two-socket box walk memory on
    the other socket, feed their GPU across the inter-socket link and pre-empt each other.  The cores local to each visible
    GPU are read from sysfs (`local_cpulist` of its PCI device); the ranks whose GPUs share a core list split it evenly.
    Falls back to an even split of the current affinity mask.  Call it BEFORE allocating ledgers or pinned buffers (first
    touch decides where they live).  Returns the list of cores."""
    cur = sorted(os.sched_getaffinity(0))

    def parse(txt):
        out = []
        for part in txt.strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            out.extend(range(int(a), int(b or a) + 1))
        return out
    lists = None
    try:
        import torch
        lists = []
        for d in range(local_world):
            pr = torch.cuda.get_device_properties(d)
            bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            lists.append(tuple(c for c in parse(open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read()) if c in cur))
        if not all(lists):
            lists = None
    except Exception:
        lists = None
    if lists is None:
        lists = [tuple(cur)] * local_world
    mine = lists[local_rank]
    peers = [r for r in range(local_world) if lists[r] == mine]
    i, k = peers.index(local_rank), len(peers)
    cores = list(mine[len(mine) * i // k:len(mine) * (i + 1) // k]) or list(mine)
    os.sched_setaffinity(0, cores)
    if verbose:
        import sys
        print("rank %d: cores %s" % (local_rank, cores), file=sys.stderr, flush=True)
    return cores


def shard_bounds(n, rank, world):
    """contiguous shard [lo, hi) of rank `rank` in a batch of n transactions"""
    return n * rank // world, n * (rank + 1) // world


def all_gather_bytes(data: bytes, group=None, device=None):
    """all-gather of one variable-length byte string per rank (lengths first, then the padded payloads)"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = device if device is not None and device.type == "cuda" else torch.device("cpu")
    ln = torch.tensor([len(data)], dtype=torch.int64, device=dev)
    lens = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(lens, ln, group=group)
    lens = [int(x) for x in lens.cpu().tolist()]
    cap = max(max(lens), 1)
    buf = torch.zeros(cap, dtype=torch.uint8)
    if data:
        buf[:len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8)
    out = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf.to(dev), group=group)
    raw = bytes(out.cpu().numpy())
    return [raw[r * cap:r * cap + lens[r]] for r in range(world)]


def sync_and_commit(ctx, ledger, group=None, device=None):
    """After an accepted sharded batch: exchange the ranks' balance updates and apply them in shard order (a later shard's
    value of a shared balance is the final one), so every rank's replica of the state equals the reference's final state."""
    import torch.distributed as dist
    from . import verifier
    rank = dist.get_rank(group)
    handle = verifier.take_pending(ctx)
    mine = verifier.export_taken(handle) if handle else b""
    parts = all_gather_bytes(mine, group, device)
    rc = 0
    for r, recs in enumerate(parts):
        if r == rank:
            if handle:
                rc = rc or verifier.commit_taken(handle, ledger)      # own shard: nonces / multisig settings / output ciphertexts too
        elif recs:
            rc = rc or ledger.apply_records(recs)
    return rc


def verify_batch_distributed(ctx, blobs, ledger, group=None, seed=None, threads=0, prepared=None, commit=True, fiat_shamir="host", gather=None,
                             sync_state=False, deterministic=False):
    """Transaction::verify_batch over a batch sharded across the ranks of `group`; EVERY rank passes the whole batch.
    Returns (code, first failing tx, timings) -- the same on every rank and equal to the single-process verdict, also when
    transactions of different shards touch the same (account, asset).  On accept every rank commits its own shard's balance
    updates to its ledger (`sync_state=True`: all shards' updates, see sync_and_commit).  `gather` (record -> list of records)
    replaces the direct all-gather when several batches are in flight (OrderedGatherer)."""
    import time
    import torch
    import torch.distributed as dist
    from . import verifier
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = prepared.n if prepared is not None else len(blobs)
    lo, hi = shard_bounds(n, rank, world)
    code, idx, s_enc, r_enc, tm = verifier.verify_batch_shard(ctx, blobs, ledger, lo, hi, seed=seed, threads=threads, prepared=prepared, fiat_shamir=fiat_shamir, deterministic=deterministic)
    t0 = time.perf_counter()
    rec = pack_local(code, idx, 0, s_enc, r_enc)            # idx is already an index into the whole batch
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    records = gather(rec) if gather else all_gather_records(rec, group, dev)
    t1 = time.perf_counter()
    verdict = decide(records, lambda encs: sum_is_identity(ctx, encs))
    t2 = time.perf_counter()
    if verdict[0] == OK and commit:
        if sync_state:
            sync_and_commit(ctx, ledger, group, dev)
        else:
            verifier.commit_pending(ctx, ledger)
    elif verdict[0] != OK:
        verifier.drop_taken(verifier.take_pending(ctx))
    tm = dict(tm); tm["exchange_ms"] = 1e3 * (t1 - t0); tm["decide_ms"] = 1e3 * (t2 - t1); tm["commit_ms"] = 1e3 * (time.perf_counter() - t2)
    return verdict[0], verdict[1], tm
