"""world_size-2 `gloo` test of the multi-GPU host logic on CPU: sharding, the all-gather of per-rank records and the joint
decision (xelis_he_b200/distributed.py).  The per-shard partial results come from the oracle here (no GPU); the decision
must equal the oracle's verdict on the whole batch."""
import os
import socket
import sys

import pytest


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, blobs, records, q):
    import torch.distributed as dist
    import oracle
    from xelis_he_b200 import distributed as xd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    led = oracle.Ledger()
    for pk, asset, ct in records:
        led.set_balance(pk, asset, ct); led.set_nonce(pk, 0)
    n = len(blobs); lo, hi = xd.shard_bounds(n, rank, world)
    # a shard sees the state as the earlier transactions of the batch left it (src/tx/verify.rs:301-374).  The GPU host layer
    # follows only the balance chains its shard reads; the oracle stand-in simply applies the whole prefix.
    for tx in blobs[:lo]:
        oracle.apply_without_verify(tx, led)
    code, idx, s_enc, r_enc = oracle.verify_batch_partial(blobs[lo:hi], led, rng_seed=100 + rank)
    recs = xd.all_gather_records(xd.pack_local(code, idx, lo, s_enc, r_enc))
    payload = xd.all_gather_bytes(bytes([rank]) * (3 * rank + 1))            # variable-length exchange used by sync_and_commit
    assert payload == [bytes([r]) * (3 * r + 1) for r in range(world)]

    def sum_is_identity(encs):
        acc = bytes(32)
        for e in encs:
            acc = oracle.point_add(acc, e)
        return acc == bytes(32)
    q.put((rank, xd.decide(recs, sum_is_identity)))
    dist.barrier(); dist.destroy_process_group()


def _run(blobs, records):
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue(); port = _free_port()
    ps = [ctxm.Process(target=_worker, args=(r, 2, port, blobs, records, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=240) for _ in range(2))
    for p in ps:
        p.join(60)
    assert out[0] == out[1]
    return out[0]


@pytest.mark.timeout(600)
def test_two_rank_sharded_verdicts_match_single_process_oracle():
    import oracle
    b = oracle.mint_transfers(21, 10, 1, 1, threads=4)
    records = b.ledger().dump()
    assert _run(b.blobs, records) == oracle.verify_batch(b.blobs, b.ledger()) == (0, -1)
    # a per-tx failure in the second shard is reported with its global index
    bad = list(b.blobs); x = bytearray(bad[7]); x[-40] ^= 1; bad[7] = bytes(x)
    assert _run(bad, records) == oracle.verify_batch(bad, b.ledger()) == (1, 7)
    # failures in both shards: the earliest transaction wins
    x = bytearray(bad[2]); x[56] ^= 1; bad[2] = bytes(x)
    assert _run(bad, records) == oracle.verify_batch(bad, b.ledger()) == (9, 2)


@pytest.mark.timeout(600)
def test_two_rank_dependent_transactions_match_single_process_oracle():
    """transactions of different shards on the same (account, asset): the realistic_test of the reference (src/lib.rs:831-949,
    tx2 spends what tx1 delivered) cut so that tx2 lands on rank 1; several senders of both shards crediting one receiver
    that then spends; a one-sender chain (benches/tx.rs:153-186)."""
    import oracle
    import scenarios
    w, txs, _ = scenarios.realistic_world()
    assert _run(txs, w.records) == oracle.verify_batch(txs, w.ledger.clone()) == (0, -1)
    w2, txs2 = scenarios.shared_receiver_world(6)
    assert _run(txs2, w2.records) == oracle.verify_batch(txs2, w2.ledger.clone()) == (0, -1)
    # the receiver's spend re-signed with a broken validity proof: rejected by the sigma check of the summed partials
    chain = oracle.mint_chain(9, 6, 1)
    assert _run(chain.blobs, chain.ledger().dump()) == oracle.verify_batch(chain.blobs, chain.ledger()) == (0, -1)
    swapped = [chain.blobs[0], chain.blobs[1], chain.blobs[2], chain.blobs[4], chain.blobs[3], chain.blobs[5]]
    assert _run(swapped, chain.ledger().dump()) == oracle.verify_batch(swapped, chain.ledger()) == (5, -1)


def test_decide_orders_checks_like_the_reference():
    from xelis_he_b200 import distributed as xd
    z = bytes(32); nz = bytes([1]) + bytes(31)
    ident = lambda encs: all(e == z for e in encs)   # stand-in group: only the all-zero encoding is the identity
    ok = xd.pack_local(0, -1, 0, z, z)
    assert xd.decide([ok, ok], ident) == (0, -1)
    assert xd.decide([ok, xd.pack_local(0, -1, 5, nz, z)], ident) == (5, -1)                 # sigma sum != identity
    assert xd.decide([ok, xd.pack_local(0, -1, 5, z, nz)], ident) == (6, -1)                 # range sum != identity
    assert xd.decide([xd.pack_local(0, -1, 0, nz, nz), ok], ident) == (5, -1)                # sigma is checked before range
    assert xd.decide([xd.pack_local(6, -1, 0, z, z), ok], ident) == (6, -1)                  # structural range failure in a shard
    assert xd.decide([xd.pack_local(6, -1, 0, z, z), xd.pack_local(0, -1, 5, nz, z)], ident) == (5, -1)
    assert xd.decide([ok, xd.pack_local(1, 3, 5, nz, nz)], ident) == (1, 8)                  # per-tx errors come first, global index
    assert xd.decide([xd.pack_local(9, 4, 0, z, z), xd.pack_local(1, 0, 5, z, z)], ident) == (9, 4)


class _OracleSummer:
    """stands in for the GPU context's sum_encodings in the CPU test of the decision thread"""

    def sum_encodings(self, encodings):
        import oracle
        acc = bytes(32)
        for i in range(0, len(encodings), 32):
            acc = oracle.point_add(acc, encodings[i:i + 32])
        return acc, acc == bytes(32)


def _worker_async(rank, world, port, batches, q):
    import threading
    import torch.distributed as dist
    import oracle
    from xelis_he_b200 import distributed as xd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    decider = xd.AsyncDecider(_OracleSummer(), None, None)

    def one(seq):
        blobs, records = batches[seq]
        led = oracle.Ledger()
        for pk, asset, ct in records:
            led.set_balance(pk, asset, ct); led.set_nonce(pk, 0)
        n = len(blobs); lo, hi = xd.shard_bounds(n, rank, world)
        for tx in blobs[:lo]:
            oracle.apply_without_verify(tx, led)
        code, idx, s_enc, r_enc = oracle.verify_batch_partial(blobs[lo:hi], led, rng_seed=7 * seq + rank)
        decider.submit(seq, xd.pack_local(code, idx, lo, s_enc, r_enc))
    # several batches in flight, submitted from threads in a rank-dependent order: decisions still pair up by sequence number
    order = list(range(len(batches)))
    if rank:
        order.reverse()
    th = [threading.Thread(target=one, args=(s,)) for s in order]
    for t in th:
        t.start()
    for t in th:
        t.join()
    verdicts = decider.drain(len(batches), timeout=200)
    decider.close()
    q.put((rank, [verdicts[s] for s in range(len(batches))]))
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_async_decider_orders_decisions_across_ranks():
    import oracle
    import torch.multiprocessing as mp
    good = oracle.mint_transfers(41, 6, 1, 1, threads=4)
    other = oracle.mint_transfers(42, 4, 1, 3, threads=4)
    bad = list(good.blobs); t = bytearray(bad[4]); t[-1] ^= 1; bad[4] = bytes(t)          # broken signature in the second shard
    batches = [(good.blobs, good.ledger().dump()), (bad, good.ledger().dump()), (other.blobs, other.ledger().dump())]
    want = [oracle.verify_batch(good.blobs, good.ledger()), oracle.verify_batch(bad, good.ledger()), oracle.verify_batch(other.blobs, other.ledger())]
    assert want == [(0, -1), (1, 4), (0, -1)]
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue(); port = _free_port()
    ps = [ctxm.Process(target=_worker_async, args=(r, 2, port, batches, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = dict(q.get(timeout=240) for _ in range(2))
    for p in ps:
        p.join(60)
    assert out[0] == out[1] == want


def test_ranks_bind_to_disjoint_core_sets():
    """one process per GPU keeps to its own share of the host cores (distributed.bind_rank_to_local_cores); without GPUs the
    current affinity mask is split evenly.  Run in child processes: the call changes the caller's affinity."""
    import multiprocessing as mp
    import os

    avail = sorted(os.sched_getaffinity(0))
    if len(avail) < 2:
        pytest.skip("needs two host cores")
    ctx = mp.get_context("spawn")
    with ctx.Pool(2) as pool:
        got = pool.starmap(_bind_child, [(0, 2), (1, 2)])
    assert got[0] and got[1] and not (set(got[0]) & set(got[1]))
    assert sorted(got[0] + got[1]) == avail[:len(got[0]) + len(got[1])] or set(got[0] + got[1]) <= set(avail)
    assert abs(len(got[0]) - len(got[1])) <= 1


def _bind_child(rank, world):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from xelis_he_b200.distributed import bind_rank_to_local_cores
    cores = bind_rank_to_local_cores(rank, world)
    assert sorted(os.sched_getaffinity(0)) == sorted(cores)
    return cores
