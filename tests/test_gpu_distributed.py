"""Two-GPU test of the sharded verify_batch over NCCL (skipped with fewer than two devices; tests/test_gpu_shards.py plays
the same ranks one after the other on a single GPU)."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, blobs, records, sync_state, q):
    import torch
    import torch.distributed as dist
    import xelis_he_b200 as xhe
    from xelis_he_b200 import distributed as xd, verifier
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = xhe.Ctx(rank, party_capacity=4)
    led = verifier.Ledger(); led.import_records(records)
    code, idx, tm = xd.verify_batch_distributed(ctx, blobs, led, seed=b"dist", threads=2, fiat_shamir="fast", sync_state=sync_state)
    q.put((rank, (code, idx), led.dump()))
    dist.barrier(); dist.destroy_process_group()


def _run(blobs, records, sync_state):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    c = mp.get_context("spawn"); q = c.Queue()
    ps = [c.Process(target=_worker, args=(r, 2, port, blobs, records, sync_state, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in ps:
        p.join(60)
    return res


def test_two_gpu_sharded_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import oracle
    import scenarios
    b = oracle.mint_transfers(41, 12, 1, 1, threads=4)
    records = b.ledger().dump()
    bad = b.blobs[:9] + [b.blobs[9][:-1] + bytes([b.blobs[9][-1] ^ 1])] + b.blobs[10:]
    w, dep = scenarios.shared_receiver_world(6)                  # both shards credit one receiver, which spends in shard 1
    chain = oracle.mint_chain(9, 10, 1)                          # one sender: every transaction depends on the previous one
    for blobs, recs, led0 in ((b.blobs, records, b.ledger()), (bad, records, b.ledger()), (dep, w.records, w.ledger.clone()), (chain.blobs, chain.ledger().dump(), chain.ledger())):
        want_led = led0.clone(); want = oracle.verify_batch(blobs, want_led)
        res = _run(blobs, recs, sync_state=True)
        assert all(r[1] == want for r in res), (res[0][1], res[1][1], want)
        if want == (0, -1):
            # with the update exchange every rank's replica equals the oracle's final state
            assert all(r[2] == sorted(want_led.dump()) for r in res)
    # without the exchange each rank holds its own shard's updates: applying them in shard order reproduces the final state
    want_led = w.ledger.clone(); assert oracle.verify_batch(dep, want_led) == (0, -1)
    res = _run(dep, w.records, sync_state=False)
    merged = dict(((pk, a), ct) for pk, a, ct in w.records)
    init = dict(merged)
    for _, _, dump in res:
        for pk, a, ct in dump:
            if init[(pk, a)] != ct:
                merged[(pk, a)] = ct
    assert sorted((pk, a, ct) for (pk, a), ct in merged.items()) == sorted(want_led.dump())
