"""Two-GPU test of the sharded verify_batch over NCCL (skipped with fewer than two devices)."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, blobs, records, q):
    import torch
    import torch.distributed as dist
    import xelis_he_b200 as xhe
    from xelis_he_b200 import distributed as xd, verifier
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = xhe.Ctx(rank, party_capacity=4)
    led = verifier.Ledger(); led.import_records(records)
    n = len(blobs); lo, hi = n * rank // world, n * (rank + 1) // world
    code, idx, tm = xd.verify_batch_distributed(ctx, blobs[lo:hi], led, lo, seed=b"dist%d" % rank, threads=2)
    q.put((rank, (code, idx), led.dump()))
    dist.barrier(); dist.destroy_process_group()


def test_two_gpu_sharded_batch():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import oracle
    b = oracle.mint_transfers(41, 12, 1, 1, threads=4)
    records = b.ledger().dump()
    for blobs in (b.blobs, b.blobs[:9] + [b.blobs[9][:-1] + bytes([b.blobs[9][-1] ^ 1])] + b.blobs[10:]):
        want_led = b.ledger(); want = oracle.verify_batch(blobs, want_led)
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        c = mp.get_context("spawn"); q = c.Queue()
        ps = [c.Process(target=_worker, args=(r, 2, port, blobs, records, q)) for r in range(2)]
        for p in ps:
            p.start()
        res = [q.get(timeout=300) for _ in range(2)]
        for p in ps:
            p.join(60)
        assert all(r[1] == want for r in res), (res[0][1], res[1][1], want)
        if want == (0, -1):
            # each rank applied its own shard: merging the two ledgers' changes reproduces the oracle's final state
            merged = dict(((pk, a), ct) for pk, a, ct in records)
            for _, _, dump in res:
                for pk, a, ct in dump:
                    if dict(((p2, a2), c2) for p2, a2, c2 in records)[(pk, a)] != ct:
                        merged[(pk, a)] = ct
            assert sorted((pk, a, ct) for (pk, a), ct in merged.items()) == sorted(want_led.dump())
