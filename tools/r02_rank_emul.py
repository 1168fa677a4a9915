"""Host-side cost of ONE rank of an 8-GPU sharded batch, measured on one GPU: the rank's shard of an 80k-transaction batch is
verified through verify_batch_shard with 6 batches in flight and the process pinned to 4 cores (what a rank has on a 32-core box
with 8 GPUs); the cross-rank exchange is left out (every shard is honest), the commit of the held-back updates is not.
  python tools/r02_rank_emul.py [rank ...]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import torch
import xelis_he_b200 as xhe
from xelis_he_b200 import verifier

WORLD, T, NFL, STEPS = 8, 10000, int(os.environ.get("NFL", "6")), int(os.environ.get("STEPS", "30"))
ncpu = len(os.sched_getaffinity(0))
t0 = time.time()
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_cache", "emul_80k.pkl")
if os.path.exists(cache):      # minted beforehand by tools/r02_mint_cache.py
    import pickle
    blobs, records = pickle.load(open(cache, "rb"))
else:
    parts = [oracle.mint_transfers(77 + r, T, 1, 1, threads=ncpu) for r in range(WORLD)]
    blobs = [b for p in parts for b in p.blobs]
    records = [rec for p in parts for rec in p.ledger().dump()]
print("minted", len(blobs), "in", round(time.time() - t0, 1), "s", flush=True)
cores = sorted(os.sched_getaffinity(0))[:int(os.environ.get("CORES", "4"))]
os.sched_setaffinity(0, cores)
prepared = verifier.prepare_blobs_pinned(blobs, index=not os.environ.get("NOINDEX"))
print("key index:", getattr(prepared, "index_bytes", 0), "bytes,", round(getattr(prepared, "index_build_ms", 0.0), 2), "ms on", len(cores), "cores", flush=True)
ledger0 = verifier.Ledger(); ledger0.import_records(records)
ctxs = [xhe.Ctx(0, party_capacity=2) for _ in range(NFL)]
streams = [torch.cuda.Stream() for _ in range(NFL)]
for c, s in zip(ctxs, streams):
    c.set_stream(s.cuda_stream)
for rank in [int(a) for a in sys.argv[1:]] or [0, 7]:
    lo, hi = rank * T, (rank + 1) * T
    commit_q, phases, lock = [], [], threading.Lock()
    done = threading.Event()

    def committer():
        n = 0
        while n < STEPS + NFL:
            with lock:
                item = commit_q.pop(0) if commit_q else None
            if item is None:
                time.sleep(0.0002); continue
            verifier.commit_taken(item[0], item[1]); n += 1
        done.set()

    def worker(w, nsteps, ledgers):
        torch.cuda.set_device(0)
        for s in range(nsteps):
            code, idx, se, re_, tm = verifier.verify_batch_shard(ctxs[w], None, ledgers[s], lo, hi, seed=b"e%d-%d" % (w, s), threads=1, prepared=prepared, fiat_shamir="fast")
            assert (code, idx) == (0, -1) and tm["fast_path"]
            with lock:
                commit_q.append((verifier.take_pending(ctxs[w]), ledgers[s])); phases.append(tm)
    threading.Thread(target=committer, daemon=True).start()
    th = [threading.Thread(target=worker, args=(w, 1, [ledger0.clone()])) for w in range(NFL)]      # warm-up
    [t.start() for t in th]; [t.join() for t in th]
    phases.clear()
    counts = [STEPS // NFL + (1 if w < STEPS % NFL else 0) for w in range(NFL)]
    fresh = [[ledger0.clone() for _ in range(counts[w])] for w in range(NFL)]
    th = [threading.Thread(target=worker, args=(w, counts[w], fresh[w])) for w in range(NFL)]
    t0 = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    done.wait(60); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mean = {k: round(sum(p[k] for p in phases) / len(phases), 3) for k in phases[0] if k.endswith("_ms")}
    print(f"rank {rank}: {T * STEPS / dt / 1e6:.3f} M TX/s on {len(cores)} cores, {NFL} in flight, {1e3 * dt / STEPS:.2f} ms per batch; mean phases {mean}", flush=True)
