"""Mint the 8 x 10k-transaction batch of the rank emulator once (here, on CPU) so that GPU-box time is not spent on it:
  python tools/r02_mint_cache.py  ->  _cache/emul_80k.pkl  (git-ignored; travels with gpurun)"""
import os
import pickle
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle

WORLD, T = 8, 10000
t0 = time.time()
parts = [oracle.mint_transfers(77 + r, T, 1, 1, threads=len(os.sched_getaffinity(0))) for r in range(WORLD)]
blobs = [bytes(b) for p in parts for b in p.blobs]
records = [rec for p in parts for rec in p.ledger().dump()]
os.makedirs("_cache", exist_ok=True)
with open("_cache/emul_80k.pkl", "wb") as f:
    pickle.dump((blobs, records), f, protocol=4)
print("minted", len(blobs), "in", round(time.time() - t0, 1), "s")
