import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import xelis_he_b200 as xhe
from xelis_he_b200 import verifier
ctx = xhe.Ctx(0, party_capacity=8)
for T in (4, 64):
    b = oracle.mint_transfers(3, T, 1, 1, threads=4)
    for mode in ("host", "device", "fast"):
        led = verifier.Ledger(); led.import_records(b.ledger().dump())
        try:
            print(T, mode, verifier.verify_batch(ctx, b.blobs, led, seed=b"x", fiat_shamir=mode)[:2], flush=True)
        except Exception as e:
            print(T, mode, "ERR", e, flush=True)
