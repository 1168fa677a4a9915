"""Decoding of decrypted amounts on B200: table build time, decodes / s at range 2^32 and decrypt + decode / s.
  python tools/ecdlp_bench.py [l1_bits ...]"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import xelis_he_b200 as xhe

ctx = xhe.Ctx(0, party_capacity=2)
G = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
r = random.Random(3)
N = int(os.environ.get("N", "32768"))
kp = oracle.Keypair.derive(b"ecdlp-bench"); rng = oracle.Rng(b"ecdlp-bench")
amounts = [r.randrange(2**32) for _ in range(N)]
t0 = time.time()
cts = b"".join(kp.encrypt(a, rng) for a in amounts[:4096])
pts = b"".join(oracle.scalarmult(a.to_bytes(32, "little"), G) for a in amounts[:4096])
reps = N // 4096
cts, pts, amounts = cts * reps, pts * reps, amounts[:4096] * reps
print("inputs minted in", round(time.time() - t0, 1), "s", flush=True)
res = {}
for l1 in [int(a) for a in sys.argv[1:]] or [20, 22, 24]:
    t0 = time.perf_counter(); tab = xhe.Ecdlp(ctx, l1_bits=l1); build = time.perf_counter() - t0
    tab.decode(pts[:32 * 1024], 32)
    t0 = time.perf_counter(); got, st = tab.decode(pts, 32); dt = time.perf_counter() - t0
    assert got == amounts and st == bytes([1]) * len(amounts)
    t0 = time.perf_counter(); got2, st2 = tab.decrypt_decode(kp.sk, cts, 32); dt2 = time.perf_counter() - t0
    assert got2 == amounts
    res["l1_%d" % l1] = {"table_MB": tab.table_bytes / 2**20, "build_ms": round(1e3 * build, 2), "points": len(amounts), "decode_per_s": round(len(amounts) / dt), "decrypt_decode_per_s": round(len(amounts) / dt2),
                         "giant_steps_per_decode": 2 ** (32 - l1 - 1)}
    print(l1, res["l1_%d" % l1], flush=True)
    tab.close()
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/ecdlp_bench.json", "w"), indent=1)
