#!/usr/bin/env python3
"""Generates tests/golden/vectors.json from the CPU oracle (the reference itself cannot run in this image: Rust, no cargo,
un-vendored crates).  The oracle is pinned separately (tests/test_oracle_kat.py); these fixtures freeze its outputs so the
CUDA path and any future oracle change are both checked against committed bytes.   python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle
import scenarios

out = {"msm": [], "ct_update": None, "batches": []}
for n in (1, 4, 64, 300):
    s, p = oracle.gen_msm_inputs(4242 + n, n)
    out["msm"].append({"n": n, "scalars": s.hex(), "points": p.hex(), "expect": oracle.msm(s, p).hex()})
rnd = hashlib.shake_256(b"golden-ct").digest(64 * 64)
pts = [oracle.from_uniform(rnd[64 * i:64 * i + 64]) for i in range(64)]
bal, delta, sub = b"".join(pts[:32]), b"".join(pts[32:]), bytes(i & 1 for i in range(16))
res, ok = oracle.ct_update(bal, delta, sub)
out["ct_update"] = {"bal": bal.hex(), "delta": delta.hex(), "sub": sub.hex(), "expect": res.hex(), "ok": ok.hex()}


def add_batch(name, world, txs):
    led = world.ledger.clone()
    code, idx = oracle.verify_batch(txs, led)
    out["batches"].append({"name": name, "txs": [t.hex() for t in txs], "records": [[a.hex(), b.hex(), c.hex()] for a, b, c in world.records],
                           "multisig": [[pk.hex(), [s.hex() for s in signers], th] for pk, signers, th in world.multisig],
                           "expect": [code, idx], "final": [[a.hex(), b.hex(), c.hex()] for a, b, c in sorted(led.dump())] if code == 0 else None})


w, txs, _ = scenarios.burn_world(); add_batch("burn", w, txs)
bad = bytearray(txs[0]); bad[-40] ^= 1; add_batch("burn_bad_signature", w, [bytes(bad)])
w, txs, _ = scenarios.realistic_world(); add_batch("realistic", w, txs); add_batch("realistic_out_of_order", w, txs[::-1])
w, d, _ = scenarios.multisig_world(); add_batch("multisig_setup_and_spend", w, [d["setup"], d["spend"]]); add_batch("multisig_wrong_threshold", w, [d["setup"], d["spend_one"]])
json.dump(out, open(os.path.join(HERE, "vectors.json"), "w"))
print("wrote", os.path.join(HERE, "vectors.json"), os.path.getsize(os.path.join(HERE, "vectors.json")), "bytes")
