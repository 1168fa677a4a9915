#!/usr/bin/env python3
"""bench.py -- verified TX/s on a 10k-transfer batch (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--shape a1k1] [--txs 10000] [--secondary full|min|off]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU path (oracle port of the reference; the Rust crate cannot run here)

A step = one verification of one batch of T synthetic transfer transactions per GPU.  At N > 1 the batch holds N x T
transactions; every rank holds all of it and verifies its own contiguous shard (weak scaling, the headline), and the same
machinery verifies ONE T-transaction batch cut N ways (`strong`).  Inside the `value` event pair at N > 1: the shard's kernels,
the rank's record (verdict word + partial sigma / range MSM encodings) built on the device, the NCCL all-gather of the
records, and the joint decision kernel (sum of the partial encodings, identity tests, reference precedence).
`value` times the device with the batch resident in HBM; `e2e` times the reference-facing call (host header walk + state
resolution + H2D + kernels + D2H + state update, or host Merlin transcripts in the north_star split) from host buffers.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# Several batches in flight means dozens of CUDA streams per process (six per context) plus NCCL's; with the default of 8
# hardware queues they alias, and a stream queued behind an all-gather that waits for a slower rank stalls with it (measured at
# 4 GPUs, 6 batches in flight: 60-84 ms per device phase instead of 6).  Must be set before CUDA initialises.
# (At one and two GPUs the default 8 measured better end to end -- 2.85-3.07 against 2.43-2.69 M TX/s at one, 5.3-5.5 against
# 4.6-4.7 M at two -- although the device-only ceiling of six contexts is 5 % higher with 32: fewer queues keep the batches'
# completions in order.  From four ranks on, with fewer host cores per rank, the stall above is what decides.)
if int(os.environ.get("WORLD_SIZE", "1")) >= 4:
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--txs", type=int, default=10000)
    ap.add_argument("--shape", default="a1k1")
    ap.add_argument("--cpu-sample", type=int, default=0, help="transactions per host thread in the CPU baseline (default: the whole batch in the cpu_baseline leg, 2,500 per step in the reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mint-cache", default=None, help="development only: a pickle (tools/r02_mint_cache.py) holding the 8 x 10k a1k1 batch minted with the same seeds; skips the minting, the oracle prefix check and the CPU baseline")
    ap.add_argument("--secondary", default="full", choices=["full", "min", "off"], help="BASELINE.json's other configs (N = 1 only): full = other shapes, one-sender chain, mixed batch with reject paths, 16x255, MSM sweep, ciphertext updates; min = MSM 2^20 + ciphertext updates")
    ap.add_argument("--no-secondary", action="store_true", help="same as --secondary off")
    ap.add_argument("--time-limit", type=float, default=160.0, help="wall-clock budget of the whole run in seconds: a secondary configuration (BASELINE.json's other configs, each minted by the CPU prover first) is skipped -- and listed as skipped -- when its estimated cost no longer fits; --time-limit 900 runs them all")
    ap.add_argument("--mixed-txs", type=int, default=100000, help="size of the mixed (config 5) batch")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling measurement (one T-transaction batch cut N ways) at N > 1")
    ap.add_argument("--no-key-index", action="store_true", help="sharded runs: find cross-shard dependencies by scanning the earlier shards' bytes instead of the batch's key-digest index")
    ap.add_argument("--no-bind", action="store_true", help="multi-GPU runs: leave the ranks floating over all host cores instead of binding each to its own cores on its GPU's NUMA node")
    ap.add_argument("--force-inflight", action="store_true", help="use --inflight as given instead of capping it at the rank's cores minus one")
    ap.add_argument("--inflight", type=int, default=6, help="batches in flight per GPU in the end-to-end measurement (one context + host thread each)")
    ap.add_argument("--fiat-shamir", default="fast", choices=["fast", "device", "host"], help="where the Merlin transcripts run (host = north_star split; device = SURVEY 8 f.1)")
    return ap.parse_args()


def shape_ak(shape):
    a, k = shape[1:].split("k")
    return int(a), int(k)


def canonical_lp_per_tx(a, k):
    """algorithmic limb products of one transaction with a assets and k transfers"""
    m = 1
    while m < a + k:
        m *= 2
    lg = 6 + m.bit_length() - 1
    return (5 + 6 * a + 9 * k + 2 * lg) * 12632 + (2 + 2 * a + 2 * k) * 12688 + ((7 * a + 8 * k) + (4 + 2 * lg + m)) * 8064 + 8 * 64 * m * 136 + 160000


def a_msm(n):
    """canonical work of an n-point MSM (16 windows), limb products"""
    return 8064.0 * n + 6.04e8


class ClockSampler(threading.Thread):
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        smax = max((int(float(r[1])) for r in self.rows if r), default=0)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline(batch, threads, sample):
    """the reference's own multi-thread model (benches/tx.rs:252-343): `threads` independent verify_batch calls, each over
    `sample` transactions of the workload, on the oracle port (kind = "port": curve25519-dalek itself cannot run here)."""
    sub = batch.slice(min(sample, batch.n))
    t, rc = sub.verify_timed(threads)
    assert rc == 0, rc
    return {"value": threads * sub.n / t, "unit": "TX/s", "cores": threads, "kind": "port",
            "sample": f"{threads} threads x verify_batch({sub.n} tx) of the same workload, scalar 64-bit backend, no SIMD; CPU restatement, not curve25519-dalek",
            "seconds": t, "readme_figure_tx_s_per_thread": 2500}


class Dev:
    """ctypes plumbing shared by the measurements"""

    def __init__(self, lib):
        self.lib = lib
        lib.xhe_batch_run.argtypes = [C.c_void_p]; lib.xhe_batch_run.restype = C.c_int32
        lib.xhe_batch_h2d_bytes.restype = C.c_size_t; lib.xhe_batch_h2d_bytes.argtypes = [C.c_void_p]
        lib.xhe_batch_d2h_bytes.restype = C.c_size_t; lib.xhe_batch_d2h_bytes.argtypes = [C.c_void_p]
        lib.xhe_ctx_timing.argtypes = [C.c_void_p, C.c_int]
        lib.xhe_ctx_timing_read.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int]
        lib.xhe_ctx_set_serial.argtypes = [C.c_void_p, C.c_int]
        lib.xhe_ctx_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
        lib.xhe_msm_workspace_bytes.restype = C.c_size_t; lib.xhe_msm_workspace_bytes.argtypes = [C.c_void_p, C.c_size_t]
        lib.xhe_msm_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.xhe_ct_update_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.xhe_ct_update_resident_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]


def timed_runs(torch, stream, fn, iters, warm, flush=None):
    """mean ms of fn() over `iters` launches on `stream` (CUDA events on that stream, L2 flushed outside the event pairs)"""
    with torch.cuda.stream(stream):
        for _ in range(warm):
            assert fn() == 0
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            if flush is not None:
                flush.fill_(i & 0xFF)
            e0.record(stream)
            assert fn() == 0
            e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def measure_config(name, torch, xhe, verifier, dev, stream, flush, batch_blobs, records, multisig, party_capacity, steps, e2e_steps, rejects, host_threads, canonical_lp=None, peak=None):
    """One BASELINE configuration end to end on one GPU: accept verdict, device-resident `value`, host-buffer `e2e` (single call),
    and for every (label, tampered blobs, expected verdict) in `rejects` the verdict and the reject-path throughput."""
    ctx = xhe.Ctx(torch.cuda.current_device(), party_capacity=party_capacity)
    ctx.set_stream(stream.cuda_stream)
    try:
        def ledger():
            led = verifier.Ledger(); led.import_records(records)
            for pk, signers, th in multisig:
                led.set_multisig(pk, signers, th)
            return led
        n = len(batch_blobs)
        prepared = verifier.prepare_blobs_pinned(batch_blobs)
        out = {"txs": n}
        for w in range(2):
            code, idx, tm = verifier.verify_batch(ctx, None, ledger(), seed=b"sec-warm%d" % w, threads=host_threads, prepared=prepared, fiat_shamir="fast")
            assert (code, idx) == (0, -1), (name, code, idx)
        out["accepted"] = True; out["path"] = "fast (device transcripts + device layout)" if tm["fast_path"] else ("exact, Merlin transcripts on the host threads: few transactions with very long transcripts (verify_batch picks north_star's split for them)" if name == "16x255" else "exact (host state walk and tables, device transcripts): multisig transactions / accounts in the batch")
        ms = timed_runs(torch, stream, lambda: dev.lib.xhe_batch_run(ctx.p), steps, 2, flush)
        out["value"] = {"value": n / ms * 1e3, "unit": "TX/s", "ms_per_step": ms}
        if canonical_lp and peak:
            out["value"]["step_frac"] = canonical_lp / (ms * 1e-3) / peak
        t = 0.0
        for s in range(e2e_steps):
            led = ledger(); flush.fill_(s); torch.cuda.synchronize()
            t0 = time.perf_counter()
            code, idx, tm = verifier.verify_batch(ctx, None, led, seed=b"sec%d" % s, threads=host_threads, prepared=prepared, fiat_shamir="fast")
            t += time.perf_counter() - t0
            assert (code, idx) == (0, -1)
        out["e2e_single_call"] = {"value": n * e2e_steps / t, "unit": "TX/s", "ms_per_step": 1e3 * t / e2e_steps, "h2d_bytes_per_step": int(dev.lib.xhe_batch_h2d_bytes(ctx.p)), "d2h_bytes_per_step": int(dev.lib.xhe_batch_d2h_bytes(ctx.p)),
                                  "phases_ms": {k: round(v, 3) for k, v in tm.items() if k.endswith("_ms")}}
        rj = {}
        for label, blobs, want in rejects:
            pb = verifier.prepare_blobs(blobs)
            verifier.verify_batch(ctx, None, ledger(), seed=b"rj-warm", threads=host_threads, prepared=pb, fiat_shamir="fast")
            t, reps = 0.0, 2
            for s in range(reps):
                led = ledger(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                code, idx, tm = verifier.verify_batch(ctx, None, led, seed=b"rj%d" % s, threads=host_threads, prepared=pb, fiat_shamir="fast")
                t += time.perf_counter() - t0
            assert (code, idx) == tuple(want), (name, label, code, idx, want)
            rj[label] = {"verdict": [code, idx], "matches_expected": True, "reject_path_tx_per_s": len(blobs) * reps / t, "ms": 1e3 * t / reps, "decided_by": "fast path + exact verdict of one transaction" if tm["fast_path"] else "exact path"}
        if rj:
            out["rejects"] = rj
            out["reject_over_accept_time"] = max(v["ms"] for v in rj.values()) / out["e2e_single_call"]["ms_per_step"]
        return out
    finally:
        ctx.close()


def tamper_classes(oracle, batch, seed, victim, a, k, resignable=True):
    """(label, blobs, expected verdict) per tamper class on transaction `victim` of a minted batch; expectations follow from the
    construction and are cross-checked against the oracle on that ONE transaction (accounts of a minted batch are independent)"""
    SIG, DECOMP, GENERIC, RANGE, TRANSCRIPT, NONCE = 1, 2, 5, 6, 7, 9
    blobs = list(batch.blobs)

    def mut(off, bit=1):
        t = bytearray(blobs[victim]); t[off] ^= bit; return bytes(t)
    cases = [("bad_signature", mut(len(blobs[victim]) - 1), (SIG, victim)), ("bad_nonce", mut(56), (NONCE, victim))]
    if resignable:
        kp = oracle.minted_keypair(seed, victim); rng = oracle.Rng(b"bench-tamper")
        rp0 = 64 + 324 * k
        cases += [("bad_validity_proof_resigned", oracle.resign(mut(64 + 160 + 128), kp, rng), (GENERIC, -1)),
                  ("bad_range_proof_resigned", oracle.resign(mut(rp0 + 128), kp, rng), (RANGE, -1)),
                  ("identity_Y0_resigned", oracle.resign(blobs[victim][:64 + 160] + bytes(32) + blobs[victim][64 + 160 + 32:], kp, rng), (TRANSCRIPT, victim)),
                  ("non_canonical_point_resigned", oracle.resign(blobs[victim][:64 + 64] + b"\xff" * 32 + blobs[victim][64 + 96:], kp, rng), (DECOMP, victim))]
    out = []
    for label, bad, want in cases:
        one = oracle.verify_batch([bad], batch.ledger())
        assert one == (want[0], 0 if want[1] >= 0 else -1), (label, one, want)      # the oracle agrees on the class
        out.append((label, blobs[:victim] + [bad] + blobs[victim + 1:], want))
    return out


def msm_sweep(torch, lib, ctx, stream, oracle, logs, cpu_logs, peak):
    """config 2: MSM over resident decompressed points, 2^16..2^22; the CPU Pippenger (oracle: dalek's algorithm choices, one
    thread) is timed beside it on the same distribution at the sizes in cpu_logs"""
    out = {}
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    for logn in logs:
        n = 1 << logn
        with torch.cuda.stream(stream):
            uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
            enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); niels = torch.empty((n, 24), dtype=torch.int32, device="cuda"); ok = torch.empty((n,), dtype=torch.uint8, device="cuda")
            sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g); sc[:, 31] &= 0x0F
            assert lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr()) == 0
            assert lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, None, niels.data_ptr(), ok.data_ptr()) == 0
        wsb = lib.xhe_msm_workspace_bytes(ctx.p, n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); res = torch.zeros(64, dtype=torch.uint8, device="cuda")
        ms = timed_runs(torch, stream, lambda: lib.xhe_msm_dev(ctx.p, sc.data_ptr(), niels.data_ptr(), n, ws.data_ptr(), wsb, res.data_ptr(), res.data_ptr() + 32), 5, 3)
        c, W = ctx.msm_plan(n)
        row = {"points": n, "ms": ms, "points_per_s": n / ms * 1e3, "alg_TLP_s": a_msm(n) / ms / 1e9, "msm_frac": a_msm(n) / (ms * 1e-3) / peak, "c": c, "windows": W}
        if logn in cpu_logs:
            s_host, p_host = bytes(sc.cpu().numpy()), bytes(enc.cpu().numpy())
            t, enc_cpu = oracle.msm_timed(s_host, p_host)
            row["cpu_pippenger"] = {"seconds": t, "points_per_s": n / t, "threads": 1, "kind": "port (dalek's Straus / Pippenger w = 6,7,8 choices, 64-bit limbs)"}
            row["bit_exact_vs_cpu"] = bytes(res[:32].cpu().numpy()) == enc_cpu
            assert row["bit_exact_vs_cpu"], "MSM encoding differs from the CPU oracle at 2^%d" % logn
        out["2^%d" % logn] = row
        del uni, enc, niels, ok, sc, ws
        torch.cuda.empty_cache()
    return out


def ct_update_metrics(torch, lib, ctx, stream, hbm_peak_gbs):
    """config 4: batched Twisted-ElGamal balance update over 1 M accounts, resident (HBM-bound) and compressed I/O (integer-bound)"""
    out = {}
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    n = 1 << 20
    with torch.cuda.stream(stream):
        uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
        enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        assert lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr()) == 0
    na_c = 1 << 19
    subc = torch.randint(0, 2, (na_c,), dtype=torch.uint8, device="cuda", generator=g)
    outb = torch.empty((na_c, 64), dtype=torch.uint8, device="cuda"); okb = torch.empty((na_c,), dtype=torch.uint8, device="cuda")
    bal = enc[: 2 * na_c].reshape(na_c, 64); delta = enc.flip(0)[: 2 * na_c].reshape(na_c, 64).contiguous()
    ms_c = timed_runs(torch, stream, lambda: lib.xhe_ct_update_dev(ctx.p, bal.data_ptr(), delta.data_ptr(), subc.data_ptr(), na_c, outb.data_ptr(), okb.data_ptr()), 5, 3)
    out["ct_update_compressed_512k"] = {"accounts": na_c, "ms": ms_c, "accounts_per_s": na_c / ms_c * 1e3, "alg_TLP_s": na_c * 76000.0 / ms_c / 1e9}
    del uni
    na = 1 << 20
    balr = torch.randint(0, 2**31 - 1, (4, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
    dn = torch.randint(0, 2**31 - 1, (3, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
    sub = torch.randint(0, 2, (na,), dtype=torch.uint8, device="cuda", generator=g)
    ms = timed_runs(torch, stream, lambda: lib.xhe_ct_update_resident_dev(ctx.p, balr.data_ptr(), dn.data_ptr(), sub.data_ptr(), na), 10, 3)
    bytes_per_account = 2 * (128 + 128 + 96)       # two points per account: extended balance read + write, affine-Niels delta read
    gbs = na * bytes_per_account / ms / 1e6
    out["ct_update_resident_1M"] = {"accounts": na, "ms": ms, "accounts_per_s": na / ms * 1e3,
                                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak_gbs, "unit": "GB/s", "frac": gbs / hbm_peak_gbs, "traffic": 688.9e6,
                                                 "traffic_note": "dram__bytes_read.sum (470.8 MB = 2 M points x 224 B, exactly the operands) + dram__bytes_write.sum (218.1 MB; the rest of the 268 MB written is still in L2 when the kernel ends) of one launch, ncu, profiles/r01_ct_resident_traffic.csv: no re-reads",
                                                 "algorithmic_bytes_per_account": bytes_per_account, "working_set_mb": na * bytes_per_account / 1e6}}
    return out


def main():
    # exactly one JSON line may reach stdout: libraries (NCCL's version banner, torchrun) write there too, so everything else
    # is routed to stderr and the line is written to the saved descriptor at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    t_bench0 = time.time()
    args = parse_args()
    if args.no_secondary:
        args.secondary = "off"
    a, k = shape_ak(args.shape)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = len(os.sched_getaffinity(0))
    workload = f"batch verify {args.txs} transfer TXs per GPU, shape {args.shape} (assets a, transfers k), 64-bit aggregated range proofs + sigma proofs + signatures"
    config = {"workload": workload, "txs_per_gpu": args.txs, "shape": args.shape, "parallelism": f"tx-shard x{world}: one batch of {world * args.txs} transactions, rank r verifies shard r" if world > 1 else "single",
              "l2": "L2 flushed (256 MiB write) between timed steps"}

    import oracle   # checker / CPU baseline / test-vector minting only (never on the product path)

    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(args.cpu_sample or 2500, args.txs)
        batch = oracle.mint_transfers(77, sample, a, k, threads=ncpu)
        for _ in range(max(args.warmup, 1)):
            batch.verify_timed(ncpu)
        tot = 0.0
        for _ in range(args.steps):
            t, rc = batch.verify_timed(ncpu); assert rc == 0
            tot += t
        v = ncpu * sample * args.steps / tot
        line = {"metric": "verified TX/s (10k-transfer batch)", "value": v, "unit": "TX/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic (minted by the oracle prover)",
                "config": config, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "TX/s", "cores": ncpu, "kind": "port",
                                 "sample": f"{ncpu} threads x verify_batch({sample} tx) per step; CPU restatement of the reference path (Rust toolchain and crates absent), scalar 64-bit backend"},
                "e2e": {"value": v, "unit": "TX/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    import torch
    import xelis_he_b200 as xhe
    from xelis_he_b200 import verifier
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1 and not args.no_bind:
        # one process per GPU: each rank keeps to its own cores on its GPU's NUMA node (before anything is allocated)
        from xelis_he_b200.distributed import bind_rank_to_local_cores
        bound = bind_rank_to_local_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", world)), verbose=True)
        ncpu_rank = len(bound)
    else:
        ncpu_rank = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- synthetic workload: every rank mints ITS OWN shard (distinct accounts), then the ranks exchange the bytes so that each
    # holds the whole batch -- as every node of a network holds the whole block
    t_mint = time.time()
    mint_threads = max(1, ncpu // max(1, min(world, 8)))
    cached = None
    if args.mint_cache and os.path.exists(args.mint_cache) and args.shape == "a1k1" and args.txs == 10000 and world <= 8:
        import pickle
        cb, cr = pickle.load(open(args.mint_cache, "rb")); per = len(cr) // 8
        cached = (cb[rank * args.txs:(rank + 1) * args.txs], cr[rank * per:(rank + 1) * per])
        args.no_cpu_baseline = True; args.secondary = "off"

        class _B:
            blobs = cached[0]
        batch = _B()
        my_records = cached[1]
    else:
        batch = oracle.mint_transfers(77 + rank, args.txs, a, k, threads=mint_threads)
        my_records = batch.ledger().dump()
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (batch.blobs, my_records))
        blobs = [b for part in gathered for b in part[0]]
        records = [r for part in gathered for r in part[1]]
        strong_blobs, strong_records = gathered[0]            # the strong-scaling batch: rank 0's T transactions, cut N ways
    else:
        blobs, records = batch.blobs, my_records
    t_mint = time.time() - t_mint
    n_total = len(blobs)
    m = 1
    while m < a + k:
        m *= 2
    ctx = xhe.Ctx(local, party_capacity=max(m, 2))
    # every context works on its own non-blocking stream, so batches in flight (and the NCCL exchange) overlap on the device
    streams = [torch.cuda.Stream()]
    ctx.set_stream(streams[0].cuda_stream)
    lib = ctx.lib
    dev = Dev(lib)
    ledger0 = verifier.Ledger(); ledger0.import_records(records)
    # the batch in one page-locked buffer, device layout: uploaded in place (zero-copy input).  With it, the batch's key-digest
    # index (8-byte digests of the balances each transaction moves): both are built where the transactions are received and
    # framed, before the clock -- like the reference's deserialisation into `Transaction` values
    prepared = verifier.prepare_blobs_pinned(blobs, index=world > 1 and not args.no_key_index)
    host_threads = ncpu_rank or max(1, ncpu // max(1, min(world, 8)))

    from xelis_he_b200 import distributed as xd

    fs_mode = [args.fiat_shamir]

    def e2e_step(seed, prep=None, led0=None):
        led = (led0 or ledger0).clone()
        t0 = time.perf_counter()
        if dist:     # sharded batch: local partial verification + 80-byte all-gather over NCCL + joint decision on every rank
            code, idx, tm = xd.verify_batch_distributed(ctx, None, led, seed=seed + b"r%d" % rank, threads=host_threads, prepared=prep or prepared, commit=True, fiat_shamir=fs_mode[0])
        else:
            code, idx, tm = verifier.verify_batch(ctx, None, led, seed=seed, threads=host_threads, prepared=prep or prepared, fiat_shamir=fs_mode[0])
        return time.perf_counter() - t0, code, idx, tm

    # ---- correctness gate + warm-up (also leaves the batch resident in HBM for the device-only timing)
    for w in range(max(args.warmup, 3)):
        dt, code, idx, tm = e2e_step(b"warm%d" % w)
        assert (code, idx) == (0, -1), (code, idx)
    if args.txs <= 20000 and cached is None:
        sl = min(64, args.txs)
        assert oracle.verify_batch(batch.blobs[:sl], batch.slice(sl).ledger()) == (0, -1)      # the oracle agrees on a prefix of this rank's shard

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(max(args.warmup, 3)):
        assert lib.xhe_batch_run(ctx.p) == 0
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    rec_local = torch.zeros(80, dtype=torch.uint8, device="cuda"); rec_all = torch.zeros(80 * world, dtype=torch.uint8, device="cuda"); decision = torch.zeros(16, dtype=torch.uint8, device="cuda")
    ts = streams[0]

    def device_steps(nsteps):
        """nsteps timed steps of the resident shard: kernels + (N > 1) record, all-gather, joint decision -- all on the ctx stream"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        for s in range(nsteps):
            with torch.cuda.stream(ts):
                flush.fill_(s & 0xFF)              # evict the previous step's working set from L2 (outside the event pair)
                ev[s][0].record(ts)
                assert lib.xhe_batch_run(ctx.p) == 0
                if dist:
                    assert lib.xhe_batch_record_dev(ctx.p, rec_local.data_ptr()) == 0
                    dist.all_gather_into_tensor(rec_all, rec_local)      # the per-batch exchange: (verdict, partial sigma / range encodings) of every rank
                    assert lib.xhe_shard_decide_dev(ctx.p, rec_all.data_ptr(), world, decision.data_ptr()) == 0
                ev[s][1].record(ts)
        torch.cuda.synchronize()
        if dist:
            d = bytes(decision.cpu().numpy())
            assert int.from_bytes(d[:4], "little", signed=True) == 0 and int.from_bytes(d[8:16], "little", signed=True) == -1, "joint decision of the timed steps is not Accept"
        return sum(e0.elapsed_time(e1) for e0, e1 in ev)

    # ---- host->device rate with every rank uploading at once (what the end-to-end figure can get out of this box)
    h2d_probe = None
    if dist:
        pb_h = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True); pb_d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        pb_d.copy_(pb_h, non_blocking=True); barrier()
        p0 = torch.cuda.Event(enable_timing=True); p1 = torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(4):
            pb_d.copy_(pb_h, non_blocking=True)
        p1.record(); torch.cuda.synchronize()
        tp = torch.tensor([4 * 64 * 2**20 / (p0.elapsed_time(p1) * 1e-3) / 1e9], device="cuda", dtype=torch.float64)
        allp = [torch.zeros_like(tp) for _ in range(world)]; dist.all_gather(allp, tp)
        h2d_probe = {"per_rank_GBs": [round(float(x), 2) for x in allp], "aggregate_GBs": round(sum(float(x) for x in allp), 1), "what": "256 MiB pinned -> device per rank, all ranks at once"}
        del pb_h, pb_d
        barrier()
    # ---- timed region 1: device kernels on the resident batch (value)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    launches0 = ctx.launches
    dev_ms = device_steps(args.steps)
    launches = ctx.launches - launches0                                     # ours only: NCCL's all-gather kernel is not counted
    # isolated per-kernel durations for the roofline: the same steps with the pipelines serialised on one stream and
    # CUDA events around the main kernels (timers under stream concurrency would include the overlapped neighbours)
    lib.xhe_ctx_set_serial(ctx.p, 1); lib.xhe_ctx_timing(ctx.p, 1)
    with torch.cuda.stream(ts):
        for s in range(args.steps):
            flush.fill_(s & 0xFF)
            assert lib.xhe_batch_run(ctx.p) == 0
    torch.cuda.synchronize()
    names = (C.c_char_p * 16)(); kms = (C.c_double * 16)(); kl = (C.c_uint64 * 16)(); ku = (C.c_double * 16)()
    nk = lib.xhe_ctx_timing_read(ctx.p, names, kms, kl, ku, 16)
    kernels = {names[i].decode(): {"ms_per_step": kms[i] / args.steps, "launches": int(kl[i]), "alg_lp_per_step": ku[i] / args.steps} for i in range(nk)}
    lib.xhe_ctx_timing(ctx.p, 0); lib.xhe_ctx_set_serial(ctx.p, 0)
    # where each kernel sits inside one concurrent step (stream pipelines): start/end in ms from the step start
    lib.xhe_ctx_timing(ctx.p, 1)
    with torch.cuda.stream(ts):
        flush.fill_(1)
        assert lib.xhe_batch_run(ctx.p) == 0
    torch.cuda.synchronize()
    lib.xhe_ctx_timing_read(ctx.p, names, kms, kl, ku, 16)
    tn = (C.c_char_p * 64)(); t0s = (C.c_float * 64)(); t1s = (C.c_float * 64)()
    nt = lib.xhe_ctx_timeline(ctx.p, tn, t0s, t1s, 64)
    timeline = sorted([[tn[i].decode(), round(t0s[i], 3), round(t1s[i], 3)] for i in range(max(nt, 0))], key=lambda r: r[1])
    lib.xhe_ctx_timing(ctx.p, 0)
    barrier()
    # ---- timed region 2: end to end through the host API, host buffers in, verdict out (e2e).
    # (a) one call at a time (latency); (b) several batches in flight on as many contexts of the same GPU, so the host phase of
    # one batch overlaps the device phase of the others -- what a node verifying a stream of batches does.
    single_s = 0.0; phases = {}
    for s in range(args.steps):
        flush.fill_(s & 0xFF); torch.cuda.synchronize()
        dt, code, idx, tm = e2e_step(b"step%d" % s)
        assert (code, idx) == (0, -1)
        single_s += dt
        for kk, vv in tm.items():
            phases[kk] = phases.get(kk, 0.0) + vv / args.steps
    barrier()
    # (a') the same call on a device-resident state: balances are read from and committed to the device table
    resident_state = None
    if not dist and args.fiat_shamir == "fast":
        dls = verifier.DeviceLedgerState(ctx, 2 * len(records) + 16); dls.import_records(records); dls.snapshot()
        for w in range(2):
            code, idx, tm = verifier.verify_batch(ctx, None, dls, seed=b"dls-warm", threads=host_threads, prepared=prepared, fiat_shamir="fast"); assert (code, idx) == (0, -1); dls.restore()
        t_dls, ph = 0.0, {}
        for s in range(args.steps):
            flush.fill_(s & 0xFF); torch.cuda.synchronize()
            t0 = time.perf_counter()
            code, idx, tm = verifier.verify_batch(ctx, None, dls, seed=b"dls%d" % s, threads=host_threads, prepared=prepared, fiat_shamir="fast")
            ctx.sync()                                   # the commit kernel is part of the call
            t_dls += time.perf_counter() - t0
            assert (code, idx) == (0, -1) and tm["fast_path"]
            for kk, vv in tm.items():
                ph[kk] = ph.get(kk, 0.0) + vv / args.steps
            dls.restore()
        resident_state = {"single_call": {"value": args.txs * args.steps / t_dls, "ms_per_step": 1e3 * t_dls / args.steps}, "h2d_bytes_per_step": int(lib.xhe_batch_h2d_bytes(ctx.p)), "d2h_bytes_per_step": 512,
                          "phases_ms": {kk: round(vv, 3) for kk, vv in ph.items() if kk.endswith("_ms")}, "note": "BlockchainVerificationState backed by xhe_ledger: no balance crosses the bus; the table is restored from a device-side snapshot between steps (outside the clock)"}
        dls.close()
        e2e_step(b"restore-host-ledger")
    barrier()
    pipelined = args.fiat_shamir != "host" and args.inflight > 1
    # batches in flight: one host thread each, so no more than the rank's cores minus one for the decision / commit threads
    # (8 GPUs on a 32-core box = 4 cores per rank: 3 in flight measured 9.3 M TX/s against 7.8 M with 6)
    nfl = (args.inflight if args.force_inflight else min(args.inflight, max(2, host_threads - 1))) if pipelined else 1
    workers = [ctx] + [xhe.Ctx(local, party_capacity=max(m, 2)) for _ in range(nfl - 1)]
    for c in workers[1:]:
        streams.append(torch.cuda.Stream()); c.set_stream(streams[-1].cuda_stream)
    # sharded + pipelined: the cross-rank decisions (ordered NCCL all-gathers, sum of the partial encodings, commit of the
    # held-back balance updates) run on one thread per process, off the verification threads
    decider = None
    if dist and pipelined:
        dec_ctx = xhe.Ctx(local, party_capacity=2)
        streams_keep = torch.cuda.Stream(priority=-1); dec_ctx.set_stream(streams_keep.cuda_stream)      # its tiny kernels must not queue behind the batches
        decider = xd.AsyncDecider(dec_ctx, None, torch.device("cuda", local))
    wthreads = max(1, host_threads // nfl)
    lo_w, hi_w = xd.shard_bounds(n_total, rank, world)

    def worker(widx, nsteps, out, ledgers=None):
        c = workers[widx]
        torch.cuda.set_device(local)
        for s in range(nsteps):
            led = ledgers[s] if ledgers else ledger0.clone()      # fresh state per step (cloned before the clock starts)
            if dist:
                seq = seq_base[0] + s * nfl + widx
                code, idx, s_enc, r_enc, tmw = verifier.verify_batch_shard(c, None, led, lo_w, hi_w, seed=b"p%d-%d-%d" % (widx, s, rank), threads=wthreads, prepared=prepared, fiat_shamir=args.fiat_shamir)
                decider.submit(seq, xd.pack_local(code, idx, 0, s_enc, r_enc), verifier.take_pending(c), led)
                pipe_phases.append(tmw)
                out.append(seq)
                continue
            else:
                code, idx, tmw = verifier.verify_batch(c, None, led, seed=b"p%d-%d" % (widx, s), threads=wthreads, prepared=prepared, fiat_shamir=args.fiat_shamir)
                pipe_phases.append(tmw)
            out.append((code, idx))
    seq_base = [0]
    decider_stats = None
    pipe_phases = []
    if pipelined:
        # warm-up of the extra contexts: every slot runs one batch (same sequence numbering on all ranks)
        wth = [threading.Thread(target=worker, args=(w, 1, [])) for w in range(nfl)]
        for t_ in wth:
            t_.start()
        for t_ in wth:
            t_.join()
        seq_base[0] = nfl
        if decider:
            assert all(v == (0, -1) for v in decider.drain(nfl).values())
        barrier()
        pipe_phases.clear()
        outs = [[] for _ in range(nfl)]
        counts = [args.steps // nfl + (1 if w < args.steps % nfl else 0) for w in range(nfl)]
        fresh = [[ledger0.clone() for _ in range(counts[w])] for w in range(nfl)]
        th = [threading.Thread(target=worker, args=(w, counts[w], outs[w], fresh[w])) for w in range(nfl)]
        t0 = time.perf_counter()
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        verdicts = decider.drain(nfl + args.steps) if decider else None      # every batch decided and committed
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        flat = [o for oo in outs for o in oo]
        if decider:
            assert len(flat) == args.steps and all(verdicts[q] == (0, -1) for q in flat)
            decider_stats = {kk: round(vv / max(1, decider.stats["n"]), 3) if kk.endswith("_ms") else vv for kk, vv in decider.stats.items()}
            decider.close()
        else:
            assert all(o == (0, -1) for o in flat) and len(flat) == args.steps
        # device-side ceiling of that pipeline: every context re-runs its resident batch, all streams in flight at once
        rounds = max(2, args.steps // nfl)
        torch.cuda.synchronize()
        c0 = torch.cuda.Event(enable_timing=True); c0.record(streams[0])
        for _ in range(rounds):
            for c in workers:
                assert lib.xhe_batch_run(c.p) == 0
        cends = []
        for st_ in streams[:nfl]:
            e_ = torch.cuda.Event(enable_timing=True); e_.record(st_); cends.append(e_)
        torch.cuda.synchronize()
        conc_ms = max(c0.elapsed_time(e_) for e_ in cends)
        concurrent_value = args.txs * rounds * nfl / (conc_ms * 1e-3)
    else:
        e2e_s = single_s
        concurrent_value = None
    barrier()
    # the other Fiat-Shamir placement, for the record (3 steps)
    other = "host" if args.fiat_shamir != "host" else "fast"
    fs_mode[0] = other
    e2e_step(b"warm-other"); t_other = 0.0
    for s in range(3):
        dt, code, idx, _ = e2e_step(b"other%d" % s); assert (code, idx) == (0, -1); t_other += dt
    fs_mode[0] = args.fiat_shamir
    e2e_step(b"restore")      # leave the primary mode's batch resident
    barrier()
    h2d, d2h = lib.xhe_batch_h2d_bytes(ctx.p), lib.xhe_batch_d2h_bytes(ctx.p)

    # ---- strong scaling: ONE batch of T transactions (rank 0's), cut N ways -- T / N transactions per GPU
    strong = None
    if dist and not args.no_strong:
        sl0 = verifier.Ledger(); sl0.import_records(strong_records)
        sprep = verifier.prepare_blobs_pinned(strong_blobs)
        for w in range(3):
            dt, code, idx, tm = e2e_step(b"strong-warm%d" % w, sprep, sl0)
            assert (code, idx) == (0, -1)
        barrier()
        s_dev_ms = device_steps(args.steps)
        barrier()
        s_e2e = 0.0
        for s in range(args.steps):
            torch.cuda.synchronize()
            dt, code, idx, tm = e2e_step(b"strong%d" % s, sprep, sl0)
            assert (code, idx) == (0, -1)
            s_e2e += dt
        barrier()
        t = torch.tensor([s_dev_ms, s_e2e * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s_dev_ms, s_e2e_ms = t.tolist()
        strong = {"scaling": "strong", "txs_total": len(strong_blobs), "txs_per_gpu": len(strong_blobs) // world,
                  "value": len(strong_blobs) * args.steps / (s_dev_ms * 1e-3), "ms_per_step": s_dev_ms / args.steps, "unit": "TX/s",
                  "e2e_single_call": {"value": len(strong_blobs) * args.steps / (s_e2e_ms * 1e-3), "ms_per_step": s_e2e_ms / args.steps},
                  "note": "one 10k batch cut N ways: kernels of the shard + record + NCCL all-gather + joint decision inside the event pair; e2e = verify_batch_distributed, one call at a time"}
        e2e_step(b"restore2")
        barrier()
    clocks = sampler.summary()

    by_rank = None
    if dist:
        mine = {"rank": rank, "single_call": {kk: round(vv, 3) for kk, vv in phases.items() if kk.endswith("_ms")},
                "pipelined": {kk: round(sum(t_[kk] for t_ in pipe_phases) / len(pipe_phases), 3) for kk in pipe_phases[0] if kk.endswith("_ms")} if pipe_phases else None}
        by_rank = [None] * world
        dist.all_gather_object(by_rank, mine)
    if dist:
        t = torch.tensor([dev_ms, e2e_s * 1e3, single_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_ms_max, single_ms_max = t.tolist()
    else:
        dev_ms_max, e2e_ms_max, single_ms_max = dev_ms, e2e_s * 1e3, single_s * 1e3

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    total_tx = args.txs * world * args.steps
    value = total_tx / (dev_ms_max * 1e-3)
    e2e = total_tx / (e2e_ms_max * 1e-3)
    # ---- roofline (integer-multiply pipe; tensor cores unused by design).  `kernel` = the kernel with the largest isolated time
    # of the step; step_frac = canonical work of the whole step / step time / peak; per_kernel_isolated lists every timed kernel.
    peak_wide = ctx.int_peak(2); peak_chain = ctx.int_peak(3)
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6553.9          # the figure MEASURED_PEAKS.json held when this was written
    leaf = {n: v for n, v in kernels.items() if not n.startswith("msm_")}      # msm_* timers wrap several kernels
    dom = max(leaf, key=lambda n: leaf[n]["ms_per_step"])
    ach = leaf[dom]["alg_lp_per_step"] / (leaf[dom]["ms_per_step"] * 1e-3)
    work = {n: {"ms": round(v["ms_per_step"], 4), "TLP_s": round(v["alg_lp_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12, 3), "frac": round(v["alg_lp_per_step"] / (v["ms_per_step"] * 1e-3) / peak_wide, 4)}
            for n, v in kernels.items() if v["alg_lp_per_step"] > 0 and v["ms_per_step"] > 0}
    step_lp = canonical_lp_per_tx(a, k) * args.txs
    step_ms = dev_ms_max / args.steps
    n_sigma = args.txs * (7 * a + 8 * k) + 2
    lgm = 6 + (m.bit_length() - 1)
    n_dyn = args.txs * (4 + 2 * lgm + m)
    msm_in_batch = {}
    # (the default step runs ONE Pippenger instance over both term sets -- msm_joint; XHE_SPLIT_MSM=1 keeps the two apart)
    for nm, npts, extra in (("msm_joint", n_sigma + n_dyn, kernels.get("msm_joint_sort", {}).get("ms_per_step", 0.0)),
                            ("msm_sigma", n_sigma, kernels.get("msm_sigma_sort", {}).get("ms_per_step", 0.0)), ("msm_range_dyn", n_dyn, 0.0)):
        if nm in kernels and kernels[nm]["ms_per_step"] > 0:
            ms_ = kernels[nm]["ms_per_step"] + extra
            msm_in_batch[nm] = {"points": npts, "ms_isolated": round(ms_, 4), "msm_frac": round(a_msm(npts) / (ms_ * 1e-3) / peak_wide, 4)}
    roofline = {"bound": "int-mul", "kernel": dom, "kernel_choice": "largest isolated time among the kernels of one step", "achieved": ach / 1e12, "peak": peak_wide / 1e12, "unit": "TLP/s (32x32->64 limb products)", "frac": ach / peak_wide,
                "step_frac": step_lp / (step_ms * 1e-3) / peak_wide, "step_canonical_TLP": step_lp / 1e12, "msm_frac_in_batch": msm_in_batch,
                "share_of_step_time_isolated": leaf[dom]["ms_per_step"] / max(1e-9, sum(v["ms_per_step"] for v in leaf.values())),
                "peak_source": "measured live: IMAD.WIDE.U32 Rd64, Ra, Rb, RZ microkernel (two vector multiplicands, both result words live; SASS checked: IMAD.WIDE only)", "peak_carry_chain": peak_chain / 1e12,
                "traffic": None, "traffic_note": "integer-bound kernel: operands are a 64-byte signature and one 32-byte key per thread (DRAM traffic of the HBM-bound kernel: secondary.ct_update_resident_1M.roofline)",
                "per_kernel_isolated": work, "canonical_units": "SURVEY.md 8d: decode 12,632 LP, encode 12,688, mixed add 504, A_msm(n) = 8,064 n + 6.04e8 (16 windows whatever window the plan picks), signature 160,000",
                "note": "every 32x32->64 form (IMAD.WIDE, IMAD.WIDE.X carry chains, IMAD.HI) issues at 4 cycles per warp instruction per SM sub-partition on sm_100a, half the rate of the 32-bit IMAD (DESIGN.md 4.1)"}
    kf = phases.get("keccak_f", -1)
    line = {"metric": "verified TX/s (10k-transfer batch)", "value": value, "unit": "TX/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "vs_baseline_note": "reference README: ~0.40 ms/TX on one CPU thread (hardware unstated)",
            "dtype": "u32 limbs (GF(2^255-19), mod l)", "data": "synthetic (valid TXs minted by the oracle prover; every rank mints its own shard, all ranks hold the whole batch)", "config": config,
            "e2e": {"value": e2e, "unit": "TX/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps, "host_threads": host_threads, "batches_in_flight": nfl,
                    "single_call": {"value": total_tx / (single_ms_max * 1e-3), "ms_per_step": single_ms_max / args.steps},
                    "phases_ms": {kk: round(vv, 3) for kk, vv in phases.items() if kk != "keccak_f"}, "host_keccak_f_per_tx": (kf / args.txs) if kf >= 0 else None,
                    "phases_ms_pipelined_mean": {kk: round(sum(t_[kk] for t_ in pipe_phases) / len(pipe_phases), 3) for kk in pipe_phases[0] if kk.endswith("_ms")} if pipe_phases else None,
                    "resident_state": resident_state, "phases_ms_by_rank": by_rank, "h2d_all_ranks_at_once": h2d_probe,
                    "key_index": ({"bytes": prepared.index_bytes, "build_ms": round(prepared.index_build_ms, 3), "built": "with the blob arena, before the clock (as the reference deserialises before verify_batch)"} if getattr(prepared, "index", None) else None), "decision_thread_ms_per_batch": decider_stats, "fiat_shamir": args.fiat_shamir,
                    "fiat_shamir_note": "fast = device transcripts + device layout (SURVEY 8 f.1 + f.2); other_mode = north_star's split (Merlin on host threads)",
                    "other_mode": {"fiat_shamir": other, "value_this_rank": args.txs * 3 / t_other}},
            "value_batches_in_flight": {"value_this_rank": concurrent_value, "unit": "TX/s", "contexts": nfl, "note": "device-resident batches of all contexts in flight at once (no L2 flush); the GPU-side ceiling of the pipelined e2e"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels_ms_per_step_isolated": {n: round(v["ms_per_step"], 4) for n, v in kernels.items()},
            "timeline_ms_one_step": timeline, "mint_seconds": round(t_mint, 1), "host_cores": ncpu}
    if world > 1:
        line["collective"] = {"what": "per batch: one ncclAllGather of 80 B per rank (verdict word + partial sigma / range MSM encodings), then the joint decision kernel (sum of the partial encodings, identity tests) -- inside the value event pair and inside e2e",
                              "cross_shard_dependencies": "followed by the host layer from the batch bytes (no extra exchange): tests/test_gpu_shards.py, tests/test_distributed_cpu.py"}
        if strong:
            line["strong"] = strong
    if world == 1 and args.secondary != "off":
        sec = {}
        sec.update(msm_sweep(torch, lib, ctx, ts, oracle, [16, 18, 20, 22] if args.secondary == "full" else [20], [16, 18, 20] if args.secondary == "full" else [], peak_wide))
        sec["msm_2p20"] = dict(sec["2^20"], note="resident decompressed points (affine Niels, 96 MB), uniform 252-bit scalars; bit-exactness against the oracle: tests/test_gpu_msm.py")
        sec.update(ct_update_metrics(torch, lib, ctx, ts, hbm_peak))
        if args.secondary == "full":
            t_sec = time.time()
            skipped = {}

            def fits(name, est_s):
                """run a configuration only if its estimated cost (16 host cores: the CPU prover mints its batch) still fits the budget"""
                if time.time() - t_bench0 + est_s <= args.time_limit:
                    return True
                skipped[name] = {"skipped": "time budget (--time-limit %g s; about %d s)" % (args.time_limit, est_s), "builder_run": "profiles/r02_bench_1gpu_full.json"}
                return False

            def timed(name, fn):
                t0_ = time.time(); sec[name] = fn(); sec[name]["seconds_incl_minting"] = round(time.time() - t0_, 1)

            # the headline shape's reject paths (all tamper classes)
            if fits("a1k1_rejects", 5):
                timed("a1k1_rejects", lambda: measure_config("a1k1", torch, xhe, verifier, dev, ts, flush, batch.blobs, my_records, [], max(m, 2), 3, 2, tamper_classes(oracle, batch, 77, (2 * args.txs) // 3, a, k), host_threads))

            # config 3, one-sender variant: the reference's own bench shape (benches/tx.rs:153-186), a length-T balance chain
            def run_chain():
                cb = oracle.mint_chain(83, args.txs, 1, threads=ncpu)
                cbl = list(cb.blobs)
                i0 = max(0, min(5000, args.txs - 2))
                swapped = cbl[:i0] + [cbl[i0 + 1], cbl[i0]] + cbl[i0 + 2:]
                bad_sig = cbl[:-1] + [cbl[-1][:-1] + bytes([cbl[-1][-1] ^ 1])]
                return measure_config("chain", torch, xhe, verifier, dev, ts, flush, cbl, cb.ledger().dump(), [], 2, 5, 3,
                                      [("bad_signature_last", bad_sig, (1, args.txs - 1)), ("two_swapped", swapped, (5, -1))], host_threads, canonical_lp_per_tx(1, 1) * args.txs, peak_wide)
            if fits("one_sender_chain_%d" % args.txs, 35):
                timed("one_sender_chain_%d" % args.txs, run_chain)

            # benches/tx.rs:231-233: 16 transactions of 255 transfers each (256-party aggregated range proofs)
            def run_255():
                rng = oracle.Rng(b"bench-255"); led255 = oracle.Ledger(); recs255 = []
                rcv = oracle.Keypair.derive(b"bench255-rcv"); ct = rcv.encrypt(0, rng); led255.set_balance(rcv.pk, oracle.NATIVE, ct); led255.set_nonce(rcv.pk, 0); recs255.append((rcv.pk, oracle.NATIVE, ct))
                kps = [oracle.Keypair.derive(b"bench255-%d" % i) for i in range(16)]
                for kp in kps:
                    ct = kp.encrypt(10**7, rng); led255.set_balance(kp.pk, oracle.NATIVE, ct); led255.set_nonce(kp.pk, 0); recs255.append((kp.pk, oracle.NATIVE, ct))
                res255 = [None] * 16

                def build255(i):
                    res255[i] = oracle.build_tx(kps[i], led255, oracle.Rng(b"bench-255-%d" % i), fee=3, transfers=[(oracle.NATIVE, rcv.pk, 1)] * 255, balances=[(oracle.NATIVE, 10**7)])
                th255 = [threading.Thread(target=build255, args=(i,)) for i in range(16)]
                for t_ in th255:
                    t_.start()
                for t_ in th255:
                    t_.join()
                return measure_config("16x255", torch, xhe, verifier, dev, ts, flush, res255, recs255, [], 256, 5, 3,
                                      [("bad_signature", res255[:9] + [res255[9][:-1] + bytes([res255[9][-1] ^ 1])] + res255[10:], (1, 9))], host_threads, canonical_lp_per_tx(1, 255) * 16, peak_wide)
            if fits("16x255_transfers", 10):
                timed("16x255_transfers", run_255)

            # config 3, other shapes: multi-destination / multi-asset transfers at 10k
            def run_shape(sa, sk_, seed):
                sb = oracle.mint_transfers(seed, args.txs, sa, sk_, threads=ncpu)
                sm_ = 1
                while sm_ < sa + sk_:
                    sm_ *= 2
                rej = tamper_classes(oracle, sb, seed, (2 * args.txs) // 3, sa, sk_)
                return measure_config("a%dk%d" % (sa, sk_), torch, xhe, verifier, dev, ts, flush, sb.blobs, sb.ledger().dump(), [], max(sm_, 2), 5, 3, rej, host_threads, canonical_lp_per_tx(sa, sk_) * args.txs, peak_wide)
            for sa, sk_, seed, est in ((1, 3, 81, 40), (2, 6, 82, 55)):
                if fits("a%dk%d_%d" % (sa, sk_, args.txs), est):
                    timed("a%dk%d_%d" % (sa, sk_, args.txs), lambda: run_shape(sa, sk_, seed))

            # config 5: mixed batch with multisig; one tampered run per class
            def run_mixed(n_mixed):
                mb = oracle.mint_mixed(84, n_mixed, threads=ncpu)
                v = (2 * n_mixed) // 3
                while mb.blobs[v][1] != 0 or mb.blobs[v][3] != 0xFF:
                    v += 1
                kk_ = int.from_bytes(mb.blobs[v][4:8], "little")
                rej = tamper_classes(oracle, mb, 84, v, mb.blobs[v][2], kk_)
                out = measure_config("mixed", torch, xhe, verifier, dev, ts, flush, mb.blobs, mb.ledger().dump(), mb.ledger().dump_multisig(), 8, 3, 2, rej, ncpu)
                out["mix"] = "60 % transfers (k 1..4, a 1..2), 15 % burn, 15 % contract call, 5 % multisig set-up, 5 % transfers from threshold-2 multisig accounts"
                return out
            if fits("mixed_%d" % args.mixed_txs, 300 * args.mixed_txs / 100000.0):
                timed("mixed_%d" % args.mixed_txs, lambda: run_mixed(args.mixed_txs))
            else:      # the same mix at the largest size that still fits the budget (named by its size)
                for n_small in (50000, 20000):
                    if n_small < args.mixed_txs and time.time() - t_bench0 + 300 * n_small / 100000.0 <= args.time_limit:
                        timed("mixed_%d" % n_small, lambda: run_mixed(n_small))
                        break
            sec.update(skipped)
            sec["secondary_seconds"] = round(time.time() - t_sec, 1)
        line["secondary"] = sec
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(batch, ncpu, args.cpu_sample or args.txs)
    emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
