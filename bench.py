#!/usr/bin/env python3
"""bench.py -- verified TX/s on a 10k-transfer batch (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--shape a1k1] [--txs 10000]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the CPU path (oracle port of the reference; the Rust crate cannot run here)

A step = one verification of one batch of T synthetic transfer transactions per GPU (weak scaling: every rank verifies a
T-transaction shard, the per-rank partial MSM points are all-gathered over NCCL and summed before the identity check).
`value` times the device kernels with the batch resident in HBM (xhe_batch_run); `e2e` times the reference-facing call
(host parsing + state resolution + Merlin transcripts + H2D + kernels + D2H + signature hashes) from host buffers.
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the MSM points/s and ciphertext-update (HBM) side measurements")
    ap.add_argument("--inflight", type=int, default=6, help="batches in flight per GPU in the end-to-end measurement (one context + host thread each)")
    ap.add_argument("--fiat-shamir", default="fast", choices=["fast", "device", "host"], help="where the Merlin transcripts run (host = north_star split; device = SURVEY 8 f.1)")
    return ap.parse_args()


def shape_ak(shape):
    a, k = shape[1:].split("k")
    return int(a), int(k)


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



def secondary_metrics(lib, ctx, stream, hbm_peak_gbs):
    """BASELINE.json's other single-GPU figures, measured live on the device (CUDA events on the ctx stream, L2-sized
    inputs): MSM points/s on 2^20 resident points (config 2) and the resident ciphertext update (config 4, HBM-bound)."""
    import ctypes as C
    import torch
    out = {}
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    n = 1 << 20

    def timed(fn, iters=5, warm=3):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                assert fn() == 0
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                assert fn() == 0
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    with torch.cuda.stream(stream):
        uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
        enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda"); aff = torch.empty((n, 16), dtype=torch.int32, device="cuda")
        niels = torch.empty((n, 24), dtype=torch.int32, device="cuda"); ok = torch.empty((n,), dtype=torch.uint8, device="cuda")
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g); sc[:, 31] &= 0x0F      # < 2^252 < l
        assert lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr()) == 0
        assert lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, aff.data_ptr(), niels.data_ptr(), ok.data_ptr()) == 0
    lib.xhe_msm_workspace_bytes.restype = C.c_size_t; lib.xhe_msm_workspace_bytes.argtypes = [C.c_void_p, C.c_size_t]
    wsb = lib.xhe_msm_workspace_bytes(ctx.p, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); res = torch.zeros(64, dtype=torch.uint8, device="cuda")
    lib.xhe_msm_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    ms = timed(lambda: lib.xhe_msm_dev(ctx.p, sc.data_ptr(), niels.data_ptr(), n, ws.data_ptr(), wsb, res.data_ptr(), res.data_ptr() + 32))
    out["msm_2p20"] = {"points": n, "ms": ms, "points_per_s": n / ms * 1e3, "alg_TLP_s": (8064.0 * n + 6.04e8) / ms / 1e9,
                       "note": "resident decompressed points (affine Niels, 96 MB), uniform 252-bit scalars; bit-exactness against the oracle: tests/test_gpu_msm.py"}
    ms_dec = timed(lambda: lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, aff.data_ptr(), niels.data_ptr(), ok.data_ptr()))
    out["decompress_2p20"] = {"ms": ms_dec, "points_per_s": n / ms_dec * 1e3, "alg_TLP_s": n * 12632.0 / ms_dec / 1e9}
    ms_both = timed(lambda: lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, aff.data_ptr(), niels.data_ptr(), ok.data_ptr())
                    or lib.xhe_msm_dev(ctx.p, sc.data_ptr(), niels.data_ptr(), n, ws.data_ptr(), wsb, res.data_ptr(), res.data_ptr() + 32))
    out["msm_2p20"]["incl_decompression"] = {"ms": ms_both, "points_per_s": n / ms_both * 1e3}
    # config 4, compressed I/O: 64-byte ciphertexts in and out (4 decodes + 2 encodes per account: integer-bound)
    na_c = 1 << 19
    subc = torch.randint(0, 2, (na_c,), dtype=torch.uint8, device="cuda", generator=g)
    outb = torch.empty((na_c, 64), dtype=torch.uint8, device="cuda"); okb = torch.empty((na_c,), dtype=torch.uint8, device="cuda")
    bal = enc[: 2 * na_c].reshape(na_c, 64); delta = enc.flip(0)[: 2 * na_c].reshape(na_c, 64).contiguous()
    lib.xhe_ct_update_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    ms_c = timed(lambda: lib.xhe_ct_update_dev(ctx.p, bal.data_ptr(), delta.data_ptr(), subc.data_ptr(), na_c, outb.data_ptr(), okb.data_ptr()))
    out["ct_update_compressed_512k"] = {"accounts": na_c, "ms": ms_c, "accounts_per_s": na_c / ms_c * 1e3, "alg_TLP_s": na_c * 76000.0 / ms_c / 1e9}
    del uni, aff, ws
    na = 1 << 20
    balr = torch.randint(0, 2**31 - 1, (4, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
    dn = torch.randint(0, 2**31 - 1, (3, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
    sub = torch.randint(0, 2, (na,), dtype=torch.uint8, device="cuda", generator=g)
    lib.xhe_ct_update_resident_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    ms = timed(lambda: lib.xhe_ct_update_resident_dev(ctx.p, balr.data_ptr(), dn.data_ptr(), sub.data_ptr(), na), iters=10)
    bytes_per_account = 2 * (128 + 128 + 96)       # two points per account: extended balance read + write, affine-Niels delta read
    gbs = na * bytes_per_account / ms / 1e6
    out["ct_update_resident_1M"] = {"accounts": na, "ms": ms, "accounts_per_s": na / ms * 1e3,
                                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak_gbs, "unit": "GB/s", "frac": gbs / hbm_peak_gbs, "traffic": 688.9e6, "traffic_note": "dram__bytes_read.sum (470.8 MB = 2 M points x 224 B, exactly the operands) + dram__bytes_write.sum (218.1 MB; the rest of the 268 MB written is still in L2 when the kernel ends) of one launch, ncu, profiles/r01_ct_resident_traffic.csv: no re-reads",
                                                 "algorithmic_bytes_per_account": bytes_per_account, "working_set_mb": na * bytes_per_account / 1e6}}
    return out

def main():
    # exactly one JSON line may reach stdout: libraries (NCCL's version banner, torchrun) write there too, so everything else
    # is routed to stderr and the line is written to the saved descriptor at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    args = parse_args()
    a, k = shape_ak(args.shape)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ncpu = len(os.sched_getaffinity(0))
    workload = f"batch verify {args.txs} transfer TXs per GPU, shape {args.shape} (assets a, transfers k), 64-bit aggregated range proofs + sigma proofs + signatures"
    config = {"workload": workload, "txs_per_gpu": args.txs, "shape": args.shape, "parallelism": f"tx-shard x{world}" if world > 1 else "single",
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- synthetic workload: rank 0 mints with all host threads, the other ranks receive the same bytes
    t_mint = time.time()
    if rank == 0:
        batch = oracle.mint_transfers(77, args.txs, a, k, threads=ncpu)
        blobs = batch.blobs; records = batch.ledger().dump()
    else:
        batch, blobs, records = None, None, None
    if world > 1:
        obj = [blobs, records]
        dist.broadcast_object_list(obj, src=0)
        blobs, records = obj
    t_mint = time.time() - t_mint
    m = 1
    while m < a + k:
        m *= 2
    ctx = xhe.Ctx(local, party_capacity=max(m, 2))
    # every context works on its own non-blocking stream, so batches in flight (and the NCCL exchange) overlap on the device
    streams = [torch.cuda.Stream()]
    ctx.set_stream(streams[0].cuda_stream)
    lib = ctx.lib
    ledger0 = verifier.Ledger(); ledger0.import_records(records)
    prepared = verifier.prepare_blobs(blobs)
    host_threads = max(1, ncpu // max(1, min(world, 8)))

    from xelis_he_b200 import distributed as xd

    fs_mode = [args.fiat_shamir]

    def e2e_step(seed):
        led = ledger0.clone()
        t0 = time.perf_counter()
        if dist:     # sharded batch: local partial verification + 80-byte all-gather over NCCL + joint decision on every rank
            code, idx, tm = xd.verify_batch_distributed(ctx, None, led, rank * args.txs, seed=seed + b"r%d" % rank, threads=host_threads, prepared=prepared, commit=False, fiat_shamir=fs_mode[0])
        else:
            code, idx, tm = verifier.verify_batch(ctx, None, led, seed=seed, threads=host_threads, prepared=prepared, fiat_shamir=fs_mode[0])
        return time.perf_counter() - t0, code, idx, tm

    # ---- correctness gate + warm-up (also leaves the batch resident in HBM for the device-only timing)
    for w in range(max(args.warmup, 3)):
        dt, code, idx, tm = e2e_step(b"warm%d" % w)
        assert (code, idx) == (0, -1), (code, idx)
    if rank == 0 and args.txs <= 20000:
        sl = min(64, args.txs)
        assert oracle.verify_batch(blobs[:sl], batch.slice(sl).ledger()) == (0, -1)      # the oracle agrees on a prefix of the workload

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib.xhe_batch_run.argtypes = [C.c_void_p]; lib.xhe_batch_run.restype = C.c_int32
    lib.xhe_batch_h2d_bytes.restype = C.c_size_t; lib.xhe_batch_h2d_bytes.argtypes = [C.c_void_p]
    lib.xhe_batch_d2h_bytes.restype = C.c_size_t; lib.xhe_batch_d2h_bytes.argtypes = [C.c_void_p]
    lib.xhe_ctx_timing.argtypes = [C.c_void_p, C.c_int]
    lib.xhe_ctx_timing_read.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.c_int]
    for _ in range(max(args.warmup, 3)):
        assert lib.xhe_batch_run(ctx.p) == 0
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    # ---- timed region 1: device kernels on the resident batch (value)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    launches0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    rec_local = torch.zeros(80, dtype=torch.uint8, device="cuda"); rec_all = torch.zeros(80 * world, dtype=torch.uint8, device="cuda")
    ts = streams[0]
    for s in range(args.steps):
        with torch.cuda.stream(ts):
            flush.fill_(s & 0xFF)              # evict the previous step's working set from L2 (outside the event pair)
            ev[s][0].record(ts)
            assert lib.xhe_batch_run(ctx.p) == 0
            if dist:
                dist.all_gather_into_tensor(rec_all, rec_local)     # the per-batch exchange of (verdict, partial encodings)
            ev[s][1].record(ts)
    torch.cuda.synchronize()
    dev_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    launches = ctx.launches - launches0
    # isolated per-kernel durations for the roofline: the same steps with the pipelines serialised on one stream and
    # CUDA events around the main kernels (timers under stream concurrency would include the overlapped neighbours)
    lib.xhe_ctx_set_serial.argtypes = [C.c_void_p, C.c_int]
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
    # where each kernel sits inside one concurrent step (four stream pipelines): start/end in ms from the step start
    lib.xhe_ctx_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
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
    # (a) one call at a time (latency); (b) two batches in flight on two contexts of the same GPU, so the host phase of
    # one batch overlaps the device phase of the other -- what a node verifying a stream of batches does.
    single_s = 0.0; phases = {}
    for s in range(args.steps):
        flush.fill_(s & 0xFF); torch.cuda.synchronize()
        dt, code, idx, tm = e2e_step(b"step%d" % s)
        assert (code, idx) == (0, -1)
        single_s += dt
        for kk, vv in tm.items():
            phases[kk] = phases.get(kk, 0.0) + vv / args.steps
    barrier()
    pipelined = args.fiat_shamir != "host" and args.inflight > 1
    nfl = args.inflight if pipelined else 1
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

    def worker(widx, nsteps, out, ledgers=None):
        c = workers[widx]
        torch.cuda.set_device(local)
        for s in range(nsteps):
            led = ledgers[s] if ledgers else ledger0.clone()      # fresh state per step (cloned before the clock starts)
            if dist:
                seq = seq_base[0] + s * nfl + widx
                code, idx, s_enc, r_enc, tmw = verifier.verify_batch_partial(c, None, led, seed=b"p%d-%d-%d" % (widx, s, rank), threads=wthreads, prepared=prepared, fiat_shamir=args.fiat_shamir)
                decider.submit(seq, xd.pack_local(code, idx, rank * args.txs, s_enc, r_enc), verifier.take_pending(c), led)
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
    clocks = sampler.summary()
    h2d, d2h = lib.xhe_batch_h2d_bytes(ctx.p), lib.xhe_batch_d2h_bytes(ctx.p)

    if dist:
        t = torch.tensor([dev_ms, e2e_s * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms_max, e2e_ms_max = t.tolist()
    else:
        dev_ms_max, e2e_ms_max = dev_ms, e2e_s * 1e3

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    total_tx = args.txs * world * args.steps
    value = total_tx / (dev_ms_max * 1e-3)
    e2e = total_tx / (e2e_ms_max * 1e-3)
    # ---- roofline of the dominant kernel (integer-multiply pipe; tensor cores unused by design)
    peak_wide = ctx.int_peak(2); peak_chain = ctx.int_peak(3)
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6553.9          # the figure MEASURED_PEAKS.json held when this was written
    leaf = {n: v for n, v in kernels.items() if not n.startswith("msm_")}      # msm_* timers wrap several kernels
    # dominant kernel = the one carrying the largest share of the step's algorithmic work (limb products); the per-kernel
    # table below lists every timed kernel, including the latency-bound one-thread-per-item kernels
    dom = max(leaf, key=lambda n: leaf[n]["alg_lp_per_step"])
    ach = leaf[dom]["alg_lp_per_step"] / (leaf[dom]["ms_per_step"] * 1e-3)
    work = {n: {"ms": round(v["ms_per_step"], 4), "TLP_s": round(v["alg_lp_per_step"] / (v["ms_per_step"] * 1e-3) / 1e12, 3), "frac": round(v["alg_lp_per_step"] / (v["ms_per_step"] * 1e-3) / peak_wide, 4)}
            for n, v in leaf.items() if v["alg_lp_per_step"] > 0 and v["ms_per_step"] > 0}
    roofline = {"bound": "int-mul", "kernel": dom, "achieved": ach / 1e12, "peak": peak_wide / 1e12, "unit": "TLP/s (32x32->64 limb products)", "frac": ach / peak_wide, "share_of_step_work": leaf[dom]["alg_lp_per_step"] / max(1.0, sum(v["alg_lp_per_step"] for n, v in leaf.items())),
                "peak_source": "measured live: IMAD.WIDE.U32 Rd64, Ra, Rb, RZ microkernel (two vector multiplicands, both result words live; SASS checked: IMAD.WIDE only)", "peak_carry_chain": peak_chain / 1e12, "frac_of_carry_chain_peak": ach / peak_chain,
                "traffic": 13.39e6, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_decompress launch of this workload (ncu --set full, profiles/r01_ncu_full_v3.md); algorithmic bytes are 65.6 MB (32 B in, 160 B out per point): the outputs stay in the 126 MB L2",
                "per_kernel_isolated": work, "note": "every 32x32->64 form (IMAD.WIDE, IMAD.WIDE.X carry chains, IMAD.HI) issues at 4 cycles per warp instruction per SM sub-partition on sm_100a, half the rate of the 32-bit IMAD; bench lines before r01 v5 divided by an 18.4 T/s figure that turned out to measure IADD3 pairs (DESIGN.md 4.1)"}
    line = {"metric": "verified TX/s (10k-transfer batch)", "value": value, "unit": "TX/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "vs_baseline_note": "reference README: ~0.40 ms/TX on one CPU thread (hardware unstated)",
            "dtype": "u32 limbs (GF(2^255-19), mod l)", "data": "synthetic (valid TXs minted by the oracle prover; ranks share one minted batch)", "config": config,
            "e2e": {"value": e2e, "unit": "TX/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms_max / args.steps, "host_threads": host_threads, "batches_in_flight": nfl,
                    "single_call": {"value": args.txs * args.steps / single_s, "ms_per_step": 1e3 * single_s / args.steps},
                    "phases_ms": {kk: round(vv, 3) for kk, vv in phases.items() if kk != "keccak_f"}, "host_keccak_f_per_tx": phases.get("keccak_f", 0) / args.txs,
                    "phases_ms_pipelined_mean": {kk: round(sum(t_[kk] for t_ in pipe_phases) / len(pipe_phases), 3) for kk in pipe_phases[0] if kk.endswith("_ms")} if pipe_phases else None,
                    "decision_thread_ms_per_batch": decider_stats, "fiat_shamir": args.fiat_shamir, "other_mode": {"fiat_shamir": other, "value_this_rank": args.txs * 3 / t_other}},
            "value_batches_in_flight": {"value_this_rank": concurrent_value, "unit": "TX/s", "contexts": nfl, "note": "device-resident batches of all contexts in flight at once (no L2 flush); the GPU-side ceiling of the pipelined e2e"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels_ms_per_step_isolated": {n: round(v["ms_per_step"], 4) for n, v in kernels.items()},
            "timeline_ms_one_step": timeline, "mint_seconds": round(t_mint, 1), "host_cores": ncpu}
    if world > 1:
        line["collective"] = {"what": "one all_gather of 80 B per rank per batch (verdict + partial sigma / range MSM encodings) over NCCL, inside both timed regions"}
    if world == 1 and not args.no_secondary:
        line["secondary"] = secondary_metrics(lib, ctx, ts, hbm_peak)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(batch, ncpu, args.cpu_sample or args.txs)
    emit(line)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
