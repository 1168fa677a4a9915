"""Host phases of the fast path without a device (diagnostics: flag bit 4 of xheh_verify_batch_ex): parse / resolve / staging times
of a 10k batch against a 160k-account ledger, single-threaded and with several walks running at once (as batches in flight do).
  python tools/host_profile.py [workers]"""
import ctypes as C
import os
import pickle
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xelis_he_b200 as xhe
from xelis_he_b200 import verifier

lib = xhe.load_library()
if os.environ.get("MALLOPT"):
    libc = C.CDLL("libc.so.6"); libc.mallopt(-3, 1 << 30); libc.mallopt(-1, 1 << 30)      # M_MMAP_THRESHOLD, M_TRIM_THRESHOLD: keep freed memory in the heap
cache = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_cache", "emul_80k.pkl")
blobs, records = pickle.load(open(cache, "rb"))
T = 10000
ledger0 = verifier.Ledger(); ledger0.import_records(records)
bl = verifier.prepare_blobs(blobs[:T])
workers = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def run(led, reps, out, fake_ctx):
    fi = C.c_long(-1); tm = (C.c_double * 7)(); acc = [1e9] * 6
    for r in range(reps):
        rc = lib.xheh_verify_batch_ex(C.c_void_p(fake_ctx), led.ptr, bl.ptrs, bl.lens, bl.n, b"x", 1, 1, 4 | 16, C.byref(fi), tm, None)
        assert rc == -1
        for i in range(6):
            acc[i] = min(acc[i], tm[i])
    out.append(acc)


outs = []
th = [threading.Thread(target=run, args=(ledger0.clone(), 12, outs, 4096 * (w + 1))) for w in range(workers)]      # (the context is never dereferenced in a dry run: it only keys the per-context staging cache)
t0 = time.perf_counter()
[t.start() for t in th]; [t.join() for t in th]
dt = time.perf_counter() - t0
mean = [sum(o[i] for o in outs) / len(outs) for i in range(6)]      # (per walk: the fastest of its repetitions)
print(f"{workers} walks at once: parse {mean[0]:.2f} resolve {mean[1]:.2f} stage {mean[2]:.2f} ms per batch; {1e3 * dt / (12 * workers):.2f} ms wall per batch")
