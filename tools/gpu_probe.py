"""First-contact probe on the B200 box: host shape, integer-pipe peaks, per-point kernel timings (CUDA events)."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xelis_he_b200 as xhe

out = {"nproc": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
try:
    out["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass
out["gpu"] = torch.cuda.get_device_name(0)
ctx = xhe.Ctx(0, party_capacity=8)
lib = ctx.lib
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
out["clocks_idle"] = clk
for rep in range(2):
    out[f"int_peak_{rep}"] = {"imad_lo": ctx.int_peak(0), "imad_hi": ctx.int_peak(1), "imad_wide": ctx.int_peak(2)}


def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(iters):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(1)
uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
aff = torch.empty((n, 16), dtype=torch.int32, device="cuda")
niels = torch.empty((n, 24), dtype=torch.int32, device="cuda")
ok = torch.empty((n,), dtype=torch.uint8, device="cuda")
ms = timed(lambda: lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr()))
out["from_uniform_1M_ms"] = ms
ms = timed(lambda: lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, aff.data_ptr(), niels.data_ptr(), ok.data_ptr()))
out["decompress_1M_ms"] = ms; out["decompress_Mpts_s"] = n / ms / 1e3; out["decompress_TLP_s"] = n * 12632 / ms / 1e9
assert int(ok.sum()) == n
# compressed-I/O ciphertext update on 512k accounts
na = 1 << 19
sub = torch.randint(0, 2, (na,), dtype=torch.uint8, device="cuda", generator=g)
outb = torch.empty((na, 64), dtype=torch.uint8, device="cuda")
okb = torch.empty((na,), dtype=torch.uint8, device="cuda")
bal = enc[: 2 * na].reshape(na, 64); delta = enc.flip(0)[: 2 * na].reshape(na, 64).contiguous()
ms = timed(lambda: lib.xhe_ct_update_dev(ctx.p, bal.data_ptr(), delta.data_ptr(), sub.data_ptr(), na, outb.data_ptr(), okb.data_ptr()))
out["ct_update_compressed_512k_ms"] = ms; out["ct_update_compressed_Macc_s"] = na / ms / 1e3
# resident update, 1M accounts: planar extended balances (4 x 2n x 8 words) and planar niels deltas
na = 1 << 20
balr = torch.randint(0, 2**31 - 1, (4, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
dn = torch.randint(0, 2**31 - 1, (3, 2 * na, 8), dtype=torch.int32, device="cuda", generator=g)
sub = torch.randint(0, 2, (na,), dtype=torch.uint8, device="cuda", generator=g)
ms = timed(lambda: lib.xhe_ct_update_resident_dev(ctx.p, balr.data_ptr(), dn.data_ptr(), sub.data_ptr(), na), iters=10)
out["ct_update_resident_1M_ms"] = ms; out["ct_update_resident_GBs"] = na * 704 / ms / 1e6
out["launches"] = ctx.launches
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
