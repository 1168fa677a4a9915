"""MSM microbench (config 2) on resident points: points/s and fraction of the measured IMAD.WIDE peak."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import xelis_he_b200 as xhe

ctx = xhe.Ctx(0, party_capacity=2)
lib = ctx.lib
peak = ctx.int_peak(2)
variant = int(os.environ.get('XHE_ACC_VARIANT', '4'))
lib.xhe_msm_set_variant(variant)
res = {"imad_wide_peak": peak, "variant": variant}
g = torch.Generator(device="cuda"); g.manual_seed(1)
logs = [float(a) for a in sys.argv[1:]] or [16, 18, 20, 22]
for logn in logs:
    n = int(round(2 ** logn)); logn = int(logn) if float(logn).is_integer() else logn
    uni = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda", generator=g)
    enc = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
    niels = torch.empty((n, 24), dtype=torch.int32, device="cuda")
    ok = torch.empty((n,), dtype=torch.uint8, device="cuda")
    lib.xhe_from_uniform_dev(ctx.p, uni.data_ptr(), n, enc.data_ptr())
    lib.xhe_decompress_dev(ctx.p, enc.data_ptr(), n, None, niels.data_ptr(), ok.data_ptr())
    sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
    sc[:, 31] &= 0x0F
    wsb = lib.xhe_msm_workspace_bytes(ctx.p, n)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")
    out = torch.zeros((64,), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def run():
        rc = lib.xhe_msm_dev(ctx.p, sc.data_ptr(), niels.data_ptr(), n, ws.data_ptr(), wsb, out.data_ptr(), out.data_ptr() + 32)
        assert rc == 0, rc
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(5):
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    ms = min(times)
    c, W = ctx.msm_plan(n)
    lp = 8064.0 * n + 6.04e8
    res[f"2^{logn}"] = {"ms": ms, "Mpts_s": n / ms / 1e3, "c": c, "W": W, "alg_TLP_s": lp / ms / 1e9, "frac_of_imad_peak": lp / (ms * 1e-3) / peak, "ws_MB": wsb / 1e6,
                        "enc": bytes(out[:32].cpu().numpy()).hex()}
    del uni, enc, niels, ok, sc, ws
    torch.cuda.empty_cache()
print(variant, {k: (round(v['ms'],3), round(v['Mpts_s'],1), round(v['frac_of_imad_peak'],3)) for k, v in res.items() if isinstance(v, dict)})
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/msm_bench_v%d.json" % variant, "w"), indent=1)
