import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xelis_he_b200 as xhe
ctx = xhe.Ctx(0, party_capacity=0)
lib = ctx.lib
lib.xhe_bench_op.restype = C.c_int32
lib.xhe_bench_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
names = ["fe_mul", "fe_sq", "ge_double", "ge_madd", "ge_add", "fe_add"]
res = {}
for op, name in enumerate(names):
    row = {}
    for tpb, blocks, label, wps in ((32, 1, "1warp", 1), (128, 148, "1w/smsp", 1), (256, 148, "2w/smsp", 2), (512, 148, "4w/smsp", 4), (512, 296, "8w/smsp", 8), (768, 296, "12w/smsp", 12)):
        c = C.c_double()
        rc = lib.xhe_bench_op(ctx.p, op, tpb, blocks, 2000, C.byref(c))
        row[label] = (round(c.value, 1), round(c.value / wps, 1)) if rc == 0 else rc
    res[name] = row
    print(name, row, flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/op_bench.json", "w"), indent=1)
