import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import xelis_he_b200 as xhe
ctx = xhe.Ctx(0, party_capacity=0)
lib = ctx.lib
lib.xhe_bench_op.restype = C.c_int32
lib.xhe_bench_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
names = ["fe_mul", "fe_sq", "ge_double", "ge_madd", "ge_add", "fe_add", "quad_double", "quad_add", "quad_madd"]
res = {}
GHZ = 1.92
for op, name in enumerate(names):
    row = {}
    # (threads/block, blocks/SM): resident warps per SMSP = tpb/32*bps/4 if registers allow
    for tpb, bps in ((32, 1), (128, 1), (128, 2), (128, 4)):
        c = (C.c_double * 2)()
        rc = lib.xhe_bench_op(ctx.p, op, tpb, 148 * bps, 1000, c)
        row[f"{tpb}x{bps}"] = (round(c[0]), round(c[1] * GHZ)) if rc == 0 else rc   # (warp-0 latency cycles, cycles per warp-op per SMSP from wall time)
    res[name] = row
    print(name, row, flush=True)
lib.xhe_bench_oct.restype = C.c_int32
lib.xhe_bench_oct.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double)]
for op, name in enumerate(["oct_mul", "oct_double", "oct_add_pt", "oct_add"]):      # warp-cooperative forms (csrc/oct.cuh): one warp, dependent chain
    c = C.c_double()
    rc = lib.xhe_bench_oct(ctx.p, op, 2000, C.byref(c))
    res[name] = {"32x1": round(c.value)} if rc == 0 else rc
    print(name, res[name], flush=True)
os.makedirs("gpurun_out", exist_ok=True); json.dump(res, open("gpurun_out/op_bench.json", "w"), indent=1)
