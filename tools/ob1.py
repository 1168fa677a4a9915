import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
import xelis_he_b200 as xhe
ctx = xhe.Ctx(0, party_capacity=0); lib = ctx.lib
lib.xhe_bench_op.restype = C.c_int32
lib.xhe_bench_op.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
c = (C.c_double * 2)()
for op in (0, 1, 3):
    lib.xhe_bench_op(ctx.p, op, 128, 148 * 8, 300, c); print(op, c[0], c[1])
