#!/bin/bash
mkdir -p gpurun_out
for SLOTS in 0 1 2 3; do
  XHE_DEVICE_SLOTS=$SLOTS timeout 600 python bench.py --steps 20 --warmup 3 --secondary off --no-cpu-baseline > gpurun_out/r02f_bench_s$SLOTS.json 2> gpurun_out/r02f_bench_s$SLOTS.err; echo "bench slots=$SLOTS rc=$?"
  python - $SLOTS <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02f_bench_s%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print("slots",sys.argv[1],"value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"],3), "single", round(d["e2e"]["single_call"]["ms_per_step"],2), "inflight", round(d["value_batches_in_flight"]["value_this_rank"]))
print("  pipe", d["e2e"]["phases_ms_pipelined_mean"])
PY
done
timeout 300 python tools/msm_bench.py 16 18 20 22 2>&1 | tail -1
