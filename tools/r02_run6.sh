#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/msm_bench.py 16 18 20 22 2>&1 | tail -1
XHE_MSM_SEG_LOG=4 timeout 300 python tools/msm_bench.py 16 18 20 22 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --secondary off --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02l_bench.json').read().strip().splitlines()[-1])
print("value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_call"]["ms_per_step"],2), "inflight", round(d["value_batches_in_flight"]["value_this_rank"]), "launches", d["gpu_launches"])
print({k:v for k,v in d["kernels_ms_per_step_isolated"].items()})
print(d["timeline_ms_one_step"])
print(d["roofline"]["msm_frac_in_batch"], d["roofline"]["step_frac"])
PY
