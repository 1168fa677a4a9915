#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_basic.py tests/test_gpu_msm.py -m gpu -x -q > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02e_tests.log
timeout 300 python tools/msm_bench.py 16 18 20 22 2>&1 | tail -2
XHE_MSM_QUAD_HORNER=1 timeout 300 python tools/msm_bench.py 16 18 20 22 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --secondary off > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench.json').read().strip().splitlines()[-1])
print("value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_call"]["ms_per_step"],2), "inflight", round(d["value_batches_in_flight"]["value_this_rank"]), "launches", d["gpu_launches"])
print({k:v for k,v in d["kernels_ms_per_step_isolated"].items()})
print(d["timeline_ms_one_step"])
print(d["roofline"]["msm_frac_in_batch"], d["roofline"]["step_frac"])
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
