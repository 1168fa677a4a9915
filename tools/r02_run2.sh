#!/bin/bash
# GPU call: parity tests with the joint MSM, quick bench joint vs split, rank emulator
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/r02c_gputests.log
for MODE in joint split; do
  if [ $MODE = split ]; then export XHE_SPLIT_MSM=1; else unset XHE_SPLIT_MSM; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --secondary off > gpurun_out/r02c_bench_$MODE.json 2> gpurun_out/r02c_bench_$MODE.err; echo "bench $MODE rc=$?"
  python - $MODE <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02c_bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1],"value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_call"]["ms_per_step"],2), "inflight", round(d["value_batches_in_flight"]["value_this_rank"]), "launches", d["gpu_launches"])
print({k:v for k,v in d["kernels_ms_per_step_isolated"].items()})
print(d["timeline_ms_one_step"])
print(d["roofline"]["msm_frac_in_batch"], d["roofline"]["step_frac"])
PY
done
unset XHE_SPLIT_MSM
STEPS=24 timeout 600 python tools/r02_rank_emul.py 7 2>&1 | tail -2
