#!/bin/bash
# GPU call: parity tests, a quick bench, the rank emulator (rank 7 of 8) with and without the key-digest index
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r02b_gputests.log
timeout 600 python bench.py --steps 10 --warmup 3 --secondary off > gpurun_out/r02b_bench_quick.json 2> gpurun_out/r02b_bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02b_bench_quick.json').read().strip().splitlines()[-1])
print("value",round(d["value"]), "ms",round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "single", d["e2e"]["single_call"])
print({k:v for k,v in d["kernels_ms_per_step_isolated"].items()})
print(d["timeline_ms_one_step"])
PY
STEPS=24 timeout 600 python tools/r02_rank_emul.py 7 0 2>&1 | tail -4
NOINDEX=1 STEPS=24 timeout 600 python tools/r02_rank_emul.py 7 2>&1 | tail -2
