#!/bin/bash
# round-2 GPU check: full parity suite, then the MSM sweep and a short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/gputests.log
tail -25 gpurun_out/gputests.log
for G in 1 4; do for CH in 0 1; do
  XHE_MSM_CHAIN=$CH XHE_MSM_GROUPS=$G timeout 300 python tools/msm_bench.py 16 17.2 17.6 18 20 22 2>&1 | tail -1 | sed "s/^/CHAIN=$CH G=$G /" | tee -a gpurun_out/msm_sweep2.log
done; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
XHE_MSM_CHAIN=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_quick_nochain.json 2> gpurun_out/bench_quick_nochain.err; echo "bench nochain rc=$?"
python - <<'PY'
import json
for f in ("bench_quick", "bench_quick_nochain"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "single", round(d["e2e"]["single_call"]["ms_per_step"], 2), "inflight", round(d["value_batches_in_flight"]["value_this_rank"] or 0))
        print(" isolated", d["kernels_ms_per_step_isolated"])
        print(" timeline", d["timeline_ms_one_step"])
    except Exception as e:
        print(f, "failed", e)
PY
tail -5 gpurun_out/bench_quick.err
