# scheduling experiment (design evidence): cap how many blocks of one kernel an SM holds with a dummy dynamic shared-memory
# request and watch the step time.  usage: bash tools/exp_residency.sh "NAME:ENV=VAL,ENV=VAL" ...
for cfg in "$@"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $(echo "$envs" | tr ',' ' ') timeout 170 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/exp_$name.json"))
    print("$name", "value", round(d["value"]/1e6,3), "ms", round(d["ms_per_step"],3), "inflight", round(d["value_batches_in_flight"]["value_this_rank"]/1e6,3), "e2e", round(d["e2e"]["value"]/1e6,3))
except Exception as e:
    print("$name failed", e)
PY
done
