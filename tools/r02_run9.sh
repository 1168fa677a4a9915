#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02o_bench_default.json 2> gpurun_out/r02o_bench_default.err; echo "default bench rc=$? in $(( $(date +%s) - S )) s"; tail -2 gpurun_out/r02o_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02o_bench_default.json').read().strip().splitlines()[-1])
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"single",d["e2e"]["single_call"]["ms_per_step"])
for k,v in d["secondary"].items():
    if isinstance(v,dict): print(k, v.get("skipped") or v.get("seconds_incl_minting") or "", (v.get("value") or {}) if isinstance(v.get("value"),dict) else "")
PY
