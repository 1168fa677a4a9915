#!/bin/bash
# final multi-GPU evidence (N GPUs visible): bench.py at N with the pre-minted batch (timed regions identical to the driver's run)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29550 bench.py --gpus $N --steps 20 --warmup 3 --mint-cache _cache/emul_80k.pkl > gpurun_out/r02q_bench_${N}gpu.json 2> gpurun_out/r02q_bench_${N}gpu.err; echo "bench N=$N rc=$?"
python - $N <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02q_bench_%sgpu.json'%sys.argv[1]).read().strip().splitlines()[-1])
print("N",sys.argv[1],"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"]),"e2e ms",round(d["e2e"]["ms_per_step"],3),"single",round(d["e2e"]["single_call"]["ms_per_step"],2),"threads",d["e2e"]["host_threads"],"nfl",d["e2e"]["batches_in_flight"])
print("  h2d",d["e2e"]["h2d_all_ranks_at_once"]); print("  strong",d.get("strong")); print("  decider",d["e2e"]["decision_thread_ms_per_batch"])
PY
