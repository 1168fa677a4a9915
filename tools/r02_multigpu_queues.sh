#!/bin/bash
# A/B of the number of hardware queues (CUDA_DEVICE_MAX_CONNECTIONS) for a sharded batch with several batches in flight
N=$1
for V in conn32 conn8 conn32b conn8b; do
  case $V in conn32*) export CUDA_DEVICE_MAX_CONNECTIONS=32;; conn8*) export CUDA_DEVICE_MAX_CONNECTIONS=8;; esac
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus $N --steps 20 --warmup 3 --mint-cache _cache/emul_80k.pkl --no-strong > gpurun_out/r02r_${N}gpu_$V.json 2> gpurun_out/r02r_${N}gpu_$V.err; echo "$V rc=$?"
  python - $N $V <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02r_%sgpu_%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
print(" ",sys.argv[2],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"e2e ms",round(d["e2e"]["ms_per_step"],3),"single",round(d["e2e"]["single_call"]["ms_per_step"],2),"nfl",d["e2e"]["batches_in_flight"])
print("    pipe rank0",d["e2e"]["phases_ms_by_rank"][0]["pipelined"])
PY
done
