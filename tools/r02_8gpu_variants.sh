#!/bin/bash
# 8-GPU variants of the end-to-end pipeline (development: pre-minted batch)
mkdir -p gpurun_out
run() { tag=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 18 --warmup 3 --mint-cache _cache/emul_80k.pkl --no-strong "$@" > gpurun_out/r02i_8gpu_$tag.json 2> gpurun_out/r02i_8gpu_$tag.err
  python - $tag <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r02i_8gpu_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"e2e ms",round(d["e2e"]["ms_per_step"],3),"single",round(d["e2e"]["single_call"]["ms_per_step"],2),"threads",d["e2e"]["host_threads"],"nfl",d["e2e"]["batches_in_flight"])
    print("   decider",d["e2e"]["decision_thread_ms_per_batch"])
    for r in d["e2e"]["phases_ms_by_rank"][::7]: print("   rank",r["rank"],"single",r["single_call"],"\n      pipe",r["pipelined"])
except Exception as e: print(sys.argv[1],"failed",e)
PY
}
lscpu | grep -i "model name\|^CPU(s)\|Thread\|Socket\|NUMA" | head -8
nvidia-smi topo -m 2>/dev/null | head -12
run bind6
run bind3 --inflight 3
run nobind3 --no-bind --inflight 3
grep "cores" gpurun_out/r02i_8gpu_bind6.err | head -8
