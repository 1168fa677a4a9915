#!/bin/bash
# per-launch durations of the MSM microbench under ncu (serialised, cold-cache: compare shares)
mkdir -p gpurun_out
for G in 1 4; do
XHE_MSM_GROUPS=$G timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/ncu_msm_G$G.csv python tools/msm_bench.py 17.2 20 > gpurun_out/ncu_msm_G$G.log 2>&1
echo "G=$G rc=$?"
done
