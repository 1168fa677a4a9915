#!/bin/bash
# round-2 MSM tail experiments: parity first, then the size sweep under different window-group counts / bucket-cost constants
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msm.py -x -q > gpurun_out/msm_tests.log 2>&1; echo "msm tests rc=$?" | tee -a gpurun_out/msm_tests.log
tail -5 gpurun_out/msm_tests.log
for G in 1 2 4 8; do
  XHE_MSM_GROUPS=$G timeout 300 python tools/msm_bench.py 14 16 17.2 17.6 18 20 22 2>&1 | tail -1 | sed "s/^/G=$G /" | tee -a gpurun_out/msm_sweep.log
  cp gpurun_out/msm_bench_v4.json gpurun_out/msm_bench_G$G.json
done
for K in 2 6; do for G in 4; do
  XHE_MSM_BUCKET_COST=$K XHE_MSM_GROUPS=$G timeout 300 python tools/msm_bench.py 16 17.2 17.6 18 20 22 2>&1 | tail -1 | sed "s/^/K=$K G=$G /" | tee -a gpurun_out/msm_sweep.log
done; done
