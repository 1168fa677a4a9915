#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_msm.py tests/test_golden.py -q 2>&1 | tail -3
for cfg in "default" "XHE_MSM_SEG_LOG=2" "XHE_MSM_SEG_LOG=4" "XHE_MSM_SMEM_ACC=1" "XHE_MSM_GROUPS=4" "XHE_MSM_CHAIN=1"; do
  if [ "$cfg" = "default" ]; then env_cmd=""; else env_cmd="env $cfg"; fi
  $env_cmd timeout 300 python tools/msm_bench.py 16 17.2 17.6 18 20 22 2>&1 | tail -1 | sed "s/^/$cfg /" | tee -a gpurun_out/r02_msm_ab.log
done
