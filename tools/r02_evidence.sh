#!/bin/bash
# round-2 evidence pass (run through gpurun; every ncu pass follows a plain run of the same command that exited 0)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r02_smi.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --secondary off --inflight 1"
timeout 300 $CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err; echo "plain rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launch.json 2> gpurun_out/r02_ncu_launch.err; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_decompress|k_msm_accum_tiles|k_sig_r|k_rp_gens|k_msm_horner_oct|k_msm_bucket_seg|k_msm_nodes_seq|k_msm_nodes32|k_rp_prep|k_fiat_shamir|k_msm_scatter|k_msm_count" --launch-skip 16 -c 14 -o gpurun_out/r02_full -f $CMD > gpurun_out/r02_ncu_full.json 2> gpurun_out/r02_ncu_full.err; echo "ncu full rc=$?"
timeout 200 ncu -i gpurun_out/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2> /dev/null; echo "export rc=$?"
rm -f gpurun_out/r02_full.ncu-rep
MCMD="python tools/msm_bench.py 20"
timeout 200 $MCMD > gpurun_out/r02_msm_plain.log 2>&1; echo "msm plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_msm_launches.csv $MCMD > gpurun_out/r02_msm_ncu.log 2>&1; echo "msm ncu rc=$?"
ls -la gpurun_out/r02_*
