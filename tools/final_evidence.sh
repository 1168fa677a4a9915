#!/bin/bash
# one GPU-box pass that produces the files summarised under profiles/ (run through gpurun; every ncu pass follows a plain
# run of the same command that exited 0)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/final_smi.txt
timeout 400 python bench.py > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference rc=$?"
timeout 200 python tools/msm_bench.py 16 18 20 22 > gpurun_out/final_msm_bench.log 2>&1; echo "msm rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --inflight 1"
timeout 200 $CMD > gpurun_out/final_plain.json 2> gpurun_out/final_plain.err; echo "plain rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ncu_launch.json 2> gpurun_out/final_ncu_launch.err; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_decompress|k_msm_accum_tiles|k_sig_r|k_rp_gens|k_fb_buckets|k_msm_horner|k_msm_seg|k_rp_prep" --launch-skip 16 -c 10 -o gpurun_out/final_full -f $CMD > gpurun_out/final_ncu_full.json 2> gpurun_out/final_ncu_full.err; echo "ncu full rc=$?"
timeout 200 ncu -i gpurun_out/final_full.ncu-rep --page raw --csv > gpurun_out/final_full_raw.csv 2> /dev/null; echo "export rc=$?"
rm -f gpurun_out/final_full.ncu-rep
ls -la gpurun_out/final_*
