#!/bin/bash
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02n_bench_default.json 2> gpurun_out/r02n_bench_default.err; echo "default bench rc=$? in $(( $(date +%s) - S )) s"
tail -2 gpurun_out/r02n_bench_default.err
bash tools/r02_evidence.sh
