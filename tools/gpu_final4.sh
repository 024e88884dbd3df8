#!/bin/bash
# C5 and C4 bench lines of the final build (no CPU baseline, no extras)
mkdir -p gpurun_out
t0=$SECONDS
timeout 30 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c5_final.json 2> gpurun_out/bench_c5_final.err; echo "c5 rc=$? in $((SECONDS-t0)) s"; cut -c1-200 gpurun_out/bench_c5_final.json
t0=$SECONDS
timeout 75 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c4_final.json 2> gpurun_out/bench_c4_final.err; echo "c4 rc=$? in $((SECONDS-t0)) s"; cut -c1-200 gpurun_out/bench_c4_final.json
