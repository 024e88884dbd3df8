#!/bin/bash
# ncu launch list of this library's kernels over a short bench run, and the full-set capture of the frame assembly kernel
mkdir -p gpurun_out
t0=$SECONDS
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_list_final.log 2>&1; echo "ncu list rc=$? in $((SECONDS-t0)) s"
t0=$SECONDS
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"k_emit_frames" -s 1 -c 1 -o gpurun_out/prof_final_emit -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final_emit.log 2>&1; echo "ncu full rc=$? in $((SECONDS-t0)) s"
ls -la gpurun_out/prof_final_emit.ncu-rep
