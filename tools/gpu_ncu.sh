#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> <kernel regex> <count> [skip]
tag=$1; rx=$2; cnt=${3:-2}; skip=${4:-0}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$tag.log
