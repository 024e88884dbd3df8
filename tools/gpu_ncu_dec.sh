#!/bin/bash
# ncu --set full of the decode kernel: launch 1 = with skim, launch 2 = tables reused (what an index would give)
tag=${1:-dec}
mkdir -p gpurun_out
export FRB_EXP_REUSE_INDEX=${REUSE:-1}
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/plain_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode_subframes -c 2 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$tag.log
