#!/bin/bash
# One gpurun call: GPU parity tests, bench line, then the ncu launch list of the same bench command.
# usage: tools/gpu_round.sh <tag> [extra pytest args]
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
echo "ncu rc=$?"
