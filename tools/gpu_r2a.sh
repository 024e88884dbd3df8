#!/bin/bash
# round 2, first GPU call: all GPU tests (not -x: collect every failure), bench line, reference arm
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -c "import os; print('cpus', len(os.sched_getaffinity(0)))"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_$tag.json
