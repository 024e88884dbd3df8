#!/bin/bash
# all GPU tests, every failure collected
tag=${1:-t}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 ${@:2} > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_$tag.log
