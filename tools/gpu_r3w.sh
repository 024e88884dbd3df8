#!/bin/bash
# full GPU test suite, then the quick encode line on C3 and C5 (kernel times)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r3w.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r3w.log
for w in c5 c3; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', round(d['value'],1), round(d['ms_per_step'],3), d['kernels_ms'], d['lossless_roundtrip_checked'], 'dec', round(d['decode']['ms_per_step'],3))"
done
