#!/bin/bash
# full GPU test suite, then the quick encode/decode line on C3 and C5 for both frame-assembly group sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r3x.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r3x.log
for cfg in "c3 0" "c3 32" "c5 0" "c5 128"; do
set -- $cfg
FRB_EMIT_TPF=$2 python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 tpf=$2', round(d['value'],1), round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['kernels_ms'].items()}, d['lossless_roundtrip_checked'], 'dec', round(d['decode']['ms_per_step'],3), {k: round(v,3) for k,v in d['decode']['kernels_ms'].items()}, 'foreign', round(d['decode']['foreign_streams']['ms_per_step'],3))"
done
