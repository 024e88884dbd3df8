#!/bin/bash
# decode-related GPU tests, then the quick bench line (no extras)
tag=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -k "decode or fused or fullsize or config or sharded or legacy or smoke or host_pipeline" > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$tag.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>gpurun_out/bench_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('enc', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['kernels_ms'].items()}, 'dec', round(d['decode']['ms_per_step'],2), {k: round(x,2) for k,x in d['decode']['kernels_ms'].items()}, d['lossless_roundtrip_checked'], 'e2e', round(d['e2e']['ms_per_step'],1), round(d['decode']['e2e']['ms_per_step'],1))"
