#!/bin/bash
# the driver's round-end sequence: full GPU test suite, smoke, the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
/usr/bin/time -v python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), d['kernels_ms'], 'dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, d['c5_bbox']['value'], d['roofline']['frac'], d['gpu_launches'])
PY
