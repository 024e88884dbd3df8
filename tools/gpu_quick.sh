#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for w in c3 c5; do
python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w enc', round(d['value'],1), round(d['ms_per_step'],2), 'dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],2), 'e2e', round(d['e2e']['value'],2), d['kernels_ms'], d['decode']['kernels_ms'], d['lossless_roundtrip_checked'])"
done
