#!/bin/bash
# what would an index (frame positions + subframe offsets known up front) buy the decode step?
for e in 0 1; do
FRB_EXP_REUSE_INDEX=$e python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reuse=$e', 'dec', round(d['decode']['ms_per_step'],2), {k: round(x,2) for k,x in d['decode']['kernels_ms'].items()}, d['lossless_roundtrip_checked'])"
done
