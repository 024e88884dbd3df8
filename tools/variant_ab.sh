#!/bin/bash
# A/B of library variants built into flac_raster_b200/lib (usage: tools/variant_ab.sh libA.so libB.so ...)
for lib in "$@"; do
for w in c3 c5; do
FRB_LIB_PATH=/root/repo/flac_raster_b200/lib/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $w enc', round(d['value'],1), 'dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],2), d['decode']['kernels_ms'], d['lossless_roundtrip_checked'])"
done; done
