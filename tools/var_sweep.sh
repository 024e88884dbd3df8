#!/bin/bash
# bench every variant library (and the default one): encode/decode step and kernel times, lossless flag
for v in flac_raster_b200/lib/var_*.so flac_raster_b200/lib/libflacraster_b200.so; do
FRB_LIB_PATH=$PWD/$v python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline --no-extras ${WORKLOAD:+--workload $WORKLOAD} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', 'enc', round(d['ms_per_step'],2), {k: round(x,2) for k,x in d['kernels_ms'].items()}, 'dec', round(d['decode']['ms_per_step'],2), {k: round(x,2) for k,x in d['decode']['kernels_ms'].items()}, d['lossless_roundtrip_checked'])"
done
