#!/bin/bash
# decode: CRC kernel launched after (default) or before (FRB_CRC_FIRST=1) the decode kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "decode or fused or legacy or boundary or config" > gpurun_out/pytest_gpu_r3z.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r3z.log
for cfg in "c3 0" "c3 1" "c5 0" "c5 1"; do
set -- $cfg
FRB_CRC_FIRST=$2 python bench.py --workload $1 --steps 5 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1 crc_first=$2 enc', round(d['ms_per_step'],3), 'dec', round(d['decode']['ms_per_step'],3), {k: round(v,3) for k,v in d['decode']['kernels_ms'].items()}, 'foreign', round(d['decode']['foreign_streams']['ms_per_step'],3), d['decode']['foreign_streams']['lossless'], d['lossless_roundtrip_checked'])"
done
