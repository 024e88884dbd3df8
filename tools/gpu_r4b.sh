#!/bin/bash
# N = 2 strong-scaling bench line (tiles split by pixel count), default flags minus the CPU baseline
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_r4b.json 2> gpurun_out/bench_n2_r4b.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_r4b.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), d['rank_ms'], d['config']['tiles_rank0'], 'dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['container_check'], d['c5_bbox']['value'], d['c5_bbox']['ok'], d['lossless_roundtrip_checked'])
PY
tail -3 gpurun_out/bench_n2_r4b.err
