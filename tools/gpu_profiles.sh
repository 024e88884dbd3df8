#!/bin/bash
# the default bench line, then the ncu launch list and one full-set capture of the kernels this session changed
mkdir -p gpurun_out
t0=$SECONDS
timeout 240 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? in $((SECONDS-t0)) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), d['kernels_ms'], 'dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['cpu_baseline']['value'] if d.get('cpu_baseline') else None, d['c5_bbox']['value'], d['roofline']['frac'], d['gpu_launches'])
PY
t0=$SECONDS
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_list_final.log 2>&1; echo "ncu list rc=$? in $((SECONDS-t0)) s"
t0=$SECONDS
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_emit_frames|k_crc16_frames|k_decode_subframes" -s 2 -c 5 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$? in $((SECONDS-t0)) s"
ls -la gpurun_out/prof_final.ncu-rep
