python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for tpf in 0 32; do
for w in c3 c5; do
FRB_EMIT_TPF=$tpf python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tpf=$tpf $w enc', round(d['value'],1), round(d['ms_per_step'],2), d['kernels_ms'], d['lossless_roundtrip_checked'])"
done; done
FRB_EMIT_TPF=32 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
