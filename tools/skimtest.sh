python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for L in 32 16; do
  FRB_SKIM_LANES=$L python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('lanes $L decode', d['decode']['ms_per_step'], 'lossless', d['lossless_roundtrip_checked'])"
done
