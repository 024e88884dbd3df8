for lib in libflacraster_b200.so libfrb_pos.so; do
for w in c3 c5; do
FRB_LIB_PATH=/root/repo/flac_raster_b200/lib/$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $w dec', round(d['decode']['value'],1), round(d['decode']['ms_per_step'],2), d['decode']['kernels_ms'], d['lossless_roundtrip_checked'])"
done; done
