#!/bin/bash
# A/B of k_emit_frames prefetch variants (tools/build_variants.py) on C3 and C5
python tools/enc_kernels_ab.py 2>&1 | tail -2
for v in l1a l1b l2a l2b; do
FRB_LIB_PATH=flac_raster_b200/lib/var_$v.so python tools/enc_kernels_ab.py 2>&1 | tail -2
done
