#!/bin/bash
# k_emit_frames: register-cached segment (seg), + two-segment boundary chunks (seg2) against the current library
mkdir -p gpurun_out
python tools/enc_kernels_ab.py 2>&1 | tail -2
for v in seg seg2; do
FRB_LIB_PATH=flac_raster_b200/lib/var_$v.so python tools/enc_kernels_ab.py 2>&1 | tail -2
done
FRB_LIB_PATH=flac_raster_b200/lib/var_seg2.so timeout 600 python -m pytest tests -m gpu -x -q -k "encode or golden or roundtrip or mid_side or config or fullsize_c3 or handle or boundary or random" > gpurun_out/pytest_gpu_r4a.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r4a.log
