#!/bin/bash
# memcheck on the fused decode launch (small parity cases)
K="fused_decode_to_raster or fused_integer_denormalise or decode_oracle_encoded or decode_reference_goldens or config2_rgb"
python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/plain_sanitize2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_sanitize2.log; exit 1; }
tail -1 gpurun_out/plain_sanitize2.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/sanitize2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize2_memcheck.log | head -20
