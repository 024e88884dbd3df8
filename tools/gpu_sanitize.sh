#!/bin/bash
# memcheck on the small parity cases (one tool per gpurun call)
python -m pytest tests -m gpu -x -q -k "encode_matches_oracle or decode_oracle_encoded or tile_mapping or golden_rgb or host_pipeline" > gpurun_out/plain_sanitize.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_sanitize.log; exit 1; }
tail -1 gpurun_out/plain_sanitize.log
FRB_TEST_TILES=64 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -x -q -k "encode_matches_oracle or decode_oracle_encoded or tile_mapping or golden_rgb or host_pipeline" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
