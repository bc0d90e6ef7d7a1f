#!/bin/bash
# compute-sanitizer over small parity tests: memcheck (out-of-bounds / misaligned) and racecheck (shared-memory hazards
# of the warp kernel's per-warp buffers and the block reductions).  Small cases only: the tools slow kernels 10-100x.
mkdir -p gpurun_out
SEL='ez_kats or readme or ragged_rows_with_gaps or long_rows or duplicate or single_row or one_by_one or matrix_without or aprod_modes'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Misaligned" gpurun_out/sanitize_memcheck.log | head -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests -m gpu -q -x -k "ez_kats or readme or long_rows or duplicate or single_row" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head -8
