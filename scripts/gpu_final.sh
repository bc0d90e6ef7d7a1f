#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x -k "banded_transpose or row_blocked or column_blocked or ragged or ez_kats or balanced_tile_schedule_power or c5_full or underdetermined or without_entries" 2>&1 | tail -4
WLS=C3 timeout 150 bash scripts/gpu_families.sh 1
