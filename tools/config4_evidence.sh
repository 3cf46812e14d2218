#!/bin/bash
# One gpurun call that reproduces profiles/config4_r02.md: the MobileNetV2 GPU tests, the config-4 bench line with its
# per-kernel HBM table, and the per-entry-point time table.  Usage (from the repo root, after `python __graft_entry__.py`):
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/config4_evidence.sh'
# SNIPER_DW_TILED=0 selects the register-window depthwise kernels (the v3 column of the profile).
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_zz_mobilenet_gpu.py -q -s > gpurun_out/mnv2_tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed" gpurun_out/mnv2_tests.log | tail -2
SNIPER_BREAKDOWN=gpurun_out/entry_points.md timeout 120 python bench.py --config4 40 --steps 8 --warmup 3 \
  > gpurun_out/config4.json 2> gpurun_out/config4.err
echo "bench rc=$?"; head -c 300 gpurun_out/config4.json; echo
