#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_shape" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
timeout 900 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" "BSG_TILES_SHAPE=5" "BSG_TILE_MIN_STAGES=2 BSG_TILE_BYTES=80000" "BSG_TILES_SHAPE=5 BSG_TILE_MIN_STAGES=2 BSG_TILE_BYTES=80000" "BSG_TILE_BYTES=40000" "BSG_TILE_MIN_STAGES=4" "BSG_TILE_MIN_STAGES=2 BSG_TILE_BYTES=70000 BSG_TILE_UNITS=8" > gpurun_out/r2g_sweep_2a.txt 2> gpurun_out/r2g_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2g_sweep_2a.txt; tail -3 gpurun_out/r2g_sweep_2a.err
