#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles or beyond or boundaries or corruption or hierarchical" > gpurun_out/r2_pytest_tiles.log 2>&1; echo "pytest(tiles) rc=$?"; tail -5 gpurun_out/r2_pytest_tiles.log
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=3" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=1" "BSG_TILES_SHAPE=2" "BSG_TILES_SHAPE=1 BSG_TILE_UNITS=1 BSG_TILE_MIN_STAGES=2" "BSG_TILES_SHAPE=0 BSG_PROBE_PDL=0" > gpurun_out/r2_sweep_2b.txt 2> gpurun_out/r2_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2_sweep_2b.txt; tail -3 gpurun_out/r2_sweep_2b.err
timeout 900 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=3" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=1" "BSG_TILES_SHAPE=2" "BSG_TILES_SHAPE=3" "BSG_TILES_SHAPE=1 BSG_TILE_BYTES=60000" "BSG_TILES_SHAPE=0 BSG_TILE_BYTES=16384" "BSG_TILES_SHAPE=0 BSG_TILE_BYTES=24000" "BSG_TILES_SHAPE=2 BSG_TILE_BYTES=24000" > gpurun_out/r2_sweep_2a.txt 2> gpurun_out/r2_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2_sweep_2a.txt; tail -3 gpurun_out/r2_sweep_2a.err
timeout 200 python scripts/trace_tiles.py 2b > gpurun_out/r2_trace_2b.txt 2>&1; tail -2 gpurun_out/r2_trace_2b.txt
timeout 200 python scripts/trace_tiles.py 2a > gpurun_out/r2_trace_2a.txt 2>&1; tail -2 gpurun_out/r2_trace_2a.txt
