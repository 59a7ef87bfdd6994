#!/bin/bash
# one-step 32-bit modulo + new tile descriptor encoding: parity, sweeps on both layouts, build kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest.log
timeout 600 python scripts/sweep_tiles.py 2a "BSG_TILES_SHAPE=1" "BSG_TILES_SHAPE=5" "BSG_PROBE_VARIANT=3" "BSG_TILES_SHAPE=3" "BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=6" > gpurun_out/r2c_sweep_2a.txt 2> gpurun_out/r2c_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2c_sweep_2a.txt; tail -3 gpurun_out/r2c_sweep_2a.err
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=1" "BSG_PROBE_VARIANT=6 BSG_TILES_SHAPE=5" > gpurun_out/r2c_sweep_2b.txt 2> gpurun_out/r2c_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2c_sweep_2b.txt; tail -3 gpurun_out/r2c_sweep_2b.err
python scripts/run_build.py 2000 file; python scripts/run_build.py 2000 blocks
