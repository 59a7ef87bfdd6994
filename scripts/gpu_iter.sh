#!/bin/bash
# iteration check: tile-kernel parity tests, 2a / 2b sweep of the default dispatch, ncu --set full of the 2a kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tiles or beyond or boundaries or hierarchical or pinned or smoke or sections" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
timeout 600 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" ${SWEEP_2A_EXTRA} > gpurun_out/r2d_sweep_2a.txt 2> gpurun_out/r2d_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2d_sweep_2a.txt; tail -3 gpurun_out/r2d_sweep_2a.err
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r2d_sweep_2b.txt 2> gpurun_out/r2d_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2d_sweep_2b.txt; tail -3 gpurun_out/r2d_sweep_2b.err
if [ -n "$NCU_2A" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_tiles -s 6 -c 1 -o gpurun_out/r2d_ncu_2a -f python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" > gpurun_out/r2d_ncu_2a.log 2>&1; echo "ncu 2a rc=$?"; tail -2 gpurun_out/r2d_ncu_2a.log
fi
