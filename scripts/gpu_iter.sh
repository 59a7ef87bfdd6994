#!/bin/bash
# iteration check: parity tests, racecheck of the staged kernels over their parity tests, 2a / 2b sweep of the default dispatch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2d_pytest.log
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" "BSG_PROBE_VARIANT=2" > gpurun_out/r2d_sweep_2b.txt 2> gpurun_out/r2d_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2d_sweep_2b.txt
timeout 600 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" "BSG_PROBE_VARIANT=3" > gpurun_out/r2d_sweep_2a.txt 2> gpurun_out/r2d_sweep_2a.err; echo "sweep 2a rc=$?"; cat gpurun_out/r2d_sweep_2a.txt
BSG_PROBE_VARIANT=3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_san_racecheck_staged2.log 2>&1; echo "racecheck smoke (staged2 forced) rc=$?"; tail -2 gpurun_out/r02_san_racecheck_staged2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged_variants and not 0-" > gpurun_out/r02_san_racecheck_staged2_tests.log 2>&1; echo "racecheck staged2 tests rc=$?"; tail -3 gpurun_out/r02_san_racecheck_staged2_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_shape or masked_fills or just_above" > gpurun_out/r02_san_racecheck_tiles_tests.log 2>&1; echo "racecheck tile tests rc=$?"; tail -3 gpurun_out/r02_san_racecheck_tiles_tests.log
