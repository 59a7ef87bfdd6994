#!/bin/bash
# staged2 with every publishing lane arriving: parity, racecheck (staged2 forced), one-stream time on 2b
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
BSG_PROBE_VARIANT=3 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r2b_san_racecheck_staged2.log 2>&1; echo "racecheck smoke (staged2 forced) rc=$?"; tail -4 gpurun_out/r2b_san_racecheck_staged2.log
timeout 600 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" "BSG_PROBE_VARIANT=3" > gpurun_out/r2b_sweep_2b.txt 2> gpurun_out/r2b_sweep_2b.err; echo "sweep 2b rc=$?"; cat gpurun_out/r2b_sweep_2b.txt; tail -3 gpurun_out/r2b_sweep_2b.err
