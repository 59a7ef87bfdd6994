#!/bin/bash
# compute-sanitizer over the small staged-probe cases (memcheck, then racecheck + synccheck on smoke()).
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_memcheck.log 2>&1; echo "memcheck smoke rc=$?"; tail -4 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_racecheck.log 2>&1; echo "racecheck smoke rc=$?"; tail -6 gpurun_out/san_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_synccheck.log 2>&1; echo "synccheck smoke rc=$?"; tail -4 gpurun_out/san_synccheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants and (3-4-700 or 3-0-700 or 3-1-700)" > gpurun_out/san_memcheck_variants.log 2>&1; echo "memcheck variants rc=$?"; tail -4 gpurun_out/san_memcheck_variants.log
