#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "distinct or counted or keyset" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
timeout 900 python scripts/bench_distinct.py > gpurun_out/r2h_distinct.json 2> gpurun_out/r2h_distinct.err; echo "bench_distinct rc=$?"; cat gpurun_out/r2h_distinct.json; tail -3 gpurun_out/r2h_distinct.err
