#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -6 gpurun_out/r2_pytest_all.log
BSG_DISTINCT_HASH_BITS=10 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "distinct or counted" > gpurun_out/r2_pytest_distinct_collide.log 2>&1; echo "pytest(distinct, forced collisions) rc=$?"; tail -3 gpurun_out/r2_pytest_distinct_collide.log
timeout 300 python scripts/bench_distinct.py > gpurun_out/r2_distinct.txt 2>&1; tail -5 gpurun_out/r2_distinct.txt
S=$(date +%s); timeout 1500 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; grep "bench\]" gpurun_out/r2_bench_1gpu.err | tail -12; tail -5 gpurun_out/r2_bench_1gpu.err; head -c 6000 gpurun_out/r2_bench_1gpu.json
