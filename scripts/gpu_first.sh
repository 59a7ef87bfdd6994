#!/bin/bash
# First GPU validation: smoke, sanitizer on smoke, gpu tests, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -8 gpurun_out/sanitizer.log
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -15 gpurun_out/bench.err; cat gpurun_out/bench.json
