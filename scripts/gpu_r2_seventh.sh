#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -6 gpurun_out/r2_pytest_all.log
timeout 300 python scripts/e2e_mask_1m.py > gpurun_out/r2_e2e_mask_1m.txt 2>&1; tail -6 gpurun_out/r2_e2e_mask_1m.txt
timeout 300 python scripts/e2e_phases.py 2b 2a > gpurun_out/r2_e2e_phases.txt 2>&1; tail -6 gpurun_out/r2_e2e_phases.txt
BSG_PROBE_SPIN=0 timeout 300 python scripts/e2e_phases.py 2b > gpurun_out/r2_e2e_phases_nospin.txt 2>&1; tail -3 gpurun_out/r2_e2e_phases_nospin.txt
/usr/bin/time -v timeout 1200 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -25 gpurun_out/r2_bench_1gpu.err | grep -v "^\s" ; head -c 3000 gpurun_out/r2_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; head -c 1500 gpurun_out/r2_bench_ref.json
