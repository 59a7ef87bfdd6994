#!/bin/bash
# Round-end evidence: GPU tests, ncu captures + launch list, timelines, extra configs, default bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
bash scripts/gpu_ncu.sh 2a 2b > gpurun_out/ncu_run.log 2>&1; tail -2 gpurun_out/ncu_run.log | cut -c1-200
for w in 2b 2a; do timeout 150 python scripts/trace_probe2.py $w 2>&1 | grep -v Warning | grep -v "^\[bench\]" > gpurun_out/trace_final_$w.txt; tail -3 gpurun_out/trace_final_$w.txt; done
timeout 600 python scripts/bench_extra.py > gpurun_out/extra.json 2> gpurun_out/extra.err; echo "extra rc=$?"; tail -c 1500 gpurun_out/extra.json
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
