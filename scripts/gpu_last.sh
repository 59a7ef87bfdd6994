#!/bin/bash
# last call of round 2: the tile kernel's round A lost 7 instructions per (key, unit) and mod_m32 one — full parity first;
# only if it is green: sweeps, bench (both arms already measured: the GPU arm only), ncu of both layouts, launch list
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_all.log 2>&1; rc=$?; echo "pytest(all) rc=$rc"; tail -4 gpurun_out/r02b_pytest_all.log
if [ $rc -ne 0 ]; then tail -60 gpurun_out/r02b_pytest_all.log; exit 1; fi
timeout 120 python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" > gpurun_out/r02b_sweep_2a.txt 2> gpurun_out/r02b_sweep_2a.err; cat gpurun_out/r02b_sweep_2a.txt
timeout 120 python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r02b_sweep_2b.txt 2> gpurun_out/r02b_sweep_2b.err; cat gpurun_out/r02b_sweep_2b.txt
S=$(date +%s); timeout 600 python bench.py > gpurun_out/r02b_bench_1gpu.json 2> gpurun_out/r02b_bench_1gpu.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r02b_bench_1gpu.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:probe_tiles -s 6 -c 1 -o gpurun_out/r02b_ncu_2a -f python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" > gpurun_out/r02b_ncu_2a.log 2>&1; echo "ncu 2a rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:probe_staged2 -s 6 -c 1 -o gpurun_out/r02b_ncu_2b -f python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r02b_ncu_2b.log 2>&1; echo "ncu 2b rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02b_launches.log 2>&1; echo "launch list rc=$?"
