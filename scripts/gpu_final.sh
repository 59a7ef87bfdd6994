#!/bin/bash
# final evidence of round 2 on one GPU, most important first: all GPU tests, bench (both arms), ncu --set full of the 2b kernel
# (the one that changed last), launch list of a timed step, racecheck over the staged-kernel parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -4 gpurun_out/r02_pytest_all.log
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_ref.json
S=$(date +%s); timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"; cut -c1-400 gpurun_out/r02_bench_1gpu.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:probe_staged2 -s 6 -c 1 -o gpurun_out/r02_ncu_2b -f python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r02_ncu_2b.log 2>&1; echo "ncu 2b rc=$?"; tail -2 gpurun_out/r02_ncu_2b.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"; tail -2 gpurun_out/r02_launches.csv | cut -c1-200
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged_variants and not 0-" > gpurun_out/r02_san_racecheck_staged2_tests.log 2>&1; echo "racecheck staged2 tests rc=$?"; tail -3 gpurun_out/r02_san_racecheck_staged2_tests.log
