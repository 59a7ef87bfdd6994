#!/bin/bash
# round 2 evidence on one GPU: ncu --set full of the kernels the bench launches, launch list, sanitizers, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_all.log 2>&1; echo "pytest(all) rc=$?"; tail -4 gpurun_out/r02_pytest_all.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_staged2 -s 6 -c 1 -o gpurun_out/r02_ncu_2b -f python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=7" > gpurun_out/r02_ncu_2b.log 2>&1; echo "ncu 2b rc=$?"; tail -2 gpurun_out/r02_ncu_2b.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_tiles -s 6 -c 1 -o gpurun_out/r02_ncu_2a -f python scripts/sweep_tiles.py 2a "BSG_PROBE_VARIANT=7" > gpurun_out/r02_ncu_2a.log 2>&1; echo "ncu 2a rc=$?"; tail -2 gpurun_out/r02_ncu_2a.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_build_file -f python scripts/run_build.py 2000 file > gpurun_out/r02_ncu_build_file.log 2>&1; echo "ncu build(file) rc=$?"; tail -2 gpurun_out/r02_ncu_build_file.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_build_blocks -f python scripts/run_build.py 2000 blocks > gpurun_out/r02_ncu_build_blocks.log 2>&1; echo "ncu build(blocks) rc=$?"; tail -2 gpurun_out/r02_ncu_build_blocks.log
python scripts/run_build.py 2000 file; python scripts/run_build.py 2000 blocks
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"; tail -3 gpurun_out/r02_launches.csv | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_san_memcheck.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/r02_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_san_racecheck.log 2>&1; echo "racecheck smoke rc=$?"; tail -4 gpurun_out/r02_san_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_san_synccheck.log 2>&1; echo "synccheck smoke rc=$?"; tail -3 gpurun_out/r02_san_synccheck.log
BSG_PROBE_VARIANT=6 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/r02_san_racecheck_tiles.log 2>&1; echo "racecheck smoke (tile kernel forced) rc=$?"; tail -4 gpurun_out/r02_san_racecheck_tiles.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "masked_fills or every_shape" > gpurun_out/r02_san_memcheck_tiles.log 2>&1; echo "memcheck tiles rc=$?"; tail -3 gpurun_out/r02_san_memcheck_tiles.log
timeout 1200 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02_bench_ref.json
