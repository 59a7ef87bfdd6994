#!/bin/bash
# args: N (gpus). Runs the gpu tests (incl. NCCL ones) and the N-rank bench.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | head -c 1500
