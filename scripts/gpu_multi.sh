#!/bin/bash
# 2-GPU (or N-GPU) validation: device collectives over peer memory, hierarchical gather, the bench legs
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; head -12 gpurun_out/r2_topo.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_multi.log 2>&1; echo "pytest(multi) rc=$?"; tail -8 gpurun_out/r2_pytest_multi.log
BSG_COMM_P2P=0 timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_multi_nccl.log 2>&1; echo "pytest(multi, NCCL path) rc=$?"; tail -3 gpurun_out/r2_pytest_multi_nccl.log
S=$(date +%s); timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench N=$N rc=$? in $(( $(date +%s) - S )) s"; tail -4 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
r=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print('value %.3e e2e %.3e' % (r['value'], r['e2e']['value']), r.get('comm'))
for k in ('build','config5','config4'):
    d=r[k]; print(k, {kk: d[kk] for kk in d if kk not in ('workload','roofline','what')})
PY
BSG_COMM_P2P=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-also --no-cpu > gpurun_out/r2_bench_${N}gpu_nccl.json 2> gpurun_out/r2_bench_${N}gpu_nccl.err; echo "bench N=$N (NCCL path) rc=$?"
python - <<PY
import json
r=json.load(open('gpurun_out/r2_bench_${N}gpu_nccl.json'))
for k in ('config5','config4'):
    d=r[k]; print('nccl', k, {kk: d[kk] for kk in d if kk not in ('workload','roofline','what')})
PY
