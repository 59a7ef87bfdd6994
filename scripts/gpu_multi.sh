#!/bin/bash
# N-GPU validation: device collectives tests + bench.py under torchrun (headline + build / config4 / config5 legs)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/r2_pytest_multi_${N}.log 2>&1; echo "pytest(multi) rc=$?"; tail -4 gpurun_out/r2_pytest_multi_${N}.log
S=$(date +%s); timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench N=$N rc=$? in $(( $(date +%s) - S )) s"; grep -E "Error|error|assert" gpurun_out/r2_bench_${N}gpu.err | head -5
python - <<PY
import json
r=json.load(open('gpurun_out/r2_bench_${N}gpu.json'))
print('value %.3e e2e %.3e' % (r['value'], r['e2e']['value']), r.get('comm'))
for k in ('build','config5','config4'):
    d=r[k]; print(k, {kk: d[kk] for kk in d if kk not in ('workload','roofline','what')})
PY
S=$(date +%s); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r2_bench_${N}gpu_ref.json 2> gpurun_out/r2_bench_${N}gpu_ref.err; echo "reference arm N=$N rc=$? in $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r2_bench_${N}gpu_ref.json
