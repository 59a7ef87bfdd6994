#!/bin/bash
# usage: gpu_sweep_env.sh VAR v1 v2 ...   (correctness first, then bench per value)
mkdir -p gpurun_out
var=$1; shift
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for w in "$@"; do
  env $var=$w timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/sweep_$w.json 2> gpurun_out/sweep_$w.err
  python - <<PY
import json
try:
    r=json.load(open('gpurun_out/sweep_$w.json'))
    a=r['also']['2a']
    print('$var=$w 2b %.1f us frac %.3f | 2a %.1f us frac %.3f | e2e2b %.2f G/s' % (r['roofline']['kernel_ms']*1e3, r['roofline']['frac'], a['roofline']['kernel_ms']*1e3, a['roofline']['frac'], r['e2e']['value']/1e9))
except Exception as e:
    print('$var=$w failed', e); print(open('gpurun_out/sweep_$w.err').read()[-800:])
PY
done
