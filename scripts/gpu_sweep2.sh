#!/bin/bash
# usage: gpu_sweep2.sh "ENV1=a ENV2=b" "ENV1=c" ...  : bench (no cpu leg) per environment setting
mkdir -p gpurun_out
i=0
for e in "$@"; do
  i=$((i+1))
  env $e timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/sw_$i.json 2> gpurun_out/sw_$i.err
  python - <<PY
import json
try:
    r=json.load(open('gpurun_out/sw_$i.json'))
    a=r['also']['2a']
    print('$e | 2b %.1f us (1 stream %.1f) frac %.3f/%.3f | 2a %.1f us (1s %.1f) | e2e2b %.2f G/s single %.2f | e2e2a %.1f single %.1f' % (r['roofline']['kernel_ms']*1e3, r['roofline']['kernel_ms_single_stream']*1e3, r['roofline']['frac'], r['roofline']['frac_single_stream'], a['roofline']['kernel_ms']*1e3, a['roofline']['kernel_ms_single_stream']*1e3, r['e2e']['value']/1e9, r['e2e']['single_caller']['value']/1e9, a['e2e']['value']/1e9, a['e2e']['single_caller']['value']/1e9))
except Exception as ex:
    print('$e failed', ex); print(open('gpurun_out/sw_$i.err').read()[-1500:])
PY
done
