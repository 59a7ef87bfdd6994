#!/bin/bash
# probe_staged2 shapes (BSG_PROBE_VARIANT): parity of every shape, bench per shape, timeline, full GPU suite.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/pytest_shapes.log 2>&1; echo "pytest(variants) rc=$?"; tail -5 gpurun_out/pytest_shapes.log
for v in ${VARIANTS:-1 2 3 4 5}; do
  env BSG_PROBE_VARIANT=$v timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/shape_$v.json 2> gpurun_out/shape_$v.err
  python - <<PY
import json
try:
    r=json.load(open('gpurun_out/shape_$v.json'))
    a=r['also']['2a']
    print('variant=$v 2b %.1f us (1 stream %.1f) frac %.3f/%.3f | 2a %.1f us (1s %.1f) | e2e2b %.2f G/s single %.2f' % (r['roofline']['kernel_ms']*1e3, r['roofline']['kernel_ms_single_stream']*1e3, r['roofline']['frac'], r['roofline']['frac_single_stream'], a['roofline']['kernel_ms']*1e3, a['roofline']['kernel_ms_single_stream']*1e3, r['e2e']['value']/1e9, r['e2e']['single_caller']['value']/1e9))
except Exception as e:
    print('variant=$v failed', e); print(open('gpurun_out/shape_$v.err').read()[-1500:])
PY
done
env BSG_PROBE_VARIANT=${TRACEV:-3} timeout 150 python scripts/trace_probe2.py 2b > gpurun_out/trace3_2b.txt 2>&1; tail -8 gpurun_out/trace3_2b.txt
env BSG_PROBE_VARIANT=${TRACEV:-3} timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest(all, variant 3) rc=$?"; tail -5 gpurun_out/pytest.log
