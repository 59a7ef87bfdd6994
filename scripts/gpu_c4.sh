#!/bin/bash
# config 4 alone (the hierarchical query): parity of the short-circuiting gather kernel, then the leg with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "short_circuit or hierarchical or reference_cases or multi or batcher or example or absent or large_filters" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
for cfg in "1 64" "0 64" "1 32"; do
set -- $cfg
BSG_PROBE_SHORT_CIRCUIT=$1 BSG_L2_FETCH=$2 timeout 600 python bench.py --leg config4 > gpurun_out/r2k_c4_$1_$2.json 2> gpurun_out/r2k_c4_$1_$2.err; echo "c4 short_circuit=$1 l2fetch=$2 rc=$?"; python - <<PY
import json
for ln in open('gpurun_out/r2k_c4_$1_$2.err').read().splitlines() + open('gpurun_out/r2k_c4_$1_$2.json').read().splitlines():
    if ln.startswith('{"leg"'):
        d=json.loads(ln); print('ms_per_query', d['ms_per_query'], 'e2e', d['e2e']['ms_per_query'], 'survivors', d['surviving_blocks_this_rank'], 'oracle_checked', d.get('oracle_checked_units'))
PY
done
