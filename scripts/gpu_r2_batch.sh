#!/bin/bash
# query batching: parity tests + the bench's probe leg (small-query numbers)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi or batcher or concurrent or reference_cases" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2e_pytest.log
timeout 900 python bench.py --no-extra --no-cpu > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2e_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d["e2e"].get("small_queries"), indent=1))
print("value", d["value"], "roofline", d["roofline"]["kernel_us"], d["roofline"]["frac"], "2a", d["also"]["2a"]["roofline"]["kernel_us"], "e2e", d["e2e"]["value"], d["e2e"]["single_caller"])
PY
