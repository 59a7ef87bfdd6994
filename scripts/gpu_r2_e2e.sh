#!/bin/bash
# e2e leg with C caller threads: 4 vs 8 callers
mkdir -p gpurun_out
for c in 4 8 6; do
timeout 600 python bench.py --no-extra --no-cpu --no-also --e2e-callers $c > gpurun_out/r2f_bench_c$c.json 2> gpurun_out/r2f_bench_c$c.err; echo "bench callers=$c rc=$?"; tail -2 gpurun_out/r2f_bench_c$c.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2f_bench_c$c.json').read().strip().splitlines()[-1])
e=d["e2e"]; print("callers", e["callers"], "e2e %.3e" % e["value"], "us/call/caller %.1f" % e["us_per_call_per_caller"], e["single_caller"], "value %.3e" % d["value"])
PY
done
