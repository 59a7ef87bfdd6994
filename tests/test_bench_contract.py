"""CPU tests of bench.py's contract lines that need no GPU: the reference arm (the C restatement of the Go
path on the host cores), what it may import, how it behaves under a multi-rank launch, and the handler that
keeps the headline line when a later leg fails."""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_line():
    o = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert o.returncode == 0, o.stderr[-2000:]
    lines = [ln for ln in o.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # ONE JSON line on stdout, nothing else
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == "bloom probes/sec (block-level)" and d["unit"] == "probes/s"
    assert d["config"]["workload"] == bench.workload_string("2b")      # same string as the GPU arm's (same_config)
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["warmup"] >= 3                                            # the W >= 3 rule is enforced, not trusted
    assert d["value"] > 0 and abs(d["value"] - 64 * 1000 * 1000 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and cb["sample"]
    assert cb["probe_only"]["value"] > d["value"]                      # decode + probe costs more than probe only
    assert d["e2e"] == {"value": d["value"], "unit": "probes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_without_work():
    o = _run(["--impl", "reference", "--gpus", "2", "--steps", "1"], {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"},
             timeout=120)
    assert o.returncode == 0 and o.stdout.strip() == ""


def test_reference_arm_imports_nothing_of_the_product():
    code = ("import sys, runpy\n"
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '1', '--gpus', '2']\n"
            "import os; os.environ.update(RANK='1', WORLD_SIZE='2')\n"   # rank 1: the import graph without the work
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "bad = [m for m in sys.modules if m.startswith('bloomsearch_b200')]\n"
            "print('PRODUCT_MODULES', bad, file=sys.stderr)\n")
    o = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert o.returncode == 0, o.stderr[-2000:]
    assert "PRODUCT_MODULES []" in o.stderr
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src[src.index("def run_reference"):src.index("_JSON_FD = None")]
    assert "bloomsearch_b200" not in body and "import bs" not in body


def test_pending_headline_is_printed_once_with_the_failure_named():
    code = ("import bench\n"
            "bench._PENDING = {'metric': 'm', 'value': 1.0}\n"
            "bench._emit_pending('RuntimeError: leg exploded')\n"
            "bench._emit_pending('again')\n")
    o = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    lines = [ln for ln in o.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.0, "extra_legs_error": "RuntimeError: leg exploded"}
