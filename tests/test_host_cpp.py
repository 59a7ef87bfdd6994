"""The C++ host layer (bloomsearch_b200/host/) run through its self-test, which mirrors the
reference's own tests (TestEvaluateBloomFilters, filter sizing, section round trip, Example)."""
from __future__ import annotations

import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bloomsearch_b200", "_build", "host_selftest")
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "bloom_golden.json")))["section"]["hex"]


def _run(*args):
    assert os.path.exists(EXE), "run __graft_entry__.build() first"
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=300)


def test_host_layer_without_gpu():
    r = _run("--host-only", "--golden-section", GOLDEN)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_selftest ok (host only)" in r.stdout


@pytest.mark.gpu
def test_host_layer_reference_cases_on_gpu():
    r = _run("--golden-section", GOLDEN)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_selftest ok (host + gpu)" in r.stdout
    # the section the C++ layer wrote for GPU-built filters equals the oracle's bytes
    assert f"section_hex {GOLDEN}" in r.stdout
