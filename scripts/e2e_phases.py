#!/usr/bin/env python
"""Host-side phase times of bsg_probe() (BSG_PROBE_TIMING=1), single caller, layout 2b / 2a."""
import os, sys, time
import numpy as np
os.environ["BSG_PROBE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
import bench

for wl in sys.argv[1:] or ["2b"]:
    ctx = bs.Context(0)
    c = bench.gen_corpus(wl, 0)
    desc, n_words = bench.size_filters(c, bs)
    words = ctx.build(c.blob, c.key_off, c.group_begin, np.arange(len(desc), dtype=np.uint32), None, desc, n_words)
    keys, kinds = bench.make_batch(c, 7)
    blob, off = N.pack_keys(keys)
    corpora = [bs.Corpus(ctx, desc, words) for _ in range(4)]
    out = np.zeros((c.n_blocks, (len(keys) + 63) // 64), dtype=np.uint64)
    for i in range(20):
        corpora[i % 4].probe_packed(blob, off, kinds, None, out, None)
    ctx.close()  # prints + resets nothing; open a fresh context for the measured calls
    ctx = bs.Context(0)
    corpora = [bs.Corpus(ctx, desc, words) for _ in range(4)]
    for i in range(10):
        corpora[i % 4].probe_packed(blob, off, kinds, None, out, None)
    t0 = time.perf_counter()
    n = 300
    for i in range(n):
        corpora[i % 4].probe_packed(blob, off, kinds, None, out, None)
    dt = time.perf_counter() - t0
    print(f"{wl}: {dt / n * 1e6:.1f} us per bsg_probe call (python caller, single thread)", flush=True)
    for cp in corpora:
        cp.close()
    ctx.close()
