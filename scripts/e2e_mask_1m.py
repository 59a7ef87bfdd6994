#!/usr/bin/env python
"""GPU-box script: the flat mask-only bsg_probe() call over 1 M units that round 1 measured at 15 ms
(VERDICT weak #4): host phase times (BSG_PROBE_TIMING=1) and wall time per call."""
import os
import sys
import time

import numpy as np

os.environ["BSG_PROBE_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bloomsearch_b200 as bs  # noqa: E402
from bloomsearch_b200 import _native as N  # noqa: E402
from synth.corpus import SynthCorpus  # noqa: E402

ctx = bs.Context(0)
c2 = SynthCorpus(42, 0, 10000, 1000, 100)
d2, nw2 = bench.size_filters(c2, bs)
w2 = ctx.build(c2.blob, c2.key_off, c2.group_begin, np.arange(len(d2), dtype=np.uint32), None, d2, nw2)
corpus = bs.Corpus(ctx, np.tile(d2, 100), w2)
ft = lambda b, j: c2.key(int(c2.group_begin[3 * b + 2]) + j)
split = lambda key: key.split(b"::", 1)
keys8 = [ft(17, 3), b"level::nope", ft(4021, 1500), b"user_id::x1", b"service::auth", b"service::nosuch", b"level::info", b"nested.az::az-1"]
q = bs.BloomQuery(bs.And(bs.Or(*[bs.FieldToken(*split(k)) for k in keys8[:4]]), bs.Or(*[bs.FieldToken(*split(k)) for k in keys8[4:6]]),
                         bs.FieldToken(*split(keys8[6])), bs.FieldToken(*split(keys8[7]))))
cq = bs.compile_bloom_query(q)
blob, off = N.pack_keys(cq.keys)
kinds = np.ascontiguousarray(cq.kinds, dtype=np.uint8)
mask = np.zeros((corpus.n_units + 63) // 64, dtype=np.uint64)
for _ in range(3):
    corpus.probe_packed(blob, off, kinds, cq.prog, None, mask)
t = time.perf_counter()
n = 20
for _ in range(n):
    corpus.probe_packed(blob, off, kinds, cq.prog, None, mask)
print("bsg_probe mask-only over %d units: %.3f ms per call, %d survivors" % (corpus.n_units, (time.perf_counter() - t) / n * 1e3,
                                                                            int(bs.unpack_mask(mask, corpus.n_units).sum())))
t = time.perf_counter()
for _ in range(n):
    corpus.probe(cq.keys, cq.kinds, cq.prog, want_matrix=False)
print("Corpus.probe (python packing incl.): %.3f ms per call" % ((time.perf_counter() - t) / n * 1e3))
corpus.close()
ctx.close()
