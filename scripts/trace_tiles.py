#!/usr/bin/env python
"""Profiling helper: per-CTA timeline of probe_tiles_kernel (TRACE instantiation): CTA start, hashes
ready, and per tile: resident / round A done / round B1 done / round B2 done (us), |L1|, |L2|."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bloomsearch_b200 as bs  # noqa: E402
from bloomsearch_b200 import _native as N  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "2b"
ctx = bs.Context(0)
c = bench.gen_corpus(wl, 0)
desc, n_words = bench.size_filters(c, bs)
words = ctx.build(c.blob, c.key_off, c.group_begin, np.arange(len(desc), dtype=np.uint32), None, desc, n_words)
keys, kinds = bench.make_batch(c, 7)
corpora = [bs.Corpus(ctx, desc, words) for _ in range(4)]
qs = [bs.Query(cp, keys, kinds, None) for cp in corpora]
for i in range(8):
    qs[i % 4].run(N.PROBE_STAGED | N.RUN_MATRIX_ONLY)
ctx.synchronize()
L = N.lib()
W = 8
n_tiles = 24
slots = 2 + W * n_tiles
L.bsg_debug_trace_enable.argtypes = [C.c_void_p, C.c_uint32]
L.bsg_debug_trace_read.argtypes = [C.c_void_p, C.c_void_p]
N.check(L.bsg_debug_trace_enable(ctx.handle, slots))
qs[0].run(N.PROBE_STAGED | N.RUN_MATRIX_ONLY)
ctx.synchronize()
sm = ctx.device_info()["sm_count"]
out = np.zeros((sm, slots), dtype=np.uint64)
N.check(L.bsg_debug_trace_read(ctx.handle, N.ptr(out)))
t0 = out[:, 0].min()
rel = (out.astype(np.int64) - int(t0)) / 1e3
rel[out == 0] = np.nan
np.set_printoptions(precision=2, suppress=True, linewidth=250)
print("workload", wl, "env", {k: v for k, v in os.environ.items() if k.startswith("BSG_")})
print("CTA start spread us %.2f..%.2f; hashes ready mean %.2f max %.2f" %
      (np.nanmin(rel[:, 0]), np.nanmax(rel[:, 0]), np.nanmean(rel[:, 1]), np.nanmax(rel[:, 1])))
names = ["resident", "A done  ", "B1 done ", "B2 done "]
for cta in (0, 73):
    for j, nm in enumerate(names):
        print("cta", cta, nm, rel[cta, 2 + j::W])
ends = np.nanmax(rel[:, 2:][:, [i for i in range(slots - 2) if i % W < 4]], axis=1)
print("end per CTA: min %.1f median %.1f max %.1f us" % (np.nanmin(ends), np.nanmedian(ends), np.nanmax(ends)))
for j, nm in enumerate(names):
    print("mean", nm, np.nanmean(rel[:, 2 + j::W], axis=0))
res, ad, b1, b2 = rel[:, 2::W], rel[:, 3::W], rel[:, 4::W], rel[:, 5::W]
n1 = out[:, 6::W].astype(float); n2 = out[:, 7::W].astype(float)
n1[np.isnan(res)] = np.nan; n2[np.isnan(res)] = np.nan
wait = res[:, 1:] - b2[:, :-1]
print("mean per tile: round A %.2f us, round B1 %.2f us, round B2 + barrier %.2f us, wait for the next tile %.2f us, period %.2f us; "
      "|L1| %.0f, |L2| %.0f" % (np.nanmean(ad - res), np.nanmean(b1 - ad), np.nanmean(b2 - b1), np.nanmean(wait),
                               np.nanmean(np.diff(res, axis=1)), np.nanmean(n1), np.nanmean(n2)))
