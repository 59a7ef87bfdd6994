#!/usr/bin/env python
"""Profiling helper: per-CTA timeline of the two-phase staged probe kernel (probe_staged2):
unit resident / phase A of warp 0 done / all A warps done / released, microseconds."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "2b"
ctx = bs.Context(0)
c = bench.gen_corpus(wl, 0)
desc, n_words = bench.size_filters(c, bs)
words = ctx.build(c.blob, c.key_off, c.group_begin, np.arange(len(desc), dtype=np.uint32), None, desc, n_words)
keys, kinds = bench.make_batch(c, 7)
corpora = [bs.Corpus(ctx, desc, words) for _ in range(4)]
qs = [bs.Query(cp, keys, kinds, None) for cp in corpora]
for i in range(8):
    qs[i % 4].run(N.PROBE_AUTO | N.RUN_MATRIX_ONLY)
ctx.synchronize()
L = N.lib()
W = 8
slots = W * 80 + 1 if wl == "2a" else W * 16 + 1
L.bsg_debug_trace_enable.argtypes = [C.c_void_p, C.c_uint32]
L.bsg_debug_trace_read.argtypes = [C.c_void_p, C.c_void_p]
N.check(L.bsg_debug_trace_enable(ctx.handle, slots))
qs[0].run(N.PROBE_AUTO | N.RUN_MATRIX_ONLY)
ctx.synchronize()
sm = ctx.device_info()["sm_count"]
out = np.zeros((sm, slots), dtype=np.uint64)
N.check(L.bsg_debug_trace_read(ctx.handle, N.ptr(out)))
t0 = out[:, 0].min()
rel = (out.astype(np.int64) - int(t0)) / 1e3
rel[out == 0] = np.nan
np.set_printoptions(precision=2, suppress=True, linewidth=220)
print("workload", wl, "variant", os.environ.get("BSG_PROBE_VARIANT", "1"), "start spread us", np.nanmin(rel[:, 0]), np.nanmax(rel[:, 0]))
n_it = min((slots - 1) // W, 10)
names = ["resident", "A0 done ", "A done  ", "released", "c0 hashes", "c0 tested", "B0 arrived"]
for cta in (0, 73):
    for j, nm in enumerate(names):
        print("cta", cta, nm, rel[cta, 1 + j:1 + j + W * n_it:W])
ends = np.nanmax(rel, axis=1)
print("end per CTA: min %.1f median %.1f max %.1f us" % (np.nanmin(ends), np.nanmedian(ends), np.nanmax(ends)))
for j, nm in enumerate(names):
    print("mean", nm, "per it:", np.nanmean(rel[:, 1 + j:1 + j + W * n_it:W], axis=0))
d_a = rel[:, 3::W] - rel[:, 1::W]
d_b = rel[:, 4::W] - rel[:, 3::W]
d_p = np.diff(rel[:, 1::W], axis=1)
print("chunk 0: A done -> hashes loaded %.2f us; hashes -> tested %.2f us; B warp 0: A done -> arrived %.2f us"
      % (np.nanmean(rel[:, 5::W] - rel[:, 3::W]), np.nanmean(rel[:, 6::W] - rel[:, 5::W]), np.nanmean(rel[:, 7::W] - rel[:, 3::W])))
print("mean A latency (resident -> all A done) %.2f us; B latency (A done -> released) %.2f us; period %.2f us"
      % (np.nanmean(d_a), np.nanmean(d_b), np.nanmean(d_p)))
