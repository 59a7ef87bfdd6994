#!/usr/bin/env python
"""(f.3) timing of bsg_count_distinct on the workload-2b emission stream with repeats; one JSON object.
CPU side: Python set() per group is NOT the reference (Go maps) — it is only a sanity check of the counts."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bloomsearch_b200 as bs
from synth.corpus import SynthCorpus

ctx = bs.Context(0)
c = SynthCorpus(42, 0, 1000, 10000, 100)          # 39 M distinct keys in 3000 groups, 10 files
n_groups = len(c.group_begin) - 1
# emit every key twice (the second copy right after the first group-wise): 78 M emissions
sizes = np.diff(c.group_begin).astype(np.int64)
lens = np.diff(c.key_off).astype(np.int64)
parent = np.array([(g // 3 // c.blocks_per_file) * 3 + g % 3 for g in range(n_groups)], dtype=np.uint32)
out = {}
for name, rep in (("unique_emissions", 1), ("each_key_twice", 2)):
    if rep == 1:
        blob, off, gb = c.blob, c.key_off, c.group_begin
    else:
        # duplicate each group's key range back to back
        pieces, offs, gb, pos = [], [0], [0], 0
        for g in range(n_groups):
            b, e = int(c.group_begin[g]), int(c.group_begin[g + 1])
            seg = c.blob[int(c.key_off[b]):int(c.key_off[e])]
            lo = (c.key_off[b:e + 1] - c.key_off[b]).astype(np.uint64)
            for _ in range(2):
                pieces.append(seg)
                offs.append(lo[1:] + np.uint64(pos))
                pos += len(seg)
            gb.append(gb[-1] + 2 * (e - b))
        blob = np.concatenate(pieces)
        off = np.concatenate([np.zeros(1, np.uint64)] + offs[1:])
        gb = np.array(gb, dtype=np.uint64)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        gc, pc = ctx.count_distinct(blob, off, gb, parent, 3 * c.n_files)
        best = min(best, time.perf_counter() - t)
    assert np.array_equal(gc.astype(np.int64), sizes), "group counts wrong"
    want_p = [int(c.file_counts[p // 3][p % 3]) for p in range(3 * c.n_files)]
    assert pc.tolist() == want_p, "parent counts wrong"
    n = len(off) - 1
    # resident form: the emissions already in HBM (bsg_keyset), counts only
    ks = bs.KeySet(ctx, blob, off, gb)
    best_res = 1e9
    for _ in range(4):
        t = time.perf_counter()
        gc2, pc2 = ks.count_distinct(parent, 3 * c.n_files)
        best_res = min(best_res, time.perf_counter() - t)
    assert np.array_equal(gc2, gc) and np.array_equal(pc2, pc)
    ks.close()
    out[name] = {"emissions": int(n), "groups": n_groups, "parents": 3 * c.n_files, "host_to_host_ms": best * 1e3,
                 "emissions_per_s": n / best, "resident_ms": best_res * 1e3, "resident_emissions_per_s": n / best_res,
                 "counts_verified": True}
print(json.dumps(out))
