#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 3, 4 and the file-level stage); prints one JSON object.
Not the driver's bench: bench.py is.  Usage: python scripts/bench_extra.py [--blocks4 N]"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bloomsearch_b200 as bs
from bloomsearch_b200 import _native as N
from oracle import cref
from synth.corpus import SynthCorpus
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--blocks4", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
ctx = bs.Context(0)
L = N.lib()
L.bsg_debug_last_build_kernel_ms.argtypes = [C.c_void_p]
L.bsg_debug_last_build_kernel_ms.restype = C.c_float
out = {}
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def timed(fn, steps):
    for _ in range(3):
        fn()
    ctx.synchronize()
    ctx.timer_begin()
    for _ in range(steps):
        fn()
    return ctx.timer_end() / steps


# ---------------- config 3: filter build (block filters + fused file-level filters) ----------------
c = SynthCorpus(42, 0, 1000, 10000, 100)  # 39 M distinct keys, 10 files x 100 blocks
desc, n_words = bench.size_filters(c, bs)
n_block_filters = len(desc)
fdesc = []
wo = n_words
for f in range(c.n_files):
    for kind in range(3):
        m, k = bs.estimate_parameters(max(int(c.file_counts[f][kind]), 1), 0.001)
        fdesc.append((m, k, wo))
        wo += (m + 63) // 64
all_desc = np.concatenate([desc, np.array(fdesc, dtype=N.DESC_DTYPE)])
gf = np.arange(n_block_filters, dtype=np.uint32)
gf2 = np.array([n_block_filters + (b // c.blocks_per_file) * 3 + kind for b in range(c.n_blocks) for kind in range(3)], dtype=np.uint32)
res = {}
for name, g2, d, nw in (("blocks_only", None, desc, n_words), ("blocks_plus_file_level", gf2, all_desc, wo)):
    best_k, best_e = 1e9, 1e9
    for _ in range(3):
        t = time.perf_counter()
        words = ctx.build(c.blob, c.key_off, c.group_begin, gf, g2, d, nw)
        best_e = min(best_e, time.perf_counter() - t)
        best_k = min(best_k, L.bsg_debug_last_build_kernel_ms(ctx.handle))
    key_bytes = int(c.key_off[-1])
    algo = key_bytes + 8 * c.n_keys + 8 * nw
    res[name] = {"keys": int(c.n_keys), "kernel_ms": best_k, "keys_per_s_kernel": c.n_keys / (best_k / 1e3),
                 "algorithmic_GBps": algo / (best_k / 1e3) / 1e9, "frac_of_hbm": algo / (best_k / 1e3) / 1e9 / peak,
                 "e2e_ms_host_to_host": best_e * 1e3, "keys_per_s_e2e": c.n_keys / best_e,
                 "h2d_bytes": key_bytes + 8 * (c.n_keys + 1), "d2h_bytes": 8 * nw}
# parity spot check of the fused file-level filter against buildSizedBloomFilter(union) of file 0, kind token
union = sorted({c.key(i) for b in range(c.blocks_per_file) for i in range(int(c.group_begin[3 * b + 1]), int(c.group_begin[3 * b + 2]))})
ref = cref.Filter.build_sized(union, 0.001)
m, k, o = (int(x) for x in all_desc[n_block_filters + 1])
res["file_filter_parity"] = bool((ref.m, ref.k) == (m, k) and np.array_equal(words[o:o + ref.nwords], ref.words()))
t = time.perf_counter()
cref.build_filters(c.blob, c.key_off, c.group_begin[:301], gf[:300], None, desc[:300], int(desc[299]["word_off"]) + (int(desc[299]["m"]) + 63) // 64, n_threads=os.cpu_count())
dt = time.perf_counter() - t
res["cpu_port_keys_per_s"] = {"value": float(c.group_begin[300]) / dt, "cores": os.cpu_count(), "sample": "first 100 blocks, all host threads"}
out["config3_build"] = res
block_words = words[:n_words]

# ---------------- config 4: 8-key AND/OR query over many 1000-row blocks ----------------
c2 = SynthCorpus(42, 0, 10000, 1000, 100)
d2, nw2 = bench.size_filters(c2, bs)
w2 = ctx.build(c2.blob, c2.key_off, c2.group_begin, np.arange(len(d2), dtype=np.uint32), None, d2, nw2)
reps = max(1, args.blocks4 // c2.n_blocks)
big_desc = np.tile(d2, reps)  # descriptors alias the same host words; the device copy is reps x distinct HBM
corpus = bs.Corpus(ctx, big_desc, w2)
n_units = corpus.n_units
ft = lambda b, j: c2.key(int(c2.group_begin[3 * b + 2]) + j)
split = lambda key: key.split(b"::", 1)
present = [ft(17, 3), ft(4021, 1500), ft(9000, 2035), ft(77, 1999)]
keys8 = [present[0], b"level::nope", present[1], b"user_id::x1", b"service::auth", b"service::nosuch", present[2][:0] + b"level::info", b"nested.az::az-1"]
q = bs.BloomQuery(bs.And(bs.Or(*[bs.FieldToken(*split(k)) for k in keys8[:4]]), bs.Or(*[bs.FieldToken(*split(k)) for k in keys8[4:6]]),
                         bs.FieldToken(*split(keys8[6])), bs.FieldToken(*split(keys8[7]))))
cq = bs.compile_bloom_query(q)
r4 = {"units": int(n_units), "keys": len(cq.keys), "bitset_bytes_fieldtoken": int(corpus.bitset_bytes(4)), "bitset_bytes_all": int(corpus.bitset_bytes(7))}
dq = bs.Query(corpus, cq.keys, cq.kinds, cq.prog)
for name, path in (("gather", N.PROBE_GATHER), ("staged", N.PROBE_STAGED)):
    ms = timed(lambda: dq.run(path), args.steps)
    _, mask = dq.fetch(want_matrix=False)
    r4[name] = {"ms": ms, "probes_per_s": n_units * len(cq.keys) / (ms / 1e3), "survivors": int(bs.unpack_mask(mask, n_units).sum()),
                "launches": dq.launches()}
    if name == "staged":
        r4[name]["GBps_vs_fieldtoken_bytes"] = corpus.bitset_bytes(4) / (ms / 1e3) / 1e9
# parity on the first 10k units against the oracle
b8, o8 = N.pack_keys(cq.keys)
wm = cref.probe_mask(d2, w2, c2.n_blocks, b8, o8, cq.kinds, cq.prog, n_threads=os.cpu_count())
r4["parity_first_10k_units"] = bool(np.array_equal(bs.unpack_mask(mask, n_units)[:c2.n_blocks], bs.unpack_mask(wm, c2.n_blocks)))
corpus.probe(cq.keys, cq.kinds, cq.prog, want_matrix=False)  # warm-up: scratch + pinned staging allocation
t = time.perf_counter()
for _ in range(5):
    corpus.probe(cq.keys, cq.kinds, cq.prog, want_matrix=False)
r4["e2e_ms_bsg_probe_mask_only"] = (time.perf_counter() - t) * 1e3 / 5
dq.close()
out["config4_and_or_8keys"] = r4

# ---------------- file-level stage: 10k files, file filters of 100 x 1000-row blocks ----------------
fd = []
wo = 0
for f in range(c2.n_files):
    for kind in range(3):
        m, k = bs.estimate_parameters(max(int(c2.file_counts[f][kind]), 1), 0.001)
        fd.append((m, k, wo)); wo += (m + 63) // 64
fd = np.array(fd, dtype=N.DESC_DTYPE)
alld = np.concatenate([d2, fd + np.array([(0, 0, nw2)], dtype=N.DESC_DTYPE)]) if False else None
fd_abs = fd.copy(); fd_abs["word_off"] += nw2
gf2b = np.array([len(d2) + (b // c2.blocks_per_file) * 3 + kind for b in range(c2.n_blocks) for kind in range(3)], dtype=np.uint32)
wall = ctx.build(c2.blob, c2.key_off, c2.group_begin, np.arange(len(d2), dtype=np.uint32), gf2b, np.concatenate([d2, fd_abs]), nw2 + wo)
file_words = wall[nw2:]
freps = max(1, 10000 // c2.n_files)
fcorpus = bs.Corpus(ctx, np.tile(fd, freps), file_words)
fq = bs.Query(fcorpus, cq.keys, cq.kinds, cq.prog)
ms = timed(lambda: fq.run(N.PROBE_AUTO), args.steps)
_, fmask = fq.fetch(want_matrix=False)
b8, o8 = N.pack_keys(cq.keys)
wmf = cref.probe_mask(fd, file_words, c2.n_files, b8, o8, cq.kinds, cq.prog)
out["file_level_stage"] = {"files": int(fcorpus.n_units), "filter_bytes": int(fcorpus.bitset_bytes(7)), "ms": ms,
                           "probes_per_s": fcorpus.n_units * len(cq.keys) / (ms / 1e3),
                           "parity": bool(np.array_equal(bs.unpack_mask(fmask, fcorpus.n_units)[:c2.n_files], bs.unpack_mask(wmf, c2.n_files)))}
# ---------------- hierarchical: file stage -> compaction -> block stage (config 4 as the engine runs it) ----
if fcorpus.n_units * c2.blocks_per_file == corpus.n_units:
    parent = (np.arange(corpus.n_units, dtype=np.int64) // c2.n_blocks * c2.n_files +
              (np.arange(corpus.n_units, dtype=np.int64) % c2.n_blocks) // c2.blocks_per_file).astype(np.uint32)
    corpus.set_parents(parent, fcorpus.n_units)
    bs.probe_hierarchical(fcorpus, corpus, q)  # warm-up (scratch allocation)
    best = 1e9
    for _ in range(5):
        t = time.perf_counter()
        fm, bm = bs.probe_hierarchical(fcorpus, corpus, q)
        best = min(best, time.perf_counter() - t)
    flat = bs.unpack_mask(mask, n_units)
    out["hierarchical_8keys"] = {"files": int(fcorpus.n_units), "blocks": int(corpus.n_units), "files_surviving": int(fm.sum()),
                                 "blocks_surviving": int(bm.sum()), "e2e_ms_host_to_host": best * 1e3,
                                 "consistent_with_flat_probe": bool(np.array_equal(bm, flat & fm[parent]))}
corpus.close()
print(json.dumps(out, indent=1))
