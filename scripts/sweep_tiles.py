#!/usr/bin/env python
"""GPU-box script: time the staged probe kernels over a list of env settings without regenerating the
corpus (one process; a fresh bsg context per setting since the knobs are read at bsg_create).

  python scripts/sweep_tiles.py 2b "BSG_PROBE_VARIANT=3" "BSG_TILES_SHAPE=0" "BSG_TILES_SHAPE=1 BSG_TILE_BYTES=16384" ...
  python scripts/sweep_tiles.py 600x15000 "BSG_PROBE_VARIANT=7" "SWEEP_PATH=gather"     # 600 blocks of 15 000 rows (~105 KB units)

A workload is a bench.py name (2a, 2b) or <blocks>x<rows per block>; SWEEP_PATH=gather in a setting times the gather
kernel instead of the staged path.

Prints per setting: us/launch on one stream (PDL as configured), on two streams, and whether the matrix
equals the first setting's (the first setting should be a trusted kernel)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bloomsearch_b200 as bs  # noqa: E402
from bloomsearch_b200 import _native as N  # noqa: E402


def main():
    wl = sys.argv[1]
    settings = sys.argv[2:] or ["BSG_PROBE_VARIANT=6"]
    steps = int(os.environ.get("SWEEP_STEPS", "200"))
    n_rep = int(os.environ.get("SWEEP_REPLICAS", "8"))
    scale = float(os.environ.get("SWEEP_SCALE", "1.0"))
    if "x" in wl:
        nb, rows = (int(x) for x in wl.split("x"))
        bench.WORKLOADS[wl] = (nb, rows, min(100, nb))
    c = bench.gen_corpus(wl, 0, scale)
    ctx0 = bs.Context(0)
    desc, n_words = bench.size_filters(c, bs)
    words = ctx0.build(c.blob, c.key_off, c.group_begin, np.arange(len(desc), dtype=np.uint32), None, desc, n_words)
    ctx0.close()
    keys, kinds = bench.make_batch(c, 7)
    n_units = c.n_blocks
    L = N.lib()
    L.bsg_debug_run_cycle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32]
    ref = None
    touched = set()
    for setting in settings:
        for k in touched:
            os.environ.pop(k, None)
        for kv in setting.split():
            k, v = kv.split("=")
            os.environ[k] = v
            touched.add(k)
        path = N.PROBE_GATHER if os.environ.get("SWEEP_PATH") == "gather" else N.PROBE_STAGED
        RUN = path | N.RUN_MATRIX_ONLY
        ctx = bs.Context(0)
        try:
            corpora = [bs.Corpus(ctx, desc, words) for _ in range(n_rep)]
            queries = [bs.Query(cp, keys, kinds, None) for cp in corpora]
            queries[0].run(path)
            got, _ = queries[0].fetch()
            if ref is None:
                ref = got
            same = bool(np.array_equal(got, ref))
            cp_arr = (C.c_void_p * n_rep)(*[cp.handle for cp in corpora])
            q_arr = (C.c_void_p * n_rep)(*[q._h for q in queries])

            def run(k, streams):
                N.check(L.bsg_debug_run_cycle(ctx.handle, cp_arr, q_arr, n_rep, k, RUN, streams))

            out = []
            for streams in (1, 2):
                run(20, streams)
                ctx.synchronize()
                best = 1e9
                for _ in range(3):
                    ctx.timer_begin()
                    run(steps, streams)
                    best = min(best, ctx.timer_end() / steps)
                out.append(best * 1e3)
            # end to end: bsg_probe with host buffers, single caller
            blob, off = N.pack_keys(keys)
            m_words = (len(keys) + 63) // 64
            outm = np.zeros((n_units, m_words), dtype=np.uint64)
            for i in range(10):
                corpora[i % n_rep].probe_packed(blob, off, kinds, None, outm, None)
            t0 = time.perf_counter()
            for i in range(100):
                corpora[i % n_rep].probe_packed(blob, off, kinds, None, outm, None)
            e2e_us = (time.perf_counter() - t0) / 100 * 1e6
            same_e2e = bool(np.array_equal(outm, ref))
            unit_kb = n_words * 8 / n_units / 1e3
            L.bsg_debug_probe_kernel_name.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
            name_buf = C.create_string_buffer(512)
            N.check(L.bsg_debug_probe_kernel_name(ctx.handle, corpora[0].handle, name_buf, 512))
            kname = "probe_gather_kernel" if path == N.PROBE_GATHER else name_buf.value.decode()
            gbs = n_words * 8 / (out[0] * 1e-6) / 1e9
            print(f"{wl} ({unit_kb:.0f} KB/unit; {kname}; {gbs:.0f} GB/s of bitsets on one stream) [{setting}] 1-stream {out[0]:.2f} us  2-stream {out[1]:.2f} us  bsg_probe {e2e_us:.1f} us/call  "
                  f"parity {'ok' if same and same_e2e else 'MISMATCH'}", flush=True)
            for q in queries:
                q.close()
            for cp in corpora:
                cp.close()
        finally:
            ctx.close()


if __name__ == "__main__":
    main()
