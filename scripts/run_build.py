#!/usr/bin/env python
"""GPU-box helper for profiling the build kernel: config 3 at a reduced size (blocks of 10 000 tokens, with and
without the fused file-level filters), keys resident, a few bsg_keyset_build launches."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bloomsearch_b200 as bs  # noqa: E402
from bloomsearch_b200 import _native as N  # noqa: E402

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
with_file = (sys.argv[2] if len(sys.argv) > 2 else "file") == "file"
kpb, bpf = 10000, 100
ctx = bs.Context(0)
blob, key_off, group_begin = bench.gen_token_blocks(list(range(n_blocks)), 43, kpb)
mb, kb = bs.estimate_parameters(kpb, 0.001)
mf, kf = bs.estimate_parameters(kpb * bpf, 0.001)
wb, wf = (mb + 63) // 64, (mf + 63) // 64
nf = n_blocks // bpf
desc = np.zeros(n_blocks + nf, dtype=N.DESC_DTYPE)
desc["m"][:n_blocks], desc["k"][:n_blocks], desc["word_off"][:n_blocks] = mb, kb, np.arange(n_blocks, dtype=np.uint64) * wb
desc["m"][n_blocks:], desc["k"][n_blocks:], desc["word_off"][n_blocks:] = mf, kf, n_blocks * wb + np.arange(nf, dtype=np.uint64) * wf
gf = np.arange(n_blocks, dtype=np.uint32)
gf2 = (n_blocks + np.arange(n_blocks, dtype=np.uint32) // bpf).astype(np.uint32) if with_file else None
ks = bs.KeySet(ctx, blob, key_off, group_begin)
ks.set_filters(gf, gf2, desc, n_blocks * wb + nf * wf)
for _ in range(3):
    ks.build()
ctx.synchronize()
ctx.timer_begin()
for _ in range(5):
    ks.build()
ms = ctx.timer_end() / 5
print(f"build {n_blocks} blocks x {kpb} keys, file-level {with_file}: {ms:.3f} ms per launch, {n_blocks * kpb / ms / 1e3:.3e} keys/s")
