"""world_size-2 gloo tests (CPU) of the N>1 host logic: file sharding, candidate-mask gather,
and the OR-combination identity behind the partial file-level build.  The
per-rank probe/build is played by the oracle here; on GPUs the same code runs over NCCL."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bloomsearch_b200.sharding import FileSharding, sharded_candidates, split_entries
from oracle import cref
from synth.corpus import SynthCorpus


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _units(c, fpr=0.001):
    counts = c.group_counts().reshape(-1)
    desc = np.zeros(len(counts), dtype=cref.DESC_DTYPE)
    wo = 0
    for g, n in enumerate(counts):
        m, k = cref.estimate_parameters(max(int(n), 1), fpr)
        desc[g] = (m, k, wo)
        wo += (m + 63) // 64
    words = cref.build_filters(c.blob, c.key_off, c.group_begin, np.arange(len(desc), dtype=np.uint32), None, desc, wo)
    return desc, words


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- probe: shard by file, local masks, all-gather, assemble ----
        c = SynthCorpus(42, 0, 12, 200, 2)              # 6 files x 2 blocks
        desc, words = _units(c)
        sh = FileSharding([c.blocks_per_file] * c.n_files, world)
        mine = sh.units_of(rank)
        keys = [c.key(int(c.group_begin[3 * 3 + 1]) + 5), c.key(int(c.group_begin[3 * 8 + 2]) + 9), b"absent"]
        kinds = np.array([1, 2, 1], dtype=np.uint8)
        blob, off = cref.pack_keys(keys)
        prog = np.array([(0, 0), (0, 1), (0, 2), (2, 3)], dtype=cref.OP_DTYPE)  # OR of the three
        local_desc = desc.reshape(-1, 3)[mine].reshape(-1)
        local = cref.probe_mask(local_desc, words, len(mine), blob, off, kinds, prog)

        def all_gather(x):
            t = torch.from_numpy(x.astype(np.int64))
            outs = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            return np.stack([o.numpy().astype(np.uint64) for o in outs])

        got = sharded_candidates(sh, rank, local, all_gather)
        want_words = cref.probe_mask(desc, words, c.n_blocks, blob, off, kinds, prog)
        want = np.unpackbits(want_words.view(np.uint8), bitorder="little")[:c.n_blocks].astype(bool)
        ok_probe = bool(np.array_equal(got, want)) and bool(want.any()) and not bool(want.all())

        # ---- file-level build: partial bitsets of identical (m,k) from disjoint entry shards, OR ----
        entries = sorted({c.key(i) for i in range(int(c.group_begin[1]), int(c.group_begin[2]))} |
                         {c.key(i) for i in range(int(c.group_begin[4]), int(c.group_begin[5]))})
        full = cref.Filter.build_sized(entries, 0.001)
        part = cref.Filter.new(full.m, full.k)
        for e in entries[split_entries(len(entries), world, rank)]:
            part.add(e)
        t = torch.from_numpy(part.words().astype(np.int64))
        outs = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        combined = np.bitwise_or.reduce(np.stack([o.numpy().astype(np.uint64) for o in outs]), axis=0)
        ok_or = bool(np.array_equal(combined, full.words())) and not np.array_equal(part.words(), full.words())
        q.put((rank, ok_probe, ok_or))
    finally:
        dist.destroy_process_group()


def test_sharded_probe_and_or_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]


def test_file_sharding_is_balanced_and_total():
    sh = FileSharding([100, 1, 50, 50, 100, 3, 96], 4)
    seen = np.concatenate([sh.units_of(r) for r in range(4)])
    assert sorted(seen.tolist()) == list(range(400))
    loads = [len(sh.units_of(r)) for r in range(4)]
    assert max(loads) - min(loads) <= 4
    g = np.zeros((4, sh.local_mask_words()), dtype=np.uint64)
    for r in range(4):
        n = len(sh.units_of(r))
        bits = np.zeros(sh.local_mask_words() * 64, dtype=np.uint8)
        bits[:n] = (sh.units_of(r) % 3 == 0)
        g[r] = np.packbits(bits, bitorder="little").view(np.uint64)
    out = sh.assemble(g)
    assert np.array_equal(out, np.arange(400) % 3 == 0)
