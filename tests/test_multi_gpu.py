"""-m gpu tests of the two real exchange steps of the path over NCCL, one
process per GPU: OR-combination of equal-(m,k) partial file-level bitsets built from
disjoint shards of a file's entries, and the all-gather of per-rank candidate masks of a
corpus sharded by file.  Needs >= 2 GPUs (skipped otherwise; the driver's 1-GPU box skips,
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu` runs them)."""
from __future__ import annotations

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid_q, res_q):
    import bloomsearch_b200 as bs
    from bloomsearch_b200 import _native as N
    from bloomsearch_b200.sharding import FileSharding, sharded_candidates, split_entries
    from oracle import cref
    from synth.corpus import SynthCorpus
    try:
        torch.cuda.set_device(rank)
        ctx = bs.Context(rank)
        if rank == 0:
            uid = bs.Context.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        ctx.comm_init(rank, world, uid)

        # ---- (1) file-level build: each rank inserts its shard of the union, then OR across ranks ----
        c = SynthCorpus(42, 0, 8, 400, 8)  # one file of 8 blocks
        union = sorted({c.key(i) for b in range(8) for i in range(int(c.group_begin[3 * b + 1]), int(c.group_begin[3 * b + 2]))})
        assert len(union) == int(c.file_counts[0][1])
        m, k = bs.estimate_parameters(len(union), 0.001)
        nw = (m + 63) // 64
        mine = union[split_entries(len(union), world, rank)]
        blob, off = N.pack_keys(mine)
        desc = np.array([(m, k, 0)], dtype=N.DESC_DTYPE)
        part = ctx.build(blob, off, np.array([0, len(mine)], np.uint64), np.zeros(1, np.uint32), None, desc, nw)
        full = ctx.or_reduce(part.copy())
        want = cref.Filter.build_sized(union, 0.001)
        ok_or = bool(np.array_equal(full, want.words())) and not np.array_equal(part, want.words())

        # ---- (2) probe sharded by file: local mask per rank, NCCL all-gather, assemble ----
        c2 = SynthCorpus(7, 0, 24, 300, 4)  # 6 files x 4 blocks
        counts = c2.group_counts().reshape(-1)
        d2 = np.zeros(len(counts), dtype=N.DESC_DTYPE)
        wo = 0
        for g, n in enumerate(counts):
            mm, kk = bs.estimate_parameters(max(int(n), 1), 0.001)
            d2[g] = (mm, kk, wo)
            wo += (mm + 63) // 64
        words = cref.build_filters(c2.blob, c2.key_off, c2.group_begin, np.arange(len(d2), dtype=np.uint32), None, d2, wo)
        sh = FileSharding([c2.blocks_per_file] * c2.n_files, world)
        units = sh.units_of(rank)
        local = bs.Corpus(ctx, d2.reshape(-1, 3)[units].reshape(-1), words)
        q = bs.BloomQuery(bs.Or(bs.Token(c2.key(int(c2.group_begin[3 * 5 + 1]) + 3)),
                                bs.And(bs.Field(b"nested.az"), bs.FieldToken(b"level", b"nope")),
                                bs.FieldToken(b"timestamp", b"%d" % (1700000000 + 300 * 17 + 5))))
        cq = bs.compile_bloom_query(q)
        _, lmask = local.probe(cq.keys, cq.kinds, cq.prog, want_matrix=False)
        got = sharded_candidates(sh, rank, lmask, lambda x: ctx.allgather_masks(x, world))
        b2, o2 = N.pack_keys(cq.keys)
        wmask = cref.probe_mask(d2, words, c2.n_blocks, b2, o2, cq.kinds, cq.prog)
        want_bits = bs.unpack_mask(wmask, c2.n_blocks)
        ok_probe = bool(np.array_equal(got, want_bits)) and bool(want_bits.any()) and not bool(want_bits.all())
        info = ctx.comm_info()
        assert info["world"] == world and info["rank"] == rank

        # ---- (3) the same two exchanges on DEVICE buffers: partial built straight into symmetric memory, one
        #      OR kernel over peer memory, twice (buffer reuse), plus an odd word count ----
        ks = bs.KeySet(ctx, blob, off, np.array([0, len(mine)], np.uint64))
        ks.set_filters(np.zeros(1, np.uint32), None, desc, nw)
        d_part = ctx.comm_alloc(nw * 8)
        ok_dev = True
        for _ in range(2):
            ks.build(d_part)
            ctx.or_reduce_device(d_part, nw)
            ok_dev &= bool(np.array_equal(ctx.read_device(d_part, nw), want.words()))
        assert info["peer_memory"] == ctx.comm_info()["peer_memory"]
        nvl = ctx.comm_info()["last_nvlink_bytes"]
        ok_dev &= nvl == 2 * (world - 1) * ((nw + world - 1) // world) * 8
        ks.close()
        # resident query -> device mask -> device all-gather
        dq = bs.Query(local, cq.keys, cq.kinds, cq.prog)
        dq.run(N.PROBE_AUTO)
        mw = sh.local_mask_words()
        d_all = ctx.comm_alloc(world * mw * 8)
        for _ in range(2):
            ctx.allgather_masks_device(dq.device_mask(), mw, d_all)
            gathered = ctx.read_device(d_all, world * mw).reshape(world, mw)
            ok_dev &= bool(np.array_equal(sh.assemble(gathered), want_bits))
        dq.close()

        # ---- (4) sharded hierarchical probe in one collective call (files dealt to ranks) ----
        fd = np.zeros(c2.n_files * 3, dtype=N.DESC_DTYPE)
        fsets = []
        for f in range(c2.n_files):
            for kind in range(3):
                ks_ = sorted({c2.key(i) for b in range(f * c2.blocks_per_file, (f + 1) * c2.blocks_per_file)
                              for i in range(int(c2.group_begin[3 * b + kind]), int(c2.group_begin[3 * b + kind + 1]))})
                fsets.append(cref.Filter.build_sized(ks_, 0.001))
        fwo = 0
        fchunks = []
        for i, flt in enumerate(fsets):
            fd[i] = (flt.m, flt.k, fwo)
            fchunks.append(flt.words())
            fwo += len(fchunks[-1])
        fwords = np.concatenate(fchunks)
        my_files = sh.files_of(rank)
        files_local = bs.Corpus(ctx, fd.reshape(-1, 3)[my_files].reshape(-1), fwords)
        parent = np.repeat(np.arange(len(my_files), dtype=np.uint32), c2.blocks_per_file)
        local.set_parents(parent, len(my_files))
        q2 = bs.BloomQuery(bs.And(bs.Or(bs.Token(c2.key(int(c2.group_begin[3 * 5 + 1]) + 3)), bs.FieldToken(b"level", b"nope")),
                                  bs.Field(b"nested.az")))
        allm = bs.probe_hierarchical_gather(files_local, local, q2, mw, world)
        got_h = sh.assemble(allm)
        cq2 = bs.compile_bloom_query(q2)
        b3, o3 = N.pack_keys(cq2.keys)
        fm = bs.unpack_mask(cref.probe_mask(fd, fwords, c2.n_files, b3, o3, cq2.kinds, cq2.prog), c2.n_files)
        bm = bs.unpack_mask(cref.probe_mask(d2, words, c2.n_blocks, b3, o3, cq2.kinds, cq2.prog), c2.n_blocks)
        want_h = bm & np.repeat(fm, c2.blocks_per_file)
        ok_dev &= bool(np.array_equal(got_h, want_h)) and bool(want_h.any())
        files_local.close()
        local.close()
        ctx.close()
        res_q.put((rank, ok_or, ok_probe and ok_dev, "" if ok_dev else "device collectives / hierarchical gather mismatch; peer_memory=%s" % info["peer_memory"]))
    except Exception as e:  # noqa: BLE001
        import traceback
        res_q.put((rank, False, False, traceback.format_exc()))


def test_or_reduce_and_mask_allgather_over_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctxmp = mp.get_context("spawn")
    uid_q, res_q = ctxmp.Queue(), ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, world, uid_q, res_q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(res_q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, ok_or, ok_probe, err in res:
        assert ok_or and ok_probe, f"rank {rank}: or={ok_or} probe={ok_probe}\n{err}"


def _timeout_worker(rank, world, uid_q, res_q, done_evt):
    """Rank 0 enters a peer-memory collective its peer never joins: the kernel must trap after
    BSG_COMM_TIMEOUT_S instead of spinning for ever, and the failure must surface as an exception."""
    import os
    import time
    os.environ["BSG_COMM_TIMEOUT_S"] = "2"
    import bloomsearch_b200 as bs
    try:
        torch.cuda.set_device(rank)
        ctx = bs.Context(rank)
        if rank == 0:
            uid = bs.Context.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        ctx.comm_init(rank, world, uid)
        peer_memory = ctx.comm_info()["peer_memory"]
        d = ctx.comm_alloc(4096 * 8)
        if rank == 0:
            if not peer_memory:   # the NCCL path has its own watchdog; nothing to test here
                res_q.put((0, "skip", 0.0))
            else:
                t0 = time.time()
                try:
                    ctx.or_reduce_device(d, 4096)
                    ctx.synchronize()
                    res_q.put((0, "no error raised", time.time() - t0))
                except Exception as e:  # noqa: BLE001
                    res_q.put((0, "raised: " + str(e)[:200], time.time() - t0))
        done_evt.wait(timeout=120)
    except Exception:  # noqa: BLE001
        import traceback
        res_q.put((rank, "setup failed: " + traceback.format_exc(), 0.0))
    os._exit(0)   # rank 0's context is dead after the trap and clean-up is collective: leave without it


def test_collective_traps_instead_of_hanging_on_a_missing_peer():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctxmp = mp.get_context("spawn")
    uid_q, res_q, done_evt = ctxmp.Queue(), ctxmp.Queue(), ctxmp.Event()
    procs = [ctxmp.Process(target=_timeout_worker, args=(r, world, uid_q, res_q, done_evt)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        rank, what, dt = res_q.get(timeout=180)
    finally:
        done_evt.set()
        for p in procs:
            p.join(timeout=60)
    if what == "skip":
        pytest.skip("communicator runs over NCCL (no peer memory)")
    assert what.startswith("raised: "), what
    assert 1.5 <= dt < 60, dt
