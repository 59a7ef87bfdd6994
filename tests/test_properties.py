"""Property tests (hypothesis) of the host logic either side of the hot path and of the oracle's
size-independent invariants: file sharding, the partial-build / OR identity behind config 5
(flush.go:221,253), no false negatives, and the section codec round trip (file_format.go:343-448)."""
from __future__ import annotations

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from bloomsearch_b200.sharding import FileSharding, sharded_candidates, split_entries
from oracle import bloomref as py
from oracle import cref

FAST = settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@FAST
@given(st.lists(st.integers(0, 300), min_size=0, max_size=40), st.integers(1, 8))
def test_file_sharding_partitions_and_balances(blocks, world):
    sh = FileSharding(blocks, world)
    owned = [sh.files_of(r) for r in range(world)]
    assert sorted(f for fs in owned for f in fs) == list(range(len(blocks)))            # every file exactly once
    units = [sh.units_of(r) for r in range(world)]
    allu = np.concatenate(units) if units else np.zeros(0, np.int64)
    assert sorted(allu.tolist()) == list(range(sum(blocks)))                             # every block exactly once
    loads = [int(sum(blocks[f] for f in fs)) for fs in owned]
    if blocks:
        assert max(loads) - min(loads) <= max(blocks)                                    # LPT greedy bound
    # a file's blocks stay together and in order on their rank (the probe needs no exchange)
    for r in range(world):
        pos = 0
        for f in owned[r]:
            n = blocks[f]
            assert units[r][pos:pos + n].tolist() == list(range(int(sh.file_first_unit[f]), int(sh.file_first_unit[f]) + n))
            pos += n


@FAST
@given(st.lists(st.integers(0, 70), min_size=1, max_size=12), st.integers(1, 5), st.integers(0, 2 ** 32 - 1))
def test_mask_gather_round_trip(blocks, world, seed):
    """Packing every rank's local candidate bits, all-gathering the padded words and assembling them gives
    back the global candidate set (what bsg_allgather_masks + FileSharding.assemble do over NCCL)."""
    sh = FileSharding(blocks, world)
    rng = np.random.default_rng(seed)
    truth = rng.random(sh.n_units) < 0.4
    local = []
    for r in range(world):
        bits = truth[sh.units_of(r)]
        words = np.packbits(np.concatenate([bits, np.zeros((-len(bits)) % 64, bool)]), bitorder="little").view(np.uint64)
        local.append(words)
    gathered = np.stack([sh.pad_local_mask(w) for w in local])
    for r in range(world):
        got = sharded_candidates(sh, r, local[r], lambda mine: gathered)
        assert np.array_equal(got, truth)


@FAST
@given(st.integers(0, 1000), st.integers(1, 8))
def test_split_entries_is_a_partition(n, world):
    seen = []
    for r in range(world):
        s = split_entries(n, world, r)
        seen += list(range(n))[s]
    assert seen == list(range(n))


keys_st = st.lists(st.binary(min_size=0, max_size=40), min_size=1, max_size=60, unique=True)


@FAST
@given(keys_st, st.integers(1, 4), st.sampled_from([0.5, 0.01, 0.001]))
def test_or_of_partial_filters_is_the_filter_of_the_union(keys, world, fpr):
    """config 5's identity: filters of identical (m, k) built from disjoint shards of an entry set OR to the
    filter built from the whole set — the library's Add is an OR of k bits (ingest.go:141-143)."""
    m, k = cref.estimate_parameters(max(len(keys), 1), fpr)
    whole = cref.Filter.new(m, k)
    for key in keys:
        whole.add(key)
    acc = np.zeros_like(whole.words())
    for r in range(world):
        part = cref.Filter.new(m, k)
        for key in keys[split_entries(len(keys), world, r)]:
            part.add(key)
        acc |= part.words()
    assert np.array_equal(acc, whole.words())
    assert all(whole.test(key) for key in keys)                                          # no false negatives
    twin = py.BloomFilter(m, k)                                                               # independent restatement
    for key in keys:
        twin.add(key)
    assert [int(w) for w in twin.words()] == [int(w) for w in whole.words()]


@FAST
@given(st.lists(st.binary(min_size=1, max_size=20), min_size=0, max_size=30, unique=True),
       st.lists(st.binary(min_size=1, max_size=20), min_size=0, max_size=30, unique=True),
       st.booleans(), st.integers(0, 2 ** 31))
def test_section_codec_round_trip_and_single_byte_corruption(fields, tokens, with_ft, flip):
    """encodeFilterSection -> parseFilterSection returns the same filters; any single flipped byte is caught by
    the CRC32C or the framing (file_format.go:343-385,392-448)."""
    f_field = cref.Filter.build_sized(fields, 0.01) if fields else None
    f_token = cref.Filter.build_sized(tokens, 0.01) if tokens else None
    f_ft = cref.Filter.build_sized([a + b"::" + b for a in fields[:3] for b in tokens[:3]] or [b"x::y"], 0.01) if with_ft else None
    sec = cref.section_encode(f_field, f_token, f_ft)
    back = cref.section_parse(sec)
    for a, b in zip((f_field, f_token, f_ft), back):
        assert (a is None) == (b is None)
        if a is not None:
            assert (a.m, a.k) == (b.m, b.k) and np.array_equal(a.words(), b.words())
    bad = bytearray(sec)
    bad[flip % len(bad)] ^= 0x5A
    try:
        cref.section_parse(bytes(bad))
        caught = False
    except Exception:  # noqa: BLE001
        caught = True
    assert caught
