// (f.3) Exact distinct counts on the device — what the three Go maps of bloomEntrySets exist for
// (ingest.go:24-45 dedup, :105-123 unionInto / counts): the number of distinct entries of every group
// (block x kind) and of every parent (file x kind) union, which is the `n` that sizes the filters
// (ingest.go:139-140).  Emissions may repeat.
//
// A Go map is a hash set, and so is this: one open-addressing table in HBM for all groups at once, one
// thread per emission, no sort and no library call.
//
//   slot      one 64-bit word: 0 = empty, else (segment id << 32) | (emission index + 1).  Claimed with ONE
//             atomicCAS, so a slot is never seen half written and no thread ever waits for another.
//   home      MurmurHash3_x64_128 of the key bytes (h0, the hash the build kernel needs anyway) mixed with the
//             segment id; linear probing; the table holds at least twice as many slots as emissions.
//   equality  an occupied slot of the same segment is compared BYTE FOR BYTE with the emission (SURVEY §8f.3:
//             equal hashes are never taken as equal keys): equal -> the emission is a repeat; different ->
//             keep probing.  The count is therefore exact for any hash quality (BSG_DISTINCT_HASH_BITS=<n>
//             keeps only n bits of h0 so that tests can force thousands of collisions).
//   counts    an emission that claims a slot is the first of its key in its group: one atomicAdd per (warp,
//             group) on the group's counter, and — file-level unions (flush.go:221,253) — the same insertion
//             into a second table keyed by the group's parent, counted per parent.
// Round 2 first ordered the emissions with cub::DeviceRadixSort (35 ms for 39 M emissions); the table needs
// one pass over the keys and about two random 32-byte accesses per emission.
#include <cstdlib>

#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

namespace {

// MurmurHash3_x64_128(key, seed 0): the first half of bloom/v3's baseHashes
__device__ __forceinline__ void murmur128(const uint8_t* key, uint32_t len, uint64_t& o1, uint64_t& o2) {
    uint64_t h1 = 0, h2 = 0, k1 = 0, k2 = 0;
    const uint32_t nblocks = len >> 4, t = len & 15;
    if (len != 0) {
        WordReader rd(key);
        for (uint32_t b = 0; b < nblocks; ++b) {
            const uint64_t a = rd.next();
            const uint64_t c = rd.next();
            bmix(h1, h2, a, c);
        }
        if (t > 0) k1 = rd.next() & low_bytes_mask(t);
        if (t > 8) k2 = rd.next() & low_bytes_mask(t - 8);
    }
    finalize(h1, h2, k1, k2, len, o1, o2);
}

__device__ __forceinline__ bool same_key_bytes(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off,
                                               uint32_t i, uint32_t p) {
    const uint64_t bi = __ldg(&key_off[i]), ei = __ldg(&key_off[i + 1]);
    const uint64_t bp = __ldg(&key_off[p]), ep = __ldg(&key_off[p + 1]);
    if (ei - bi != ep - bp) return false;
    uint32_t left = static_cast<uint32_t>(ei - bi);
    if (left == 0) return true;
    WordReader ri(keys + bi), rp(keys + bp);
    while (left >= 8) {
        if (ri.next() != rp.next()) return false;
        left -= 8;
    }
    if (left) {
        const uint64_t m = low_bytes_mask(left);
        if ((ri.next() & m) != (rp.next() & m)) return false;
    }
    return true;
}

__device__ __forceinline__ uint64_t ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// true: emission i is the first of its key in segment `seg` (it now owns a slot); false: a repeat
__device__ __forceinline__ bool set_insert(unsigned long long* __restrict__ slots, uint64_t slot_mask, uint64_t h0, uint32_t seg,
                                           uint32_t i, const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off) {
    uint64_t x = h0 ^ (static_cast<uint64_t>(seg) * 0x9E3779B97F4A7C15ull);
    x ^= x >> 29;
    uint64_t pos = x & slot_mask;
    const unsigned long long mine = (static_cast<unsigned long long>(seg) << 32) | (static_cast<unsigned long long>(i) + 1ull);
    for (;;) {
        unsigned long long v = ld_relaxed_u64(&slots[pos]);
        if (v == 0ull) {
            v = atomicCAS(&slots[pos], 0ull, mine);
            if (v == 0ull) return true;
        }
        if (static_cast<uint32_t>(v >> 32) == seg &&
            same_key_bytes(keys, key_off, i, static_cast<uint32_t>(v) - 1u))
            return false;
        pos = (pos + 1) & slot_mask;
    }
}

// counts[seg] += 1 for every lane with `add`, one atomic per distinct segment in the warp
__device__ __forceinline__ void warp_count(unsigned long long* __restrict__ counts, uint32_t seg, bool add) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t key = add ? seg : 0xffffffffu;
    const uint32_t same = __match_any_sync(0xffffffffu, key);
    if (add && lane == static_cast<uint32_t>(__ffs(static_cast<int>(same)) - 1))
        atomicAdd(&counts[seg], static_cast<unsigned long long>(__popc(same)));
}

__global__ void __launch_bounds__(256)
distinct_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off, uint64_t n_keys,
                const uint64_t* __restrict__ group_begin, uint32_t n_groups, const uint32_t* __restrict__ group_parent,
                uint64_t h0_mask, unsigned long long* __restrict__ tab_g, unsigned long long* __restrict__ tab_p,
                uint64_t slot_mask, unsigned long long* __restrict__ group_counts,
                unsigned long long* __restrict__ parent_counts) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool first_in_group = false, first_in_parent = false;
    uint32_t g = 0, p = 0;
    // the group of the warp's first emission by binary search (one lane), then every lane walks forward from it:
    // emissions are grouped (CSR), so 32 consecutive ones rarely span more than two groups
    {
        const uint64_t i0 = i - (threadIdx.x & 31);
        uint32_t lo = 0;
        if ((threadIdx.x & 31) == 0 && i0 < n_keys) {
            uint32_t hi = n_groups;  // last g with group_begin[g] <= i0
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&group_begin[mid]) <= i0) lo = mid; else hi = mid;
            }
        }
        g = __shfl_sync(0xffffffffu, lo, 0);
    }
    if (i < n_keys) {
        while (g + 1 < n_groups && __ldg(&group_begin[g + 1]) <= i) ++g;   // also skips empty groups
        const uint64_t b = __ldg(&key_off[i]), e = __ldg(&key_off[i + 1]);
        uint64_t h0, h1;
        murmur128(keys + b, static_cast<uint32_t>(e - b), h0, h1);
        h0 &= h0_mask;
        first_in_group = set_insert(tab_g, slot_mask, h0, g, static_cast<uint32_t>(i), keys, key_off);
        if (first_in_group && tab_p) {
            p = __ldg(&group_parent[g]);
            // a different mix constant for the second table is not needed: it is a different table
            first_in_parent = set_insert(tab_p, slot_mask, h0, p, static_cast<uint32_t>(i), keys, key_off);
        }
    }
    warp_count(group_counts, g, first_in_group);
    if (tab_p) warp_count(parent_counts, p, first_in_parent);
}

uint64_t table_slots(uint64_t n_keys) {   // power of two, >= 2 n (load factor <= 1/2)
    uint64_t t = 1024;
    while (t < 2 * n_keys) t <<= 1;
    return t;
}

}  // namespace

cudaError_t launch_count_distinct(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                                  const uint64_t* d_group_begin, uint32_t n_groups, const uint32_t* d_group_parent,
                                  uint32_t n_parents, void* d_scratch, unsigned long long* d_group_counts,
                                  unsigned long long* d_parent_counts, cudaStream_t s) {
    (void)n_parents;
    if (n_keys == 0 || n_groups == 0) return cudaSuccess;
    if (n_keys > 0xfffffff0ull) return cudaErrorInvalidValue;   // emission indexes are 32-bit
    uint64_t h0_mask = ~0ull;
    if (const char* w = getenv("BSG_DISTINCT_HASH_BITS")) {
        const int b = atoi(w);
        if (b > 0 && b < 64) h0_mask = (1ull << b) - 1ull;
    }
    const uint64_t T = table_slots(n_keys);
    const bool unions = d_group_parent && d_parent_counts;
    unsigned long long* tab_g = static_cast<unsigned long long*>(d_scratch);
    unsigned long long* tab_p = unions ? tab_g + T : nullptr;
    cudaError_t e = cudaMemsetAsync(tab_g, 0, T * 8 * (unions ? 2 : 1), s);
    if (e != cudaSuccess) return e;
    const uint32_t blocks = static_cast<uint32_t>((n_keys + 255) / 256);
    distinct_kernel<<<blocks, 256, 0, s>>>(d_keys, d_key_off, n_keys, d_group_begin, n_groups, d_group_parent, h0_mask, tab_g,
                                           tab_p, T - 1, d_group_counts, d_parent_counts);
    return cudaGetLastError();
}

size_t count_distinct_scratch_bytes(uint64_t n_keys) {
    return static_cast<size_t>(table_slots(std::max<uint64_t>(n_keys, 1))) * 8 * 2 + 256;
}

}  // namespace bsg
