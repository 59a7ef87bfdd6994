// (f.3) Exact distinct counts on the device — what the three Go maps of bloomEntrySets exist for
// (ingest.go:24-45 dedup, :105-123 unionInto / counts): the number of distinct entries of every group
// (block x kind) and of every parent (file x kind) union, which is the `n` that sizes the filters
// (ingest.go:139-140).  Emissions may repeat.
//
//   1. emit    one thread per emission: MurmurHash3_x64_128 of the key bytes -> (h0, h1), its group id
//   2. sort    LSD radix sort (cub::DeviceRadixSort, library code like a plain cuBLAS GEMM) of
//              (h0 -> emission index), then a stable sort of those indexes by group id: order (group, h0)
//   3. count   an emission is a run head when its (group, h0) differs from its left neighbour's, OR when the
//              KEY BYTES differ (the tie check of SURVEY §8f.3: equal 64-bit hashes are not taken as equal
//              keys).  One atomic per (warp, group run).
//   4. unions  the h0-sorted order is re-sorted by parent id and counted the same way.
// A byte mismatch inside an equal-hash run means a real 64-bit collision between two distinct keys of one
// group; their repeats could interleave, so the pass is repeated for that call with h1 as a third sort key
// (order (group, h0, h1): equal keys are adjacent again).  BSG_DISTINCT_HASH_BITS=<n> keeps only n bits
// of h0 to force that path in tests.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

namespace {

struct Scratch {
    uint64_t *h0, *h1, *ka, *kb;
    uint32_t *seg, *va, *vb, *sa, *sb;
    unsigned long long* collisions;
    void* temp;
    size_t temp_bytes;
};

size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

size_t temp_bytes_for(uint64_t n) {
    size_t t64 = 0, t32 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t64, static_cast<const uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), n);
    cub::DeviceRadixSort::SortPairs(nullptr, t32, static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr),
                                    static_cast<const uint32_t*>(nullptr), static_cast<uint32_t*>(nullptr), n);
    return std::max(t64, t32);
}

Scratch carve(void* base, uint64_t n) {
    Scratch s;
    uint8_t* p = static_cast<uint8_t*>(base);
    auto take = [&](size_t bytes) { uint8_t* r = p; p += align_up(bytes); return r; };
    s.h0 = reinterpret_cast<uint64_t*>(take(n * 8));
    s.h1 = reinterpret_cast<uint64_t*>(take(n * 8));
    s.ka = reinterpret_cast<uint64_t*>(take(n * 8));
    s.kb = reinterpret_cast<uint64_t*>(take(n * 8));
    s.seg = reinterpret_cast<uint32_t*>(take(n * 4));
    s.va = reinterpret_cast<uint32_t*>(take(n * 4));
    s.vb = reinterpret_cast<uint32_t*>(take(n * 4));
    s.sa = reinterpret_cast<uint32_t*>(take(n * 4));
    s.sb = reinterpret_cast<uint32_t*>(take(n * 4));
    s.collisions = reinterpret_cast<unsigned long long*>(take(8));
    s.temp_bytes = temp_bytes_for(n);
    s.temp = take(s.temp_bytes);
    return s;
}

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

// hash every key and tag it with its group (binary search of the CSR group_begin)
__global__ void __launch_bounds__(256)
emit_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off, uint64_t n_keys,
            const uint64_t* __restrict__ group_begin, uint32_t n_groups, uint64_t h0_mask, uint64_t* __restrict__ h0,
            uint64_t* __restrict__ h1, uint64_t* __restrict__ sort_key, uint32_t* __restrict__ idx,
            uint32_t* __restrict__ seg) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_keys) return;
    const uint64_t b = __ldg(&key_off[i]), e = __ldg(&key_off[i + 1]);
    uint64_t a, c;
    murmur128(keys + b, static_cast<uint32_t>(e - b), a, c);
    a &= h0_mask;
    uint32_t lo = 0, hi = n_groups;  // last g with group_begin[g] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&group_begin[mid]) <= i) lo = mid; else hi = mid;
    }
    h0[i] = a;
    h1[i] = c;
    sort_key[i] = a;
    idx[i] = static_cast<uint32_t>(i);
    seg[i] = lo;
}

// out[j] = table[in_idx[j]] (optionally mapped through `parent`)
__global__ void __launch_bounds__(256)
gather_seg_kernel(const uint32_t* __restrict__ order, uint64_t n, const uint32_t* __restrict__ seg,
                  const uint32_t* __restrict__ parent, uint32_t* __restrict__ out) {
    const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t g = seg[order[j]];
    out[j] = parent ? __ldg(&parent[g]) : g;
}
__global__ void __launch_bounds__(256)
gather_u64_kernel(const uint32_t* __restrict__ order, uint64_t n, const uint64_t* __restrict__ table, uint64_t* __restrict__ out) {
    const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j < n) out[j] = table[order[j]];
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

// counts[seg] += run heads of the array sorted by (seg, h0[, h1]); one atomic per (warp, segment run)
__global__ void __launch_bounds__(256)
count_heads_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ segs, uint64_t n,
                   const uint64_t* __restrict__ h0, const uint64_t* __restrict__ h1, int use_h1,
                   const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off,
                   unsigned long long* __restrict__ counts, unsigned long long* __restrict__ collisions) {
    const uint64_t j = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    bool head = false;
    uint32_t seg = 0xffffffffu;
    if (j < n) {
        seg = segs[j];
        head = true;
        if (j > 0 && segs[j - 1] == seg) {
            const uint32_t i = order[j], p = order[j - 1];
            if (h0[i] == h0[p] && (!use_h1 || h1[i] == h1[p])) {
                // tie check: equal hashes are not taken as equal keys
                if (same_key_bytes(keys, key_off, i, p)) head = false;
                else atomicAdd(collisions, 1ull);
            }
        }
    }
    // lanes of one segment are contiguous (sorted): the first lane of each run adds the run's head count
    const uint32_t same = __match_any_sync(0xffffffffu, seg);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (j < n && lane == static_cast<uint32_t>(__ffs(same) - 1)) {
        const uint32_t c = __popc(heads & same);
        if (c) atomicAdd(&counts[seg], static_cast<unsigned long long>(c));
    }
}

int bits_for(uint32_t n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) ++b;
    return b;
}

#define CUB_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return _e; } while (0)

// order_h (indexes sorted by h0 [then h1]) -> stable sort by segment id -> count
cudaError_t count_by_segment(const Scratch& S, uint64_t n, const uint32_t* order_h, const uint32_t* d_parent, uint32_t n_segments,
                             int use_h1, const uint8_t* d_keys, const uint64_t* d_key_off, unsigned long long* d_counts,
                             uint32_t* order_out, cudaStream_t s) {
    const uint32_t blocks = static_cast<uint32_t>((n + 255) / 256);
    gather_seg_kernel<<<blocks, 256, 0, s>>>(order_h, n, S.seg, d_parent, S.sa);
    CUB_TRY(cudaGetLastError());
    size_t tb = S.temp_bytes;
    CUB_TRY(cub::DeviceRadixSort::SortPairs(S.temp, tb, S.sa, S.sb, order_h, order_out, n, 0, bits_for(n_segments), s));
    count_heads_kernel<<<blocks, 256, 0, s>>>(order_out, S.sb, n, S.h0, S.h1, use_h1, d_keys, d_key_off, d_counts, S.collisions);
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_count_distinct(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                                  const uint64_t* d_group_begin, uint32_t n_groups, const uint32_t* d_group_parent,
                                  uint32_t n_parents, void* d_scratch, unsigned long long* d_group_counts,
                                  unsigned long long* d_parent_counts, cudaStream_t s) {
    if (n_keys == 0 || n_groups == 0) return cudaSuccess;
    if (n_keys > 0xfffffff0ull) return cudaErrorInvalidValue;   // emission indexes are 32-bit
    const Scratch S = carve(d_scratch, n_keys);
    uint64_t h0_mask = ~0ull;
    if (const char* w = getenv("BSG_DISTINCT_HASH_BITS")) {
        const int b = atoi(w);
        if (b > 0 && b < 64) h0_mask = (1ull << b) - 1ull;
    }
    const uint32_t blocks = static_cast<uint32_t>((n_keys + 255) / 256);
    CUB_TRY(cudaMemsetAsync(S.collisions, 0, 8, s));
    emit_kernel<<<blocks, 256, 0, s>>>(d_keys, d_key_off, n_keys, d_group_begin, n_groups, h0_mask, S.h0, S.h1, S.ka, S.va, S.seg);
    CUB_TRY(cudaGetLastError());
    size_t tb = S.temp_bytes;
    CUB_TRY(cub::DeviceRadixSort::SortPairs(S.temp, tb, S.ka, S.kb, S.va, S.vb, n_keys, 0, 64, s));   // S.vb: order by h0
    for (int use_h1 = 0; use_h1 < 2; ++use_h1) {
        if (use_h1) {
            // a real 64-bit collision between distinct keys of one group: add h1 as a sort key so that repeats of
            // equal keys are adjacent again.  LSD: h1 first, then h0 (both stable), then the segment ids.
            CUB_TRY(cudaMemsetAsync(d_group_counts, 0, static_cast<size_t>(n_groups) * 8, s));
            if (d_parent_counts) CUB_TRY(cudaMemsetAsync(d_parent_counts, 0, static_cast<size_t>(std::max<uint32_t>(n_parents, 1)) * 8, s));
            CUB_TRY(cudaMemsetAsync(S.collisions, 0, 8, s));
            emit_kernel<<<blocks, 256, 0, s>>>(d_keys, d_key_off, n_keys, d_group_begin, n_groups, h0_mask, S.h0, S.h1, S.ka, S.va, S.seg);
            CUB_TRY(cudaGetLastError());
            gather_u64_kernel<<<blocks, 256, 0, s>>>(S.va, n_keys, S.h1, S.ka);
            CUB_TRY(cudaGetLastError());
            tb = S.temp_bytes;
            CUB_TRY(cub::DeviceRadixSort::SortPairs(S.temp, tb, S.ka, S.kb, S.va, S.vb, n_keys, 0, 64, s));   // by h1
            gather_u64_kernel<<<blocks, 256, 0, s>>>(S.vb, n_keys, S.h0, S.ka);
            CUB_TRY(cudaGetLastError());
            tb = S.temp_bytes;
            CUB_TRY(cub::DeviceRadixSort::SortPairs(S.temp, tb, S.ka, S.kb, S.vb, S.va, n_keys, 0, 64, s));   // by (h0, h1)
            CUB_TRY(cudaMemcpyAsync(S.vb, S.va, n_keys * 4, cudaMemcpyDeviceToDevice, s));
        }
        CUB_TRY(count_by_segment(S, n_keys, S.vb, nullptr, n_groups, use_h1, d_keys, d_key_off, d_group_counts, S.va, s));
        if (d_group_parent && d_parent_counts)
            CUB_TRY(count_by_segment(S, n_keys, S.vb, d_group_parent, n_parents, use_h1, d_keys, d_key_off, d_parent_counts, S.va, s));
        unsigned long long coll = 0;
        CUB_TRY(cudaMemcpyAsync(&coll, S.collisions, 8, cudaMemcpyDeviceToHost, s));
        CUB_TRY(cudaStreamSynchronize(s));
        if (coll == 0) break;   // no equal-hash / different-bytes neighbours: the counts are exact
    }
    return cudaSuccess;
}

size_t count_distinct_scratch_bytes(uint64_t n_keys) {
    const uint64_t n = std::max<uint64_t>(n_keys, 1);
    return 4 * align_up(n * 8) + 5 * align_up(n * 4) + align_up(8) + align_up(temp_bytes_for(n)) + 256;
}

}  // namespace bsg
