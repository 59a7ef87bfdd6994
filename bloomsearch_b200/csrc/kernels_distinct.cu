// (f.3) Exact distinct counts on the device — what the three Go maps of bloomEntrySets exist for
// (ingest.go:24-45 dedup, :105-123 unionInto / counts): the number of distinct entries of every
// group (block x kind) and of every parent (file x kind) union, which is the `n` that sizes the
// filters (ingest.go:139-140).  Emissions may repeat; they are hashed once (the same four base
// hashes the build uses), sorted by (segment, h0, h1, h2, h3) and counted.  Two entries are taken
// as equal when all 256 bits of their base hashes agree (a false merge needs a 256-bit
// MurmurHash3 collision; the byte-compare tie check of SURVEY §8f.3 is not performed).
// Sorting is thrust's merge sort (library code, like calling cuBLAS for a plain GEMM); the hash
// and count kernels are ours.
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>

#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

struct __align__(8) Emission {
    uint64_t h0, h1, h2, h3;
    uint32_t seg;   // segment (group or parent) id
    uint32_t pad;
};

struct EmissionLess {
    __host__ __device__ bool operator()(const Emission& a, const Emission& b) const {
        if (a.seg != b.seg) return a.seg < b.seg;
        if (a.h0 != b.h0) return a.h0 < b.h0;
        if (a.h1 != b.h1) return a.h1 < b.h1;
        if (a.h2 != b.h2) return a.h2 < b.h2;
        return a.h3 < b.h3;
    }
};

// hash every key and tag it with its group (binary search of the CSR group_begin)
__global__ void __launch_bounds__(256)
emit_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off, uint64_t n_keys,
            const uint64_t* __restrict__ group_begin, uint32_t n_groups, Emission* __restrict__ out) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_keys) return;
    const uint64_t b = __ldg(&key_off[i]), e = __ldg(&key_off[i + 1]);
    uint64_t h[4];
    base_hashes(keys + b, static_cast<uint32_t>(e - b), h);
    uint32_t lo = 0, hi = n_groups;  // last g with group_begin[g] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&group_begin[mid]) <= i) lo = mid; else hi = mid;
    }
    out[i] = Emission{h[0], h[1], h[2], h[3], lo, 0};
}

__global__ void __launch_bounds__(256)
retag_kernel(Emission* __restrict__ em, uint64_t n, const uint32_t* __restrict__ group_parent) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) em[i].seg = __ldg(&group_parent[em[i].seg]);
}

// counts[seg] += number of run heads in the sorted array; one atomic per (warp, segment run)
__global__ void __launch_bounds__(256)
count_heads_kernel(const Emission* __restrict__ em, uint64_t n, unsigned long long* __restrict__ counts) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    bool head = false;
    uint32_t seg = 0xffffffffu;
    if (i < n) {
        const Emission a = em[i];
        seg = a.seg;
        head = true;
        if (i > 0) {
            const Emission p = em[i - 1];
            head = p.seg != a.seg || p.h0 != a.h0 || p.h1 != a.h1 || p.h2 != a.h2 || p.h3 != a.h3;
        }
    }
    // lanes of one segment are contiguous (sorted): the first lane of each run adds the run's head count
    const uint32_t same = __match_any_sync(0xffffffffu, seg);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (i < n && lane == static_cast<uint32_t>(__ffs(same) - 1)) {
        const uint32_t c = __popc(heads & same);
        if (c) atomicAdd(&counts[seg], static_cast<unsigned long long>(c));
    }
}

// out_counts[n_segments] zeroed by the caller
static cudaError_t sort_and_count(Emission* d_em, uint64_t n, unsigned long long* d_counts, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    try {
        thrust::sort(thrust::cuda::par.on(s), thrust::device_pointer_cast(d_em), thrust::device_pointer_cast(d_em + n),
                     EmissionLess());
    } catch (...) {
        return cudaErrorUnknown;
    }
    const uint64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    count_heads_kernel<<<static_cast<uint32_t>(blocks), 256, 0, s>>>(d_em, n, d_counts);
    return cudaGetLastError();
}

cudaError_t launch_count_distinct(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                                  const uint64_t* d_group_begin, uint32_t n_groups, const uint32_t* d_group_parent,
                                  void* d_emissions /* n_keys * 40 B scratch */, unsigned long long* d_group_counts,
                                  unsigned long long* d_parent_counts, cudaStream_t s) {
    if (n_keys == 0 || n_groups == 0) return cudaSuccess;
    Emission* em = static_cast<Emission*>(d_emissions);
    const uint64_t blocks = (n_keys + 255) / 256;
    if (blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    emit_kernel<<<static_cast<uint32_t>(blocks), 256, 0, s>>>(d_keys, d_key_off, n_keys, d_group_begin, n_groups, em);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = sort_and_count(em, n_keys, d_group_counts, s);
    if (e != cudaSuccess || !d_group_parent || !d_parent_counts) return e;
    retag_kernel<<<static_cast<uint32_t>(blocks), 256, 0, s>>>(em, n_keys, d_group_parent);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return sort_and_count(em, n_keys, d_parent_counts, s);
}

size_t count_distinct_scratch_bytes(uint64_t n_keys) { return static_cast<size_t>(n_keys) * sizeof(Emission); }

}  // namespace bsg
