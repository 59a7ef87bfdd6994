// Corpus re-layout: copies filter words from the caller's layout into the
// probe layout (unit-contiguous, every filter 16-byte aligned), byte-swapping
// the on-disk big-endian words of bitset.WriteTo on the device — the decode the
// reference performs per block per query in parseFilterSection / ReadFrom
// (file_format.go:392-448) happens here once, at load time.
#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
    const uint32_t lo = static_cast<uint32_t>(v), hi = static_cast<uint32_t>(v >> 32);
    return (static_cast<uint64_t>(__byte_perm(lo, 0, 0x0123)) << 32) | __byte_perm(hi, 0, 0x0123);
}

// One CTA per filter slot (unit*3+kind).
__global__ void __launch_bounds__(256)
repack_kernel(const uint64_t* __restrict__ src, const uint64_t* __restrict__ src_off,
              const DevFilter* __restrict__ udesc, uint64_t* __restrict__ dst, int big_endian) {
    const uint64_t f = blockIdx.x;
    const DevFilter d = udesc[f];
    if (d.m == 0) return;
    const uint64_t* s = src + src_off[f];
    uint64_t* o = dst + d.word_off;
    for (uint32_t w = threadIdx.x; w < d.nwords; w += blockDim.x) {
        uint64_t v = __ldg(&s[w]);
        o[w] = big_endian ? bswap64(v) : v;
    }
}

cudaError_t launch_repack(const uint64_t* d_src, const uint64_t* d_src_off, const DevFilter* d_udesc,
                          uint64_t n_filters, uint64_t* d_dst, int big_endian, cudaStream_t s) {
    if (n_filters == 0) return cudaSuccess;
    if (n_filters > 0x7fffffffull) return cudaErrorInvalidValue;
    repack_kernel<<<static_cast<uint32_t>(n_filters), 256, 0, s>>>(d_src, d_src_off, d_udesc, d_dst, big_endian);
    return cudaGetLastError();
}

}  // namespace bsg
