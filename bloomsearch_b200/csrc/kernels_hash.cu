// K1 — base hashes of packed keys.
// Replaces bloom/v3 baseHashes() inside every AddString/TestString
// (ingest.go:142; query_exec.go:141,147,154): MurmurHash3_x64_128(seed 0) of the
// key and of key||0x01, one thread per key, aligned 8-byte loads only.
#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

__global__ void __launch_bounds__(256)
hash_keys_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off, uint64_t n_keys,
                 uint64_t* __restrict__ hashes) {
    griddep_launch_dependents();  // PDL: the probe kernel of this batch may start filling its ring now
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_keys) return;
    const uint64_t b = __ldg(&key_off[i]), e = __ldg(&key_off[i + 1]);
    uint64_t h[4];
    base_hashes(keys + b, static_cast<uint32_t>(e - b), h);
    ulonglong2* out = reinterpret_cast<ulonglong2*>(hashes + 4 * i);
    out[0] = make_ulonglong2(h[0], h[1]);
    out[1] = make_ulonglong2(h[2], h[3]);
}

cudaError_t launch_hash_keys(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                             uint64_t* d_hashes, cudaStream_t s) {
    if (n_keys == 0) return cudaSuccess;
    const uint64_t n_blocks = (n_keys + 255) / 256;
    if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    hash_keys_kernel<<<static_cast<uint32_t>(n_blocks), 256, 0, s>>>(d_keys, d_key_off, n_keys, d_hashes);
    return cudaGetLastError();
}

}  // namespace bsg
