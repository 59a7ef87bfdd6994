// K2/K3 — filter construction.
//
// Replaces the AddString loop of buildSizedBloomFilter (ingest.go:139-145),
// called per block from flush.go:204 / merge.go:771 and per file from
// flush.go:253 / merge.go:516.  One CTA per key group: every key is hashed
// once (bloom/v3 baseHashes), its k locations are OR-ed into the group's
// primary (block-level) filter — staged in shared memory with ATOMS.OR when the
// bitset fits, else with RED.OR straight to HBM — and, when the group names a
// secondary filter (the file-level union filter, flush.go:221,253), into that
// one too with RED.OR.  The staged bitset is merged to HBM with RED.OR so that
// several groups (shards of one large entry set) may share a primary filter.
#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

template <typename Sink>
__device__ __forceinline__ void scatter_locations(const uint64_t h[4], uint64_t m, uint64_t inv, uint32_t k,
                                                  Sink&& sink) {
    // location(h,i): i%4 == 0: h0+i*h2, 1: h1+i*h3, 2: h0+i*h3, 3: h1+i*h2
    uint64_t ih2 = 0, ih3 = 0;
    for (uint32_t i = 0; i < k; i += 4) {
        sink(mod_m(h[0] + ih2, m, inv));
        if (i + 1 >= k) break;
        sink(mod_m(h[1] + ih3 + h[3], m, inv));
        if (i + 2 >= k) break;
        sink(mod_m(h[0] + ih3 + 2 * h[3], m, inv));
        if (i + 3 >= k) break;
        sink(mod_m(h[1] + ih2 + 3 * h[2], m, inv));
        ih2 += 4 * h[2];
        ih3 += 4 * h[3];
    }
}

// Same for m < 2^30 (every block filter and every file filter below 128 MB): the exact 32-bit reduction
// mod_m32 (9 instructions) instead of the emulated 64x64 multiply-high of mod_m (~35).  ncu of round 2's first
// build capture: 612 thread instructions per key, FMA-heavy pipe 75 % busy, most of it the 64-bit modulo.
template <typename Sink>
__device__ __forceinline__ void scatter_locations32(const uint64_t h[4], uint32_t m, uint32_t ih, uint32_t il, uint32_t k,
                                                    Sink&& sink) {
    uint64_t ih2 = 0, ih3 = 0;
    for (uint32_t i = 0; i < k; i += 4) {
        sink(mod_m32(h[0] + ih2, m, ih, il));
        if (i + 1 >= k) break;
        sink(mod_m32(h[1] + ih3 + h[3], m, ih, il));
        if (i + 2 >= k) break;
        sink(mod_m32(h[0] + ih3 + 2 * h[3], m, ih, il));
        if (i + 3 >= k) break;
        sink(mod_m32(h[1] + ih2 + 3 * h[2], m, ih, il));
        ih2 += 4 * h[2];
        ih3 += 4 * h[3];
    }
}

// k locations of one key into the primary filter (shared-memory staged or global) and the optional secondary one
__device__ __forceinline__ void insert_key(const uint64_t h[4], const BuildFilter& f1, const BuildFilter& f2, bool staged,
                                           bool has2, uint32_t* s32, unsigned long long* g1, unsigned long long* g2) {
    if (f1.m < kSmallModLimit) {
        const uint32_t m = static_cast<uint32_t>(f1.m), ih = static_cast<uint32_t>(f1.inv >> 32), il = static_cast<uint32_t>(f1.inv);
        if (staged)
            scatter_locations32(h, m, ih, il, f1.k, [&](uint32_t bit) { atomicOr(&s32[bit >> 5], 1u << (bit & 31u)); });
        else
            scatter_locations32(h, m, ih, il, f1.k, [&](uint32_t bit) { atomicOr(&g1[bit >> 6], 1ull << (bit & 63u)); });
    } else if (staged) {
        scatter_locations(h, f1.m, f1.inv, f1.k, [&](uint64_t bit) {
            atomicOr(&s32[static_cast<uint32_t>(bit >> 5)], 1u << (static_cast<uint32_t>(bit) & 31u));
        });
    } else {
        scatter_locations(h, f1.m, f1.inv, f1.k, [&](uint64_t bit) {
            atomicOr(&g1[bit >> 6], 1ull << (static_cast<uint32_t>(bit) & 63u));
        });
    }
    if (has2) {
        if (f2.m < kSmallModLimit)
            scatter_locations32(h, static_cast<uint32_t>(f2.m), static_cast<uint32_t>(f2.inv >> 32), static_cast<uint32_t>(f2.inv),
                                f2.k, [&](uint32_t bit) { atomicOr(&g2[bit >> 6], 1ull << (bit & 63u)); });
        else
            scatter_locations(h, f2.m, f2.inv, f2.k, [&](uint64_t bit) {
                atomicOr(&g2[bit >> 6], 1ull << (static_cast<uint32_t>(bit) & 63u));
            });
    }
}

__global__ void __launch_bounds__(256)
build_kernel(const uint8_t* __restrict__ keys, const uint64_t* __restrict__ key_off,
             const uint64_t* __restrict__ group_begin, const uint32_t* __restrict__ group_filter,
             const uint32_t* __restrict__ group_filter2, const BuildFilter* __restrict__ filters,
             uint64_t* __restrict__ out_words, uint32_t smem_cap_words64) {
    extern __shared__ __align__(16) uint64_t s_words[];
    const uint32_t g = blockIdx.x;
    const BuildFilter f1 = filters[group_filter[g]];
    const uint32_t f2_id = group_filter2 ? group_filter2[g] : BSG_NO_FILTER;
    const bool has2 = f2_id != BSG_NO_FILTER;
    BuildFilter f2 = {0, 1, 0, 0, 0};
    if (has2) f2 = filters[f2_id];
    const uint64_t kb = group_begin[g], ke = group_begin[g + 1];
    const bool staged = f1.nwords <= smem_cap_words64;

    if (staged) {
        for (uint32_t w = threadIdx.x; w < f1.nwords; w += blockDim.x) s_words[w] = 0;
        __syncthreads();
    }
    uint32_t* s32 = reinterpret_cast<uint32_t*>(s_words);
    // global merges use 64-bit RED.OR only (no mixed-size atomics on one word)
    unsigned long long* g1 = reinterpret_cast<unsigned long long*>(out_words + f1.word_off);
    unsigned long long* g2 = reinterpret_cast<unsigned long long*>(out_words + f2.word_off);

    for (uint64_t i = kb + threadIdx.x; i < ke; i += blockDim.x) {
        const uint64_t b = __ldg(&key_off[i]), e = __ldg(&key_off[i + 1]);
        uint64_t h[4];
        base_hashes(keys + b, static_cast<uint32_t>(e - b), h);
        insert_key(h, f1, f2, staged, has2, s32, g1, g2);
    }
    if (staged) {
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < f1.nwords; w += blockDim.x) {
            const uint64_t v = s_words[w];
            if (v) atomicOr(&g1[w], static_cast<unsigned long long>(v));
        }
    }
}

// Fused field::token build (§8 f.2): entry i of a group is the pair (path_of[i], token_of[i]) of
// indexes into one string table; its key is strings[path] + "::" + strings[token]
// (makeFieldTokenKey, tokenizer.go:508-511), hashed as a byte stream without materialising it.
__global__ void __launch_bounds__(256)
build_ft_kernel(const uint8_t* __restrict__ strings, const uint64_t* __restrict__ str_off,
                const uint32_t* __restrict__ pair_path, const uint32_t* __restrict__ pair_token,
                const uint64_t* __restrict__ group_begin, const uint32_t* __restrict__ group_filter,
                const uint32_t* __restrict__ group_filter2, const BuildFilter* __restrict__ filters,
                uint64_t* __restrict__ out_words, uint32_t smem_cap_words64) {
    extern __shared__ __align__(16) uint64_t s_words[];
    const uint32_t g = blockIdx.x;
    const BuildFilter f1 = filters[group_filter[g]];
    const uint32_t f2_id = group_filter2 ? group_filter2[g] : BSG_NO_FILTER;
    const bool has2 = f2_id != BSG_NO_FILTER;
    BuildFilter f2 = {0, 1, 0, 0, 0};
    if (has2) f2 = filters[f2_id];
    const uint64_t kb = group_begin[g], ke = group_begin[g + 1];
    const bool staged = f1.nwords <= smem_cap_words64;
    if (staged) {
        for (uint32_t w = threadIdx.x; w < f1.nwords; w += blockDim.x) s_words[w] = 0;
        __syncthreads();
    }
    uint32_t* s32 = reinterpret_cast<uint32_t*>(s_words);
    unsigned long long* g1 = reinterpret_cast<unsigned long long*>(out_words + f1.word_off);
    unsigned long long* g2 = reinterpret_cast<unsigned long long*>(out_words + f2.word_off);
    for (uint64_t i = kb + threadIdx.x; i < ke; i += blockDim.x) {
        const uint32_t pi = __ldg(&pair_path[i]), ti = __ldg(&pair_token[i]);
        const uint64_t pb = __ldg(&str_off[pi]), pe = __ldg(&str_off[pi + 1]);
        const uint64_t tb = __ldg(&str_off[ti]), te = __ldg(&str_off[ti + 1]);
        StreamHasher sh;
        sh.feed_bytes(strings + pb, static_cast<uint32_t>(pe - pb));
        sh.feed(0x3a3aull, 2);  // "::"
        sh.feed_bytes(strings + tb, static_cast<uint32_t>(te - tb));
        uint64_t h[4];
        sh.finish(h);
        insert_key(h, f1, f2, staged, has2, s32, g1, g2);
    }
    if (staged) {
        __syncthreads();
        for (uint32_t w = threadIdx.x; w < f1.nwords; w += blockDim.x) {
            const uint64_t v = s_words[w];
            if (v) atomicOr(&g1[w], static_cast<unsigned long long>(v));
        }
    }
}

static int g_build_max_smem = 48 * 1024;

cudaError_t build_configure(int max_smem_optin) {
    g_build_max_smem = max_smem_optin;
    cudaError_t e = cudaFuncSetAttribute(build_ft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
}

cudaError_t launch_build_ft(const uint8_t* d_strings, const uint64_t* d_str_off, const uint32_t* d_pair_path,
                            const uint32_t* d_pair_token, const uint64_t* d_group_begin, uint32_t n_groups,
                            const uint32_t* d_group_filter, const uint32_t* d_group_filter2,
                            const BuildFilter* d_filters, uint64_t* d_out_words, uint32_t smem_cap_bytes,
                            cudaStream_t s) {
    if (n_groups == 0) return cudaSuccess;
    if (smem_cap_bytes > static_cast<uint32_t>(g_build_max_smem)) smem_cap_bytes = g_build_max_smem;
    smem_cap_bytes &= ~15u;
    build_ft_kernel<<<n_groups, 256, smem_cap_bytes, s>>>(d_strings, d_str_off, d_pair_path, d_pair_token,
                                                          d_group_begin, d_group_filter, d_group_filter2, d_filters,
                                                          d_out_words, smem_cap_bytes / 8);
    return cudaGetLastError();
}

cudaError_t launch_build(const uint8_t* d_keys, const uint64_t* d_key_off, const uint64_t* d_group_begin,
                         uint32_t n_groups, const uint32_t* d_group_filter, const uint32_t* d_group_filter2,
                         const BuildFilter* d_filters, uint64_t* d_out_words, uint32_t smem_cap_bytes,
                         cudaStream_t s) {
    if (n_groups == 0) return cudaSuccess;
    if (smem_cap_bytes > static_cast<uint32_t>(g_build_max_smem)) smem_cap_bytes = g_build_max_smem;
    smem_cap_bytes &= ~15u;
    build_kernel<<<n_groups, 256, smem_cap_bytes, s>>>(d_keys, d_key_off, d_group_begin, d_group_filter,
                                                       d_group_filter2, d_filters, d_out_words, smem_cap_bytes / 8);
    return cudaGetLastError();
}

__global__ void or_words_kernel(uint64_t* __restrict__ dst, const uint64_t* __restrict__ src, uint64_t n) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] |= src[i];
}

cudaError_t launch_or_words(uint64_t* d_dst, const uint64_t* d_src, uint64_t n_words, cudaStream_t s) {
    if (n_words == 0) return cudaSuccess;
    uint64_t blocks = (n_words + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    or_words_kernel<<<static_cast<uint32_t>(blocks), 256, 0, s>>>(d_dst, d_src, n_words);
    return cudaGetLastError();
}

}  // namespace bsg
