// Device-side decode of raw filter sections (file_format.go:343-385 framing):
//   [u8 flags][per present filter: u32 LE length | u64 BE m | u64 BE k | u64 BE bitlen | words u64 BE][u32 LE CRC32C]
// parse_sections_kernel does what parseFilterSection (file_format.go:392-448) does per
// block per query in the reference — CRC32C (Castagnoli) over the payload, flags / length
// checks, bloom header decode — once, at load time, one warp per section.
// repack_sections_kernel then copies every filter's big-endian words (arbitrarily aligned
// inside the byte stream) into the probe layout, byte-swapped to native order.
#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

constexpr uint32_t kCrc32cPoly = 0x82F63B78u;  // reflected Castagnoli, file_format.go:44

__constant__ uint32_t c_x2n[32];  // x^(2^n) mod P, reflected (for CRC combination)

static uint32_t h_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kCrc32cPoly : b >> 1;
    }
    return p;
}

cudaError_t sections_configure() {
    uint32_t t[32];
    uint32_t p = 1u << 30;  // x^1
    t[0] = p;
    for (int n = 1; n < 32; ++n) t[n] = p = h_multmodp(p, p);
    return cudaMemcpyToSymbol(c_x2n, t, sizeof(t));
}

__device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kCrc32cPoly : b >> 1;
    }
    return p;
}
// x^(8*n) mod P
__device__ __forceinline__ uint32_t x8nmodp(uint64_t n) {
    uint32_t p = 1u << 31;  // x^0
    uint32_t k = 3;
    while (n) {
        if (n & 1) p = multmodp(c_x2n[k & 31], p);
        n >>= 1;
        ++k;
    }
    return p;
}
// crc(A || B) from crc(A), crc(B), len(B) — standard (pre/post-conditioned) CRC values
__device__ __forceinline__ uint32_t crc_combine(uint32_t crc_a, uint32_t crc_b, uint64_t len_b) {
    return multmodp(x8nmodp(len_b), crc_a) ^ crc_b;
}

__device__ __forceinline__ uint64_t load_be64(const uint8_t* p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
    return v;
}
__device__ __forceinline__ uint32_t load_le32(const uint8_t* p) {
    return static_cast<uint32_t>(p[0]) | (static_cast<uint32_t>(p[1]) << 8) | (static_cast<uint32_t>(p[2]) << 16) |
           (static_cast<uint32_t>(p[3]) << 24);
}

// one warp per section
__global__ void __launch_bounds__(256)
parse_sections_kernel(const uint8_t* __restrict__ sections, const uint64_t* __restrict__ sec_off, uint64_t n_units,
                      int verify_crc, SectionInfo* __restrict__ info) {
    __shared__ uint32_t s_table[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
#pragma unroll
        for (int j = 0; j < 8; ++j) c = (c & 1) ? (c >> 1) ^ kCrc32cPoly : c >> 1;
        s_table[i] = c;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t u = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (u >= n_units) return;
    const uint64_t b = sec_off[u], e = sec_off[u + 1];
    const uint8_t* sec = sections + b;
    const uint64_t len = e - b;
    int32_t status = 0;
    SectionInfo out;
    out.present = 0;
    for (int k = 0; k < 3; ++k) { out.m[k] = 0; out.k[k] = 0; out.words_byte_off[k] = 0; }

    if (len < 4 + 1) {
        status = -1;  // "bloom filter section too small"
    } else {
        const uint64_t plen = len - 4;
        if (verify_crc) {
            // each lane: standard CRC32C of its contiguous chunk; then a shuffle tree combines them
            const uint64_t chunk = (plen + 31) / 32;
            const uint64_t lo = min(plen, chunk * lane), hi = min(plen, chunk * (lane + 1));
            uint32_t c = 0xFFFFFFFFu;
            for (uint64_t i = lo; i < hi; ++i) c = s_table[(c ^ sec[i]) & 0xffu] ^ (c >> 8);
            uint32_t crc = c ^ 0xFFFFFFFFu;
            uint64_t clen = hi - lo;
            for (uint32_t d = 1; d < 32; d <<= 1) {
                const uint32_t crc_hi = __shfl_down_sync(0xffffffffu, crc, d);
                const uint64_t len_hi = __shfl_down_sync(0xffffffffu, clen, d);
                if ((lane & (2 * d - 1)) == 0) {
                    crc = crc_combine(crc, crc_hi, len_hi);
                    clen += len_hi;
                }
            }
            const uint32_t actual = __shfl_sync(0xffffffffu, crc, 0);
            if (actual != load_le32(sec + plen)) status = -2;  // ErrInvalidHash
        }
        if (status == 0) {
            const uint8_t flags = sec[0];
            if (flags & ~7u) {
                status = -3;  // unrecognized leading byte
            } else {
                uint64_t pos = 1;
                for (int k = 0; k < 3 && status == 0; ++k) {
                    if (!(flags & (1u << k))) continue;
                    if (plen - pos < 4) { status = -4; break; }  // truncated length prefix
                    const uint64_t flen = load_le32(sec + pos);
                    pos += 4;
                    if (flen > plen - pos) { status = -5; break; }  // length exceeds remainder
                    if (flen < 24) { status = -6; break; }          // bloom header does not fit
                    const uint64_t m = load_be64(sec + pos), kk = load_be64(sec + pos + 8),
                                   bitlen = load_be64(sec + pos + 16);
                    const uint64_t nw = (bitlen + 63) >> 6;
                    // the probe needs bitlen == m (bloom.New always writes it so) and sane sizes
                    if (bitlen != m || m == 0 || m > (1ull << 62) || kk == 0 || kk > 0x7fffffffull ||
                        flen != 24 + 8 * nw) { status = -6; break; }
                    out.m[k] = m;
                    out.k[k] = kk;
                    out.words_byte_off[k] = b + pos + 24;
                    out.present |= 1u << k;
                    pos += flen;
                }
                if (status == 0 && pos != plen) status = -7;  // trailing bytes
            }
        }
    }
    if (lane == 0) {
        if (status != 0) {
            out.present = 0;
            for (int k = 0; k < 3; ++k) { out.m[k] = 0; out.k[k] = 0; out.words_byte_off[k] = 0; }
        }
        out.status = status;
        info[u] = out;
    }
}

cudaError_t launch_parse_sections(const uint8_t* d_sections, const uint64_t* d_sec_off, uint64_t n_units,
                                  int verify_crc, SectionInfo* d_info, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_blocks = (n_units + 7) / 8;  // 8 warps per block
    if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    parse_sections_kernel<<<static_cast<uint32_t>(n_blocks), 256, 0, s>>>(d_sections, d_sec_off, n_units, verify_crc,
                                                                         d_info);
    return cudaGetLastError();
}

// One CTA per filter slot (unit*3+kind): big-endian words at an arbitrary byte offset -> native
// words in the probe layout.
__global__ void __launch_bounds__(256)
repack_sections_kernel(const uint8_t* __restrict__ sections, const SectionInfo* __restrict__ info,
                       const DevFilter* __restrict__ udesc, uint64_t* __restrict__ dst) {
    const uint64_t f = blockIdx.x;
    const DevFilter d = udesc[f];
    if (d.m == 0) return;
    const uint8_t* src = sections + info[f / 3].words_byte_off[f % 3];
    uint64_t* o = dst + d.word_off;
    // aligned 8-byte loads + funnel shift (the section buffer is padded by 16 bytes)
    const uintptr_t a = reinterpret_cast<uintptr_t>(src);
    const uint64_t* p = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
    const uint32_t sh = static_cast<uint32_t>(a & 7) * 8;
    for (uint32_t w = threadIdx.x; w < d.nwords; w += blockDim.x) {
        uint64_t lo = __ldg(p + w);
        uint64_t v = lo;
        if (sh) {
            const uint64_t hi = __ldg(p + w + 1);
            v = (lo >> sh) | (hi << (64 - sh));
        }
        // v holds the 8 bytes in memory order (little-endian load): big-endian value = bswap
        const uint32_t vl = static_cast<uint32_t>(v), vh = static_cast<uint32_t>(v >> 32);
        o[w] = (static_cast<uint64_t>(__byte_perm(vl, 0, 0x0123)) << 32) | __byte_perm(vh, 0, 0x0123);
    }
}

cudaError_t launch_repack_sections(const uint8_t* d_sections, const SectionInfo* d_info, const DevFilter* d_udesc,
                                   uint64_t n_units, uint64_t* d_dst, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    if (n_units * 3 > 0x7fffffffull) return cudaErrorInvalidValue;
    repack_sections_kernel<<<static_cast<uint32_t>(n_units * 3), 256, 0, s>>>(d_sections, d_info, d_udesc, d_dst);
    return cudaGetLastError();
}

}  // namespace bsg
