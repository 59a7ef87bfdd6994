// Device-side arithmetic shared by every kernel: bloom/v3 base hashes
// (MurmurHash3_x64_128 of key and key||0x01, seed 0), location(h,i), exact
// reduction mod m by a precomputed reciprocal, and the mbarrier / bulk-copy
// (TMA 1-D) PTX wrappers used by the staged probe kernel.
//
// Reference arithmetic: bits-and-blooms/bloom/v3 v3.7.0 bloom.go (baseHashes,
// location) + murmur.go (sum256), reached from ingest.go:142 (AddString) and
// query_exec.go:141-154 (TestString).  sm_100a only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsg {

// ---------------------------------------------------------------- murmur3 ---
constexpr uint64_t kC1 = 0x87c37b91114253d5ULL;
constexpr uint64_t kC2 = 0x4cf5ad432745937fULL;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

__device__ __forceinline__ void bmix(uint64_t& h1, uint64_t& h2, uint64_t k1, uint64_t k2) {
    k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729ULL;
    k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5ULL;
}

// Tail (k1,k2 already zero-padded little-endian words; a zero word mixes to a
// no-op so no branch on the tail length is needed) + finalisation.
__device__ __forceinline__ void finalize(uint64_t h1, uint64_t h2, uint64_t k1, uint64_t k2,
                                         uint64_t total_len, uint64_t& o1, uint64_t& o2) {
    k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; h2 ^= k2;
    k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; h1 ^= k1;
    h1 ^= total_len; h2 ^= total_len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    o1 = h1; o2 = h2;
}

__device__ __forceinline__ uint64_t low_bytes_mask(uint32_t nbytes) {  // nbytes in [0,8]
    return nbytes >= 8 ? ~0ULL : ((1ULL << (8 * nbytes)) - 1ULL);
}

// Streams little-endian 64-bit words out of an arbitrarily aligned byte range
// using only aligned 8-byte loads (the buffer must be readable up to the next
// 8-byte boundary past its end; device key buffers are padded by 16 bytes).
struct WordReader {
    const uint64_t* p;
    uint64_t cur;
    uint32_t sh;  // bit shift 0..56
    __device__ __forceinline__ explicit WordReader(const uint8_t* addr) {
        uintptr_t a = reinterpret_cast<uintptr_t>(addr);
        p = reinterpret_cast<const uint64_t*>(a & ~uintptr_t(7));
        sh = static_cast<uint32_t>(a & 7) * 8;
        cur = __ldg(p);
    }
    __device__ __forceinline__ uint64_t next() {
        uint64_t w;
        ++p;
        if (sh == 0) {
            w = cur;
            cur = __ldg(p);
        } else {
            uint64_t nx = __ldg(p);
            w = (cur >> sh) | (nx << (64 - sh));
            cur = nx;
        }
        return w;
    }
};

// Byte-stream murmur: segments of bytes are fed one after another and mixed 16 bytes at a
// time — used to hash field + "::" + token without materialising the joined key
// (makeFieldTokenKey, tokenizer.go:508-511; addFieldToken, ingest.go:95-102).
struct StreamHasher {
    uint64_t h1 = 0, h2 = 0;   // running murmur state (seed 0)
    uint64_t w[2] = {0, 0};    // pending (not yet mixed) bytes, little-endian
    uint32_t nb = 0;           // number of pending bytes, 0..15
    uint64_t total = 0;        // bytes fed so far

    // append the low n (1..8) bytes of v at byte position nb of the pending 16-byte block
    __device__ __forceinline__ void feed(uint64_t v, uint32_t n) {
        v &= low_bytes_mask(n);
        uint64_t spill = 0;  // bytes that fall beyond the 16-byte block (at most 7)
        if (nb < 8u) {
            const uint32_t sh = nb * 8u;  // 0..56
            w[0] |= v << sh;
            if (sh) w[1] |= v >> (64u - sh);  // zero when everything fitted: v's high bytes are zero
        } else {
            const uint32_t sh = (nb - 8u) * 8u;  // 0..56
            w[1] |= v << sh;
            if (sh) spill = v >> (64u - sh);
        }
        nb += n;
        total += n;
        if (nb >= 16u) {
            bmix(h1, h2, w[0], w[1]);
            w[0] = spill;
            w[1] = 0;
            nb -= 16u;
        }
    }
    __device__ __forceinline__ void feed_bytes(const uint8_t* p, uint32_t len) {
        if (len == 0) return;
        WordReader rd(p);
        uint32_t left = len;
        while (left >= 8u) { feed(rd.next(), 8u); left -= 8u; }
        if (left) feed(rd.next(), left);
    }
    // baseHashes of everything fed so far: murmur(data) and murmur(data || 0x01)
    __device__ __forceinline__ void finish(uint64_t h[4]) const {
        finalize(h1, h2, w[0], w[1], total, h[0], h[1]);
        uint64_t k1 = w[0], k2 = w[1];
        uint64_t g1 = h1, g2 = h2;
        if (nb < 8u) k1 |= 1ull << (8u * nb); else k2 |= 1ull << (8u * (nb - 8u));
        if (nb == 15u) { bmix(g1, g2, k1, k2); k1 = 0; k2 = 0; }
        finalize(g1, g2, k1, k2, total + 1, h[2], h[3]);
    }
};

// baseHashes(key): h[0..1] = murmur(key), h[2..3] = murmur(key || 0x01).
__device__ __forceinline__ void base_hashes(const uint8_t* key, uint32_t len, uint64_t h[4]) {
    uint64_t h1 = 0, h2 = 0;
    const uint32_t nblocks = len >> 4;
    const uint32_t t = len & 15;
    uint64_t k1 = 0, k2 = 0;
    if (len != 0) {
        WordReader rd(key);
        for (uint32_t b = 0; b < nblocks; ++b) {
            uint64_t a = rd.next();
            uint64_t c = rd.next();
            bmix(h1, h2, a, c);
        }
        if (t > 0) k1 = rd.next() & low_bytes_mask(t);
        if (t > 8) k2 = rd.next() & low_bytes_mask(t - 8);
    }
    finalize(h1, h2, k1, k2, len, h[0], h[1]);
    // virtual extra byte 0x01 at position len
    if (t < 8) k1 |= 1ULL << (8 * t); else k2 |= 1ULL << (8 * (t - 8));
    if (t == 15) {  // the extra byte completes a 16-byte block; empty tail
        bmix(h1, h2, k1, k2);
        k1 = 0; k2 = 0;
    }
    finalize(h1, h2, k1, k2, static_cast<uint64_t>(len) + 1, h[2], h[3]);
}

// ------------------------------------------------------ location + modulo ---
// location(h,i) = h[i%2] + i*h[2+(((i+(i%2))%4)/2)]  (uint64 wraparound)
//   i%4: 0 -> h0+i*h2, 1 -> h1+i*h3, 2 -> h0+i*h3, 3 -> h1+i*h2
__device__ __forceinline__ uint64_t location(const uint64_t h[4], uint32_t i) {
    const uint64_t a = (i & 1) ? h[1] : h[0];
    const uint64_t b = (((i + (i & 1)) & 3) >> 1) ? h[3] : h[2];
    return a + static_cast<uint64_t>(i) * b;
}

// x mod m, exact, with inv = floor(2^64/m) for m >= 2 and inv = 2^64-1 for m == 1
// (host: bsg::reciprocal).  q = hi64(x*inv) is floor(x/m) or one less, so a single
// conditional subtraction finishes it.  Requires m < 2^63.
__device__ __forceinline__ uint64_t mod_m(uint64_t x, uint64_t m, uint64_t inv) {
    const uint64_t q = __umul64hi(x, inv);
    uint64_t r = x - q * m;
    if (r >= m) r -= m;
    return r;
}

// Same reduction for m < 2^30 in SIX 32-bit instructions (2 IMAD.WIDE, 2 IMAD, 2 VIADDMNMX) instead of ~35 for the
// emulated 64x64 multiply-high.  With I = inv = floor(2^64/m) = Ih*2^32 + Il and x = xh*2^32 + xl:
//   q_full = xh*Ih + floor((xh*Il + xl*Ih) / 2^32)  is floor(x/m) - {0,1,2}   (it drops two fractional parts of
//   x*I/2^64, each < 1, and x/m - x*I/2^64 is in [0,1)), so x - q_full*m is in [0, 3m), below 2^32 for m < 2^30.
// A value below 2^32 is determined by its residue mod 2^32, and that residue needs only q_full mod 2^32:
//   q = lo32(xh*Ih) + hi32((xh*Il + xl*Ih) mod 2^64),   r = lo32(xl - q*m)  — exact, then two conditional
// subtractions of m (3m -> 2m -> m; no 2m has to be formed), each min(v, v - m): v - m wraps above v exactly when v < m.
// No quotient ever has to fit 32 bits, so there is no separate reduction of xh (rounds 1-2 spent 12 instructions
// on a two-step form).  (proof in DESIGN.md §5; m == 1 works through inv = 2^64-1; checked against 128-bit
// arithmetic by tests/test_host_logic.py::test_mod_m32_formula.)
__device__ __forceinline__ uint32_t mod_m32(uint64_t x, uint32_t m, uint32_t ih, uint32_t il) {
    const uint32_t xh = static_cast<uint32_t>(x >> 32), xl = static_cast<uint32_t>(x);
    uint32_t shi;
    asm("{\n\t.reg .u64 s;\n\t.reg .u32 lo;\n\t"
        "mul.wide.u32 s, %1, %2;\n\t"
        "mad.wide.u32 s, %3, %4, s;\n\t"
        "mov.b64 {lo, %0}, s;\n\t}"
        : "=r"(shi)
        : "r"(xl), "r"(ih), "r"(xh), "r"(il));
    const uint32_t q = xh * ih + shi;
    const uint32_t nm = 0u - m;               // one negation per filter (hoisted), not one per location: written as
    uint32_t r;                               // C++, q * nm + xl is turned back into m * (-q) + xl
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(nm), "r"(xl));   // x - q_full*m, in [0, 3m)
    r = min(r, r - m);   // [0, 3m) -> [0, 2m)
    r = min(r, r - m);   //         -> [0, m)
    return r;
}
// Word index of a bit, opaque to the optimiser: `w32[bit >> 5]` is canonicalised to ((bit >> 3) & ~3) + base
// (SHF, LOP3, IADD); with the shift hidden the address is formed as base + idx * 4 (SHF, LEA).
__device__ __forceinline__ uint32_t word_index(uint32_t bit) {
    uint32_t i;
    asm("shr.u32 %0, %1, 5;" : "=r"(i) : "r"(bit));
    return i;
}
constexpr uint64_t kSmallModLimit = 1ull << 30;

// locations START..k-1 of a survivor (k > START, START in 1..4), in groups of up to four locations:
// the tests of a group are independent (no branch between them -> ILP 4, the dependent chain of a
// test is ~14 instructions + one shared-memory load), the early exit sits between groups.  A warp
// runs until its slowest lane is done anyway, so testing a whole group costs no extra issue slots
// but a quarter of the latency.  location(i): i%4 == 0: h0+i*h2, 1: h1+i*h3, 2: h0+i*h3, 3: h1+i*h2.
// A location index >= k is computed but masked out (its bit index is < m, the load is in bounds).
template <int START>
__device__ __forceinline__ bool test_tail_s32(uint64_t h0, uint64_t h1, uint64_t h2, uint64_t h3, uint32_t m,
                                              uint32_t ih, uint32_t il, uint32_t k,
                                              const uint32_t* __restrict__ w32) {
    static_assert(START >= 1 && START <= 4, "phase A runs 1..4 tests");
    auto probe = [&](uint64_t loc) -> uint32_t {
        const uint32_t bit = mod_m32(loc, m, ih, il);
        return (w32[word_index(bit)] >> (bit & 31u)) & 1u;
    };
    if (START < 4) {  // first group: locations START..3
        uint32_t ok = 1u;
        if (START <= 1) ok &= probe(h1 + h3);
        if (START <= 2) ok &= probe(h0 + 2 * h3) | static_cast<uint32_t>(k <= 2u);
        if (START <= 3) ok &= probe(h1 + 3 * h2) | static_cast<uint32_t>(k <= 3u);
        if (!ok) return false;
    }
    uint64_t ih2 = 4 * h2, ih3 = 4 * h3;  // i*h2, i*h3 at i = 4, 8, ...
    for (uint32_t i = 4; i < k; i += 4) {
        uint32_t ok = probe(h0 + ih2);
        ok &= probe(h1 + ih3 + h3) | static_cast<uint32_t>(i + 1 >= k);
        ok &= probe(h0 + ih3 + 2 * h3) | static_cast<uint32_t>(i + 2 >= k);
        ok &= probe(h1 + ih2 + 3 * h2) | static_cast<uint32_t>(i + 3 >= k);
        if (!ok) return false;
        ih2 += 4 * h2;
        ih3 += 4 * h3;
    }
    return true;
}

// locations START..k-1 of a survivor in groups of FOUR starting at START (so the first group already has
// ILP 4: a phase-B warp is latency bound, and 15/16 of the surviving absent keys die in it).  The location
// pattern i%4 -> (h0|h1) + i*(h2|h3) is fixed per group member because the groups advance by 4.
template <int START>
__device__ __forceinline__ bool test_from_s32(uint64_t h0, uint64_t h1, uint64_t h2, uint64_t h3, uint32_t m,
                                              uint32_t ih, uint32_t il, uint32_t k,
                                              const uint32_t* __restrict__ w32) {
    uint64_t loc[4], step[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        constexpr int dummy = 0;
        (void)dummy;
        const int i = START + j;
        const uint64_t a = (i & 1) ? h1 : h0;
        const uint64_t b = (((i + (i & 1)) & 3) >> 1) ? h3 : h2;
        loc[j] = a + static_cast<uint64_t>(i) * b;
        step[j] = 4 * b;
    }
    for (uint32_t i0 = START; i0 < k; i0 += 4) {
        uint32_t ok = 1u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t bit = mod_m32(loc[j], m, ih, il);
            ok &= ((w32[word_index(bit)] >> (bit & 31u)) & 1u) | static_cast<uint32_t>(i0 + j >= k);
            loc[j] += step[j];
        }
        if (!ok) return false;
    }
    return true;
}

// -------------------------------------------------- mbarrier / bulk copy ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe of a phase (try_wait may suspend the warp for a HW time slice).
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// Wait of a warp that expects to wait long (phase-B warps waiting for phase A): try_wait with a
// suspend-time hint, and a short sleep between polls, so the poll loop does not eat the issue slots
// of the warps it is waiting for (measured: 16 % of all issued instructions were this loop).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t hint_ns, uint32_t sleep_ns) {
    uint32_t ok = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
            : "memory");
        if (ok) break;
        if (sleep_ns) __nanosleep(sleep_ns);
    }
}
__device__ __forceinline__ uint32_t ld_volatile_shared_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); size and both
// addresses must be multiples of 16 bytes.  Completion is signalled on `bar`.
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Coherent 16-byte global load (never the read-only .nc path): for data that the running kernel
// itself may have written earlier (the fused hash table of probe_staged2).
__device__ __forceinline__ ulonglong2 ld_global_u64x2(const void* p) {
    ulonglong2 v;
    asm volatile("ld.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
// Programmatic dependent launch (PDL): a kernel launched with programmatic stream serialization may
// start while its predecessor in the stream is still draining; it must not touch anything the
// predecessor wrote before griddep_wait() returns.  Both are no-ops for a normal launch.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ uint32_t atom_add_acq_rel_shared(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t atom_add_relaxed_shared(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void fence_acq_rel_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace bsg
