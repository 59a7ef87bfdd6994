// K4 — the staged MaybeContains probe over a ring of TILES (the default staged path).
//
// Replaces the reference's per-block loop (query_exec.go:572-615): for every unit and every key of
// the batch, TestString on the unit's filter of the key's kind (query_exec.go:128-159).  Same bits
// as probe_staged / probe_staged2 (kernels_probe.cu) and as the gather path; what changed is how
// the work is cut:
//
//   * the ring's unit is a TILE (bsg_internal.h): several small units per stage (one mbarrier wait,
//     one refill and one descriptor copy serve up to 8 units), or one unit's {field,token} /
//     {fieldtoken} filters per stage for large units (half-size stages: a ring twice as deep, first
//     data resident twice as early);
//   * the batch's keys are sorted by kind on the host (slotinfo: sorted slot -> caller index | kind),
//     so in KIND mode only the warps that hold keys of the tile's kinds touch it;
//   * every CTA keeps the batch's base hashes in shared memory (hashed in place by the CTA while the
//     first fills are in flight, or copied from the hash kernel's output): phase B never leaves the SM;
//   * TestString is an AND over the k locations, so the order in which clear bits are found is free: the
//     kernel runs it in three ROUNDS per tile with the survivors re-compacted in between (see the kernel),
//     instead of one early-exit loop per lane (a third of the lanes busy) or specialised warps.
#include <type_traits>

#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

struct ProbeTilesArgs {
    const TileRec* tiles;
    const uint32_t* n_items_dev;   // nullable: item count of a device-compacted list (hierarchical probe)
    uint32_t n_items_host;
    uint32_t parts;                // tiles per item (1 UNIT mode, 2 KIND mode)
    const uint64_t* words;
    const uint64_t* hashes;        // nullable: base hashes in caller key order (hash_keys_kernel)
    const uint8_t* key_bytes;      // fused hashing (hashes == nullptr)
    const uint64_t* key_off;
    const uint16_t* slotinfo;      // [key_base + slot] = caller index within the pass | kind << 14
    uint32_t key_base, n_keys, kind_mask;
    uint32_t* matrix32;
    uint32_t row_words32;
    uint32_t n_stages, stage_bytes, hdr_bytes, units_cap;
    uint64_t* trace;
    uint32_t trace_slots;
};

__device__ __forceinline__ void fill_tile(uint8_t* st, uint64_t* bar, const TileRec* rec, bool has_next,
                                          const TileRec* next_rec, const uint64_t* __restrict__ words,
                                          const uint4 f0, const uint16_t* nb16, uint32_t kind_mask,
                                          uint32_t hdr_bytes) {
    // f0 = first 16 bytes of the tile's TileFill: word_base (lo, hi), data_bytes, rec_bytes
    const uint64_t word_base = (static_cast<uint64_t>(f0.y) << 32) | f0.x;
    const uint32_t data_bytes = f0.z, rec_bytes = f0.w;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(words + word_base);
    uint8_t* data = st + hdr_bytes;
    if (kind_mask == 7u) {
        mbar_arrive_expect_tx(bar, rec_bytes + (has_next ? 64u : 0u) + data_bytes);
        bulk_g2s(st, rec, rec_bytes, bar);
        if (has_next) bulk_g2s(st + kTileNextFillOff, &next_rec->fill, 64u, bar);
        if (data_bytes) bulk_g2s(data, src, data_bytes, bar);
        return;
    }
    // the batch touches only some kinds: copy only those filters (runs of wanted neighbours merged)
    const uint32_t n_slots = ((rec_bytes - kTileRecFixedBytes) / 48u) * 3u;
    uint32_t want = 0;
    for (uint32_t i = 0; i < n_slots; ++i)
        if ((kind_mask >> (i % 3u)) & 1u) want += nb16[i];
    mbar_arrive_expect_tx(bar, rec_bytes + (has_next ? 64u : 0u) + want * 16u);
    bulk_g2s(st, rec, rec_bytes, bar);
    if (has_next) bulk_g2s(st + kTileNextFillOff, &next_rec->fill, 64u, bar);
    uint32_t off = 0, run_off = 0, run = 0;
    for (uint32_t i = 0; i < n_slots; ++i) {
        const uint32_t b = static_cast<uint32_t>(nb16[i]) * 16u;
        if ((kind_mask >> (i % 3u)) & 1u) {
            if (run == 0) run_off = off;
            run += b;
        } else if (run) {
            bulk_g2s(data + run_off, src + run_off, run, bar);
            run = 0;
        }
        off += b;
    }
    if (run) bulk_g2s(data + run_off, src + run_off, run, bar);
}

// NT membership tests of one key against one staged filter, no branch between them.
// f = {m, ih, il, rel << 8 | k}, rel counted from the stage start `st`; SMALLK: some filter of the tile has
// k < 4, so location t exists only when t < k.
template <int NT, bool SMALLK>
__device__ __forceinline__ bool first_tests(const uint64_t (&loc)[NT], const uint4 f, const uint8_t* st) {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(st + tile_rel(f.w));
    uint32_t bit[NT], wv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) bit[t] = mod_m32(loc[t], f.x, f.y, f.z);
#pragma unroll
    for (int t = 0; t < NT; ++t) wv[t] = w32[word_index(bit[t])];
    uint32_t pass = 1u;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        uint32_t b = (wv[t] >> (bit[t] & 31u)) & 1u;
        if (SMALLK) b |= static_cast<uint32_t>(tile_k(f.w) <= static_cast<uint32_t>(t));
        pass &= b;
    }
    return pass != 0u;
}

// One tile at a time, the whole CTA in lock step, three rounds separated by CTA barriers:
//   round A   thread t owns sorted key slot t (locations 0..NT-1 in registers).  For every unit of the tile:
//             NT branch-free tests; survivors (one in 2^NT of the absent keys, and every present key) are
//             appended to list L1 as (unit, slot) with one shared-memory atomic per warp per unit.
//   round B1  L1 is dense: entry e goes to thread e mod 1024.  Locations NT..NT+3 as four independent tests.
//             A key that fails is final; a key whose last location was among them is final (bit set in the
//             unit's result row); the rest (one in 16 of L1's absent keys) are appended to L2.
//   round B2  L2, dense again: the remaining locations in groups of four with the early exit between groups.
// Then the rows of the tile's units are written with one coalesced 128-byte store each and thread 0 refills
// the stage.  Every round runs on all 32 warps, so no warp role can starve another (measured with
// specialised A / B warps: the scheduler favoured the A warps and phase B ran 5-10x slower than its
// instruction count), and the geometric tail of TestString is re-compacted twice instead of
// idling lanes.
template <int NT, int NTHR, bool TRACE>
__global__ void __launch_bounds__(NTHR, 1024 / NTHR) probe_tiles_kernel(const ProbeTilesArgs a) {
    static_assert(NT >= 1 && NT <= 4 && (NTHR == 1024 || NTHR == 512), "shape");
    constexpr int KPT = static_cast<int>(kProbeMaxKeysPerPass) / NTHR;   // key slots per thread in round A
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + 128);          // cnt[0] = |L1|, cnt[1] = |L2|
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(smem + kTilesPrefixBytes);
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem + kTilesPrefixBytes + kTilesSlotInfoBytes);
    uint16_t* L1 = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(rows) + a.units_cap * 128u);
    uint16_t* L2 = L1 + a.units_cap * kProbeMaxKeysPerPass;
    ulonglong2* htab = reinterpret_cast<ulonglong2*>(L2 + a.units_cap * kProbeMaxKeysPerPass);
    const uint32_t hash_bytes = ((a.n_keys + 31u) & ~31u) * 32u;
    uint8_t* stages = reinterpret_cast<uint8_t*>(htab) + hash_bytes;   // 128-byte aligned: every term is

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t warp = tid >> 5;
    const uint32_t G = gridDim.x;
    const uint32_t S = a.n_stages;
    const uint32_t P = a.parts;
    // optional timeline (profiling only), per CTA: [0] start, [1] hashes ready; per tile n, base b = 2+8n:
    // [b] resident, [b+1] round A done, [b+2] round B1 done, [b+3] round B2 done, [b+4] |L1|, [b+5] |L2|
    uint64_t* tr = (TRACE && a.trace) ? a.trace + static_cast<size_t>(blockIdx.x) * a.trace_slots : nullptr;
    if (TRACE && tr && tid == 0) tr[0] = globaltimer_ns();

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) mbar_init(&full[s], 1);
        cnt[0] = 0;
        cnt[1] = 0;
        fence_barrier_init();
    }
    for (uint32_t i = tid; i < a.units_cap * 32u; i += blockDim.x) rows[i] = 0;
    __syncthreads();

    // PDL: the first fills read only the immutable corpus, so they may overlap the tail of the previous
    // kernel in the stream; everything a predecessor wrote (hashes, a compacted item list) is touched
    // after the wait.
    if (a.n_items_dev) griddep_wait();
    const uint32_t n_items = a.n_items_dev ? __ldg(a.n_items_dev) : a.n_items_host;
    const uint32_t my_items = n_items > blockIdx.x ? (n_items - blockIdx.x + G - 1) / G : 0;
    const uint32_t my_tiles = my_items * P;
    // position n of this CTA's tile sequence -> record: item (blockIdx + (n / P) * G), part n % P
    auto rec_of = [&](uint32_t n) -> const TileRec* {
        const uint32_t item = blockIdx.x + (P == 1 ? n : (n >> 1)) * G;
        return a.tiles + (P == 1 ? item : item * 2u + (n & 1u));
    };

    if (warp == 0 && lane < S && lane < my_tiles) {  // prologue: lane l fills stage l with tile l
        const TileRec* r = rec_of(lane);
        const uint4* fp = reinterpret_cast<const uint4*>(&r->fill);
        const uint4 f0 = __ldg(fp);
        uint16_t nb16[kTileMaxUnits * 3] = {};
        if (a.kind_mask != 7u) {
            const uint4 q1 = __ldg(fp + 1), q2 = __ldg(fp + 2), q3 = __ldg(fp + 3);
            const uint32_t w[12] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int i = 0; i < 12; ++i) { nb16[2 * i] = static_cast<uint16_t>(w[i]); nb16[2 * i + 1] = static_cast<uint16_t>(w[i] >> 16); }
        }
        const bool has_next = lane + S < my_tiles;
        fill_tile(stages + static_cast<size_t>(lane) * a.stage_bytes, &full[lane], r, has_next,
                  has_next ? rec_of(lane + S) : r, a.words, f0, nb16, a.kind_mask, a.hdr_bytes);
    }
    griddep_launch_dependents();
    if (!a.n_items_dev) griddep_wait();

    // ---- the batch: slot table + base hashes into shared memory (sorted-slot order) ----
    for (uint32_t t = tid; t < a.n_keys; t += blockDim.x) {
        const uint32_t si = __ldg(&a.slotinfo[a.key_base + t]);
        s_slot[t] = static_cast<uint16_t>(si);
        const uint32_t q = a.key_base + (si & 0x3ffu);
        ulonglong2 h01, h23;
        if (a.hashes) {
            const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(a.hashes + 4ull * q);
            h01 = __ldg(hp);
            h23 = __ldg(hp + 1);
        } else {
            const uint64_t b = __ldg(&a.key_off[q]), e = __ldg(&a.key_off[q + 1]);
            uint64_t h[4];
            base_hashes(a.key_bytes + b, static_cast<uint32_t>(e - b), h);
            h01 = make_ulonglong2(h[0], h[1]);
            h23 = make_ulonglong2(h[2], h[3]);
        }
        htab[2 * t] = h01;
        htab[2 * t + 1] = h23;
    }
    __syncthreads();
    if (TRACE && tr && tid == 0) tr[1] = globaltimer_ns();

    // ---- this thread's keys for round A: sorted slots tid + j*NTHR ----
    uint64_t loc[KPT][NT];   // location(h, 0..NT-1) = h0, h1+h3, h0+2*h3, h1+3*h2
    uint32_t koff[KPT], kbit[KPT];
    uint32_t warp_kinds = 0;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const uint32_t slot = tid + j * NTHR;
#pragma unroll
        for (int t = 0; t < NT; ++t) loc[j][t] = 0;
        koff[j] = 0;
        kbit[j] = 0;
        if (slot < a.n_keys) {
            const ulonglong2 x = htab[2 * slot], y = htab[2 * slot + 1];
            const uint64_t l4[4] = {x.x, x.y + y.y, x.x + 2 * y.y, x.y + 3 * y.x};
#pragma unroll
            for (int t = 0; t < NT; ++t) loc[j][t] = l4[t];
            const uint32_t kind = s_slot[slot] >> 14;
            koff[j] = kind * 16u;
            kbit[j] = 1u << kind;
        }
        warp_kinds |= kbit[j];
    }
    warp_kinds = __reduce_or_sync(0xffffffffu, warp_kinds);
    const uint32_t lt_mask = (1u << lane) - 1u;
    // every word of the row that belongs to this pass, pad word included (the smem rows are zero there)
    const uint32_t out_words = min(32u, a.row_words32 - (a.key_base >> 5));
    uint32_t* out_base = a.matrix32 + (a.key_base >> 5);

    uint32_t s = 0, ph = 0;
    uint8_t* st = stages;
    for (uint32_t n = 0; n < my_tiles; ++n) {
        mbar_wait(&full[s], ph);
        if (TRACE && tr && tid == 0 && 2 + 8 * n < a.trace_slots) tr[2 + 8 * n] = globaltimer_ns();
        const uint4 head = *reinterpret_cast<const uint4*>(st);  // n_units, part_kinds, flags
        const uint32_t my_unit = warp < head.x ? *reinterpret_cast<const uint32_t*>(st + 16 + 64 + 4 * warp) : 0u;
        // ---------------------------------------------------------------- round A ---
        if (warp_kinds & head.y) {
            // every thread keeps, per key slot, the set of units whose first tests its key passed: one funnel shift
            // per (key, unit) pushes the AND of the tested bits into the mask from the top (bit 0 of `p` is the
            // verdict, the bits above it are junk and fall off), no predicate, no ballot, no shared-memory traffic
            // inside the loop, so consecutive units overlap; the mask is shifted down once after the loop.  Then
            // ONE warp scan + ONE atomic per warp per tile reserve the warp's range of L1 and every thread appends
            // its own few survivors (L1's order is free)
            uint32_t umask[KPT];
#pragma unroll
            for (int j = 0; j < KPT; ++j) umask[j] = 0;
            auto unit_loop = [&](auto small_k) {
                constexpr bool SMALLK = decltype(small_k)::value;
                const uint8_t* desc = st + kTileDescOff;
#pragma unroll 2
                for (uint32_t u = 0; u < head.x; ++u, desc += 48) {
#pragma unroll
                    for (int j = 0; j < KPT; ++j) {
                        // a key slot of a kind this tile does not carry reads an absent-filter record (see
                        // TileFilter) and is masked out after the loop
                        const uint4 f = *reinterpret_cast<const uint4*>(desc + koff[j]);
                        // absent filter (k == 0): a one-bit filter whose word is all ones (TileRec::ones), so the
                        // tests pass by themselves — cannot disqualify (query_exec.go:137-151) — and round B1 sets the bit
                        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(st + tile_rel(f.w));
                        uint32_t bit[NT], wv[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) bit[t] = mod_m32(loc[j][t], f.x, f.y, f.z);
#pragma unroll
                        for (int t = 0; t < NT; ++t) wv[t] = w32[word_index(bit[t])];
                        uint32_t p = 0xffffffffu;
#pragma unroll
                        for (int t = 0; t < NT; ++t) {
                            uint32_t b = wv[t] >> (bit[t] & 31u);
                            if (SMALLK) b |= static_cast<uint32_t>(tile_k(f.w) <= static_cast<uint32_t>(t));
                            p &= b;
                        }
                        umask[j] = __funnelshift_r(umask[j], p, 1);   // (umask >> 1) | (p << 31)
                    }
                }
            };
            if (head.z & kTileSmallK) unit_loop(std::true_type{});
            else unit_loop(std::false_type{});
            const uint32_t down = 32u - head.x;   // unit u entered u-th: it sits at bit 32 - n_units + u
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                umask[j] >>= down;
                if ((kbit[j] & head.y) == 0) umask[j] = 0;   // not a key of this tile's kinds (or no key at all)
                mine += __popc(umask[j]);
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= static_cast<uint32_t>(d)) incl += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt[0], total);
                uint32_t off = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
#pragma unroll
                for (int j = 0; j < KPT; ++j) {
                    uint32_t mk = umask[j];
                    while (mk) {
                        const uint32_t u = static_cast<uint32_t>(__ffs(static_cast<int>(mk))) - 1u;
                        mk &= mk - 1u;
                        L1[off++] = static_cast<uint16_t>((u << 10) | (tid + j * NTHR));
                    }
                }
            }
        }
        __syncthreads();
        const uint32_t n1 = ld_volatile_shared_u32(&cnt[0]);
        if (TRACE && tr && tid == 0 && 7 + 8 * n < a.trace_slots) { tr[3 + 8 * n] = globaltimer_ns(); tr[6 + 8 * n] = n1; }
        // --------------------------------------------------------------- round B1 ---
        for (uint32_t e0 = warp * 32; e0 < n1; e0 += blockDim.x) {
            const uint32_t e = e0 + lane;
            bool more = false;
            uint32_t entry = 0;
            if (e < n1) {
                entry = L1[e];
                const uint32_t u = entry >> 10, slot = entry & 0x3ffu;
                const uint32_t si = s_slot[slot];
                const uint4 f = *reinterpret_cast<const uint4*>(st + kTileDescOff + u * 48u + (si >> 14) * 16u);
                const uint32_t k = tile_k(f.w);
                bool fin = k <= static_cast<uint32_t>(NT);   // absent filter (k == 0), or every location already passed
                if (!fin) {
                    const ulonglong2 x = htab[2 * slot], y = htab[2 * slot + 1];
                    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(st + tile_rel(f.w));
                    uint32_t ok = 1u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {   // locations NT..NT+3: i%4 == 0: h0+i*h2, 1: h1+i*h3, 2: h0+i*h3, 3: h1+i*h2
                        constexpr int dummy = 0;
                        (void)dummy;
                        const int i = NT + j;
                        const uint64_t aa = (i & 1) ? x.y : x.x;
                        const uint64_t bb = (((i + (i & 1)) & 3) >> 1) ? y.y : y.x;
                        const uint32_t bit = mod_m32(aa + static_cast<uint64_t>(i) * bb, f.x, f.y, f.z);
                        ok &= ((w32[word_index(bit)] >> (bit & 31u)) & 1u) | static_cast<uint32_t>(static_cast<uint32_t>(i) >= k);
                    }
                    fin = ok != 0u && k <= static_cast<uint32_t>(NT + 4);
                    more = ok != 0u && !fin;
                }
                if (fin) {
                    const uint32_t pos = si & 0x3ffu;
                    atomicOr(&rows[u * 32u + (pos >> 5)], 1u << (pos & 31u));
                }
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, more);
            if (bits) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&cnt[1], static_cast<uint32_t>(__popc(bits)));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (more) L2[base + __popc(bits & lt_mask)] = static_cast<uint16_t>(entry);
            }
        }
        __syncthreads();
        if (tid == 0) cnt[0] = 0;   // L1 is consumed; the next tile's round A starts after the third barrier
        const uint32_t n2 = ld_volatile_shared_u32(&cnt[1]);
        if (TRACE && tr && tid == 0 && 7 + 8 * n < a.trace_slots) { tr[4 + 8 * n] = globaltimer_ns(); tr[7 + 8 * n] = n2; }
        // --------------------------------------------------------------- round B2 ---
        for (uint32_t e = tid; e < n2; e += blockDim.x) {
            const uint32_t entry = L2[e];
            const uint32_t u = entry >> 10, slot = entry & 0x3ffu;
            const uint32_t si = s_slot[slot];
            const uint4 f = *reinterpret_cast<const uint4*>(st + kTileDescOff + u * 48u + (si >> 14) * 16u);
            const ulonglong2 x = htab[2 * slot], y = htab[2 * slot + 1];
            if (test_from_s32<NT + 4>(x.x, x.y, y.x, y.y, f.x, f.y, f.z, tile_k(f.w),
                                      reinterpret_cast<const uint32_t*>(st + tile_rel(f.w)))) {
                const uint32_t pos = si & 0x3ffu;
                atomicOr(&rows[u * 32u + (pos >> 5)], 1u << (pos & 31u));
            }
        }
        __syncthreads();
        if (TRACE && tr && tid == 0 && 5 + 8 * n < a.trace_slots) tr[5 + 8 * n] = globaltimer_ns();
        // ---- rows out (one coalesced store per unit), stage refill ----
        if ((head.z & kTileLastPart) && warp < head.x) {
            const uint32_t v = rows[warp * 32u + lane];
            if (lane < out_words) out_base[static_cast<size_t>(my_unit) * a.row_words32 + lane] = v;
            rows[warp * 32u + lane] = 0;
        }
        if (tid == 0) {
            cnt[1] = 0;
            const uint32_t nxt = n + S;
            if (nxt < my_tiles) {
                const uint4* fp = reinterpret_cast<const uint4*>(st + kTileNextFillOff);
                const uint4 f0 = fp[0];
                uint16_t nb16[kTileMaxUnits * 3] = {};
                if (a.kind_mask != 7u) {
                    const uint16_t* src16 = reinterpret_cast<const uint16_t*>(st + kTileNextFillOff + 16);
#pragma unroll
                    for (int i = 0; i < static_cast<int>(kTileMaxUnits) * 3; ++i) nb16[i] = src16[i];
                }
                const bool has_next = nxt + S < my_tiles;
                // no proxy fence: the stage was only READ through the generic proxy, and those reads are ordered
                // before this point by the CTA barrier (write-after-read needs no cross-proxy fence)
                fill_tile(st, &full[s], rec_of(nxt), has_next, has_next ? rec_of(nxt + S) : rec_of(nxt), a.words, f0, nb16,
                          a.kind_mask, a.hdr_bytes);
            }
        }
        st += a.stage_bytes;
        if (++s == S) { s = 0; ph ^= 1u; st = stages; }
    }
}

// ---- compiled shapes: <tests of round A, threads per CTA> (1024 / threads CTAs share an SM) ----
// (a software-pipelined variant of the same rounds — mbarrier hand-offs instead of CTA barriers, double-buffered
// lists, ring of >= 5 stages — measured 16 % slower on both layouts and was removed; profiles/r02_kernel_experiments.txt)
#define BSG_TILES_SHAPES(X) X(0, 3, 512) X(1, 3, 1024) X(2, 2, 512) X(3, 2, 1024) X(4, 4, 512) X(5, 4, 1024)

int probe_tiles_n_shapes() {
    int n = 0;
#define X(id, nt, thr) ++n;
    BSG_TILES_SHAPES(X)
#undef X
    return n;
}
int probe_tiles_threads(int shape) {
    switch (shape) {
#define X(id, nt, thr) case id: return thr;
        BSG_TILES_SHAPES(X)
#undef X
        default: return 1024;
    }
}
const char* probe_tiles_shape_name(int shape) {
    switch (shape) {
#define X(id, nt, thr) case id: return "probe_tiles_kernel<NT=" #nt "," #thr " threads>";
        BSG_TILES_SHAPES(X)
#undef X
        default: return "probe_tiles_kernel<?>";
    }
}

cudaError_t probe_tiles_configure(int max_smem_optin) {
    cudaError_t e = cudaSuccess;
#define X(id, nt, thr)                                                                                              \
    e = cudaFuncSetAttribute(probe_tiles_kernel<nt, thr, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin); \
    if (e != cudaSuccess) return e;                                                                                       \
    e = cudaFuncSetAttribute(probe_tiles_kernel<nt, thr, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);  \
    if (e != cudaSuccess) return e;
    BSG_TILES_SHAPES(X)
#undef X
    return e;
}

template <int NT, int NTHR>
static cudaError_t tiles_launch(const ProbeTilesPlan& plan, const ProbeTilesArgs& args, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3(NTHR);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = plan.pdl ? 1 : 0;
    if (args.trace) return cudaLaunchKernelEx(&cfg, probe_tiles_kernel<NT, NTHR, true>, args);
    return cudaLaunchKernelEx(&cfg, probe_tiles_kernel<NT, NTHR, false>, args);
}

cudaError_t launch_probe_tiles(const ProbeTilesPlan& plan, const TileRec* d_tiles, uint32_t n_items,
                               const uint32_t* d_n_items, const uint64_t* d_words, const uint64_t* d_hashes,
                               const uint16_t* d_slotinfo, uint32_t key_base, uint32_t n_keys, uint32_t kind_mask,
                               uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s, uint64_t* d_trace,
                               uint32_t trace_slots) {
    if ((n_items == 0 && !d_n_items) || n_keys == 0) return cudaSuccess;
    if (n_keys > kProbeMaxKeysPerPass) return cudaErrorInvalidValue;
    ProbeTilesArgs a;
    a.tiles = d_tiles;
    a.n_items_dev = d_n_items;
    a.n_items_host = n_items;
    a.parts = plan.parts;
    a.words = d_words;
    a.hashes = plan.fuse_keys ? nullptr : d_hashes;
    a.key_bytes = plan.fuse_keys;
    a.key_off = plan.fuse_key_off;
    a.slotinfo = d_slotinfo;
    a.key_base = key_base;
    a.n_keys = n_keys;
    a.kind_mask = kind_mask;
    a.matrix32 = d_matrix32;
    a.row_words32 = row_words32;
    a.n_stages = static_cast<uint32_t>(plan.n_stages);
    a.hdr_bytes = tile_header_bytes(plan.units_cap);
    a.stage_bytes = a.hdr_bytes + plan.stage_data_bytes;
    a.units_cap = plan.units_cap;
    a.trace = d_trace;
    a.trace_slots = d_trace ? trace_slots : 0;
    switch (plan.shape) {
#define X(id, nt, thr) case id: return tiles_launch<nt, thr>(plan, a, s);
        BSG_TILES_SHAPES(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
}

// Stream compaction between the two stages of a hierarchical probe: keeps the items with at least one
// unit whose parent (file) survived.  Order is not preserved (no consumer needs it).
__global__ void __launch_bounds__(256)
compact_tiles_kernel(const TileRec* __restrict__ tiles, uint32_t n_items, uint32_t parts,
                     const uint32_t* __restrict__ parent, const uint32_t* __restrict__ parent_mask32,
                     TileRec* __restrict__ out, uint32_t* __restrict__ n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n_items) {
        const TileRec* r = tiles + static_cast<size_t>(i) * parts;
        const uint32_t nu = r->n_units;
        for (uint32_t u = 0; u < nu && !keep; ++u) {
            const uint32_t f = __ldg(&parent[r->unit[u]]);
            keep = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
        }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && bal) base = atomicAdd(n_out, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const uint4* src = reinterpret_cast<const uint4*>(tiles + static_cast<size_t>(i) * parts);
        uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(base + __popc(bal & ((1u << lane) - 1u))) * parts);
        const uint32_t n16 = parts * (sizeof(TileRec) / 16);
        for (uint32_t j = 0; j < n16; ++j) dst[j] = __ldg(src + j);
    }
}

cudaError_t launch_compact_tiles(const TileRec* d_tiles, uint32_t n_items, uint32_t parts, const uint32_t* d_parent,
                                 const uint32_t* d_parent_mask32, TileRec* d_out, uint32_t* d_n_out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_n_out, 0, 4, s);
    if (e != cudaSuccess || n_items == 0) return e;
    compact_tiles_kernel<<<(n_items + 255) / 256, 256, 0, s>>>(d_tiles, n_items, parts, d_parent, d_parent_mask32, d_out,
                                                              d_n_out);
    return cudaGetLastError();
}

}  // namespace bsg
