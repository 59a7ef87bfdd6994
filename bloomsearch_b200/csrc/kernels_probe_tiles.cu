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
//   * phase A publishes survivors as BALLOT words (one store per 32 keys, no atomics, no queue); the
//     phase-B team that owns the unit expands the 1024-bit survivor bitmap into one dense list (each
//     warp of the team expands its share of the words), tests locations NT..k-1 with dense lanes in
//     groups of four independent tests, ORs passing keys into the team's result row (caller key
//     order) and writes the row with one coalesced 128-byte store.
//
//   phase A  warps 0..NA-1, KPT keys per thread, locations 0..NT-1 of each key in registers; walks the
//            tiles in order; per unit of the tile: NT branch-free tests per key, one ballot per 32 keys.
//   phase B  warps NA..NA+NB-1 in NB/T teams of T warps; team g owns the units whose ordinal in this CTA's
//            sequence is g mod NB/T (all parts of a unit go to the same team, so the row is complete when
//            its last part is).  Two named barriers per task order list build -> tests -> row write-out; the
//            survivor counter only grows, so nothing is reset between tasks.  Every B warp visits every
//            tile and arrives on its done counter (so no warp can fall two mbarrier phases behind a
//            stage); the last arriver clears the bitmaps and refills the stage.  (Measured: one warp per
//            unit, 6-8 B warps, left phase B latency bound: 2b 25 us, 2a 68-80 us; see DESIGN.md.)
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
// f = {m, ih, il, k << 16 | rel16}; SMALLK: some filter of the tile has k < 4, so location t exists
// only when t < k.
template <int NT, bool SMALLK>
__device__ __forceinline__ bool first_tests(const uint64_t (&loc)[NT], const uint4 f, const uint8_t* data) {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(data + ((f.w & 0xffffu) << 4));
    uint32_t bit[NT], wv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) bit[t] = mod_m32(loc[t], f.x, f.y, f.z);
#pragma unroll
    for (int t = 0; t < NT; ++t) wv[t] = w32[bit[t] >> 5];
    uint32_t pass = 1u;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        uint32_t b = (wv[t] >> (bit[t] & 31u)) & 1u;
        if (SMALLK) b |= static_cast<uint32_t>((f.w >> 16) <= static_cast<uint32_t>(t));
        pass &= b;
    }
    return pass != 0u;
}

__device__ __forceinline__ void team_barrier(uint32_t id, uint32_t n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

template <int NA, int KPT, int NT, int NB, int T, bool TRACE>
__global__ void __launch_bounds__((NA + NB) * 32, 1) probe_tiles_kernel(const ProbeTilesArgs a) {
    static_assert(NA * KPT * 32 == static_cast<int>(kProbeMaxKeysPerPass), "A warps must cover one pass of keys");
    static_assert(NA + NB <= 32 && NT >= 1 && NT <= 4 && NB % T == 0 && 32 % T == 0 && (T == 1 || NB / T <= 15), "shape");
    constexpr uint32_t NTEAMS = NB / T;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* aready = full + kProbeMaxStages;
    uint32_t* done = reinterpret_cast<uint32_t*>(aready + kProbeMaxStages);
    uint16_t* s_slot = reinterpret_cast<uint16_t*>(smem + kTilesPrefixBytes);
    uint8_t* bwarp_area = smem + kTilesPrefixBytes + kTilesSlotInfoBytes;
    ulonglong2* htab = reinterpret_cast<ulonglong2*>(bwarp_area + NTEAMS * kTilesPerTeamBytes);
    const uint32_t hash_bytes = ((a.n_keys + 31u) & ~31u) * 32u;
    uint8_t* stages = reinterpret_cast<uint8_t*>(htab) + hash_bytes;   // 128-byte aligned: every term is

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t warp = tid >> 5;
    const uint32_t G = gridDim.x;
    const uint32_t S = a.n_stages;
    const uint32_t P = a.parts;
    // optional timeline (profiling only), per CTA: [0] start, [1] hashes ready; per tile n, base b = 2+8n:
    // [b] resident (A warp 0), [b+1] A warp 0 done, [b+2] all A done (a B task starts), [b+3] released,
    // [b+4] list expanded, [b+5] after team barrier 1, [b+6] tests done, [b+7] after team barrier 2 (member 0 of a team)
    uint64_t* tr = (TRACE && a.trace) ? a.trace + static_cast<size_t>(blockIdx.x) * a.trace_slots : nullptr;
    if (TRACE && tr && tid == 0) tr[0] = globaltimer_ns();

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&aready[s], NA);
            done[s] = 0;
        }
        fence_barrier_init();
    }
    for (uint32_t i = tid; i < NTEAMS * 64u; i += blockDim.x)  // the teams' result rows start at zero
        reinterpret_cast<uint32_t*>(bwarp_area + (i / 64u) * kTilesPerTeamBytes)[i % 64u] = 0;
    for (uint32_t i = tid; i < S * a.units_cap * 32u; i += blockDim.x) {  // survivor bitmaps start clear
        const uint32_t s = i / (a.units_cap * 32u), r = i % (a.units_cap * 32u);
        reinterpret_cast<uint32_t*>(stages + static_cast<size_t>(s) * a.stage_bytes + kTileBitmapOff)[r] = 0;
    }
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

    uint32_t s = 0, ph = 0;
    uint8_t* st = stages;
    if (warp < NA) {
        // ------------------------------------------------------------------ phase A ---
        uint64_t loc[KPT][NT];   // location(h, 0..NT-1) = h0, h1+h3, h0+2*h3, h1+3*h2
        uint32_t koff[KPT];      // byte offset of the key's kind inside a unit's descriptor triple
        uint32_t kbit[KPT];      // 1 << kind, 0 = no key in this slot
        uint32_t warp_kinds = 0;
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const uint32_t slot = (warp * KPT + j) * 32 + lane;
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
        for (uint32_t n = 0; n < my_tiles; ++n) {
            mbar_wait_relaxed(&full[s], ph, 128u, 0u);
            if (TRACE && tr && tid == 0 && 2 + 8 * n < a.trace_slots) tr[2 + 8 * n] = globaltimer_ns();
            const uint4 head = *reinterpret_cast<const uint4*>(st);  // n_units, part_kinds, flags
            if (warp_kinds & head.y) {
                const uint8_t* data = st + a.hdr_bytes;
                const uint8_t* desc = st + kTileDescOff;
                // Ballot words stay in registers until the tile is done (lane u*KPT+j keeps the word of unit u,
                // key slot j): no shared-memory store inside the loop, so the loads and modulo chains of
                // consecutive keys / units can overlap (phase A is latency bound, not issue bound).
                uint32_t mybits = 0;
                auto unit_loop = [&](auto small_k) {
                    constexpr bool SMALLK = decltype(small_k)::value;
#pragma unroll 2
                    for (uint32_t u = 0; u < head.x; ++u, desc += 48) {
                        uint4 f[KPT];
                        bool act[KPT], pass[KPT];
#pragma unroll
                        for (int j = 0; j < KPT; ++j) {
                            act[j] = (kbit[j] & head.y) != 0;
                            f[j] = make_uint4(0, 0, 0, 0);
                            if (act[j]) f[j] = *reinterpret_cast<const uint4*>(desc + koff[j]);
                        }
#pragma unroll
                        for (int j = 0; j < KPT; ++j) {
                            // absent filter (m == 0): cannot disqualify (query_exec.go:137-151) -> phase B sets the bit
                            const bool test = act[j] && f[j].x != 0;
                            const uint32_t* w32 = reinterpret_cast<const uint32_t*>(data + ((f[j].w & 0xffffu) << 4));
                            uint32_t bit[NT], wv[NT];
#pragma unroll
                            for (int t = 0; t < NT; ++t) bit[t] = mod_m32(loc[j][t], f[j].x, f[j].y, f[j].z);
#pragma unroll
                            for (int t = 0; t < NT; ++t) wv[t] = test ? w32[bit[t] >> 5] : 0u;
                            uint32_t p = 1u;
#pragma unroll
                            for (int t = 0; t < NT; ++t) {
                                uint32_t b = (wv[t] >> (bit[t] & 31u)) & 1u;
                                if (SMALLK) b |= static_cast<uint32_t>((f[j].w >> 16) <= static_cast<uint32_t>(t));
                                p &= b;
                            }
                            pass[j] = act[j] && (f[j].x == 0 || p != 0u);
                        }
#pragma unroll
                        for (int j = 0; j < KPT; ++j) {
                            const uint32_t bits = __ballot_sync(0xffffffffu, pass[j]);
                            if (lane == u * KPT + j) mybits = bits;
                        }
                    }
                };
                if (head.z & kTileSmallK) unit_loop(std::true_type{});
                else unit_loop(std::false_type{});
                // lane 0 publishes the words (it is also the lane that arrives on the barrier below)
                uint32_t* bm = reinterpret_cast<uint32_t*>(st + kTileBitmapOff) + warp * KPT;
                for (uint32_t i = 0; i < head.x * KPT; ++i) {
                    const uint32_t v = __shfl_sync(0xffffffffu, mybits, i);
                    if (lane == 0) bm[(i / KPT) * 32 + (i % KPT)] = v;
                }
            }
            // lane 0 wrote this warp's bitmap words and lane 0 arrives (release): a B warp's try_wait
            // (acquire) on aready orders its loads after them
            if (lane == 0) mbar_arrive(&aready[s]);
            __syncwarp();
            if (TRACE && tr && tid == 0 && 3 + 8 * n < a.trace_slots) tr[3 + 8 * n] = globaltimer_ns();
            st += a.stage_bytes;
            if (++s == S) { s = 0; ph ^= 1u; st = stages; }
        }
    } else {
        // ------------------------------------------------------------------ phase B ---
        // Per-lane state machines over the survivor bitmap: lane l of the team's member j owns bits
        // [j*32/T, (j+1)*32/T) of bitmap word l.  A lane pulls its next survivor as soon as its current one
        // is decided, so the geometric tail of TestString (half of the remaining absent keys die at every
        // location) does not idle the warp: no list, no scan, no atomics, no barrier inside a task.
        const uint32_t wb = warp - NA;
        const uint32_t team = wb / T, member = wb % T;
        uint32_t* rows = reinterpret_cast<uint32_t*>(bwarp_area + team * kTilesPerTeamBytes);  // two rows, by unit parity
        uint32_t units_done = 0;
        const uint32_t out_words = (a.n_keys + 31) >> 5;
        uint32_t* out_base = a.matrix32 + (a.key_base >> 5);
        constexpr uint32_t BPM = 32 / T;                                // bits of every bitmap word a member owns
        const uint32_t my_bits = (T == 1 ? 0xffffffffu : ((1u << BPM) - 1u)) << (member * BPM);
        for (uint32_t n = 0; n < my_tiles; ++n) {
            mbar_wait_relaxed(&full[s], ph, 128u, 0u);  // the bulk copies' bytes (async proxy) are visible
            const uint4 head = *reinterpret_cast<const uint4*>(st);
            bool waited = false;
            for (uint32_t u = 0; u < head.x; ++u) {
                // ordinal of the unit in this CTA's sequence; all parts of a unit share it
                const uint32_t ord = P == 1 ? n * a.units_cap + u : (n >> 1);
                if (ord % NTEAMS != team) continue;
                if (!waited) {  // every A warp has published its survivor words for this tile
                    mbar_wait_relaxed(&aready[s], ph, 256u, 0u);
                    waited = true;
                    if (TRACE && tr && lane == 0 && member == 0 && 4 + 8 * n < a.trace_slots) tr[4 + 8 * n] = globaltimer_ns();
                }
                uint32_t* row = rows + (units_done & 1u) * 32u;
                uint32_t w = ld_volatile_shared_u32(st + kTileBitmapOff + u * 128u + lane * 4u) & my_bits;
                const uint8_t* desc = st + kTileDescOff + u * 48u;
                const uint8_t* data = st + a.hdr_bytes;
                bool alive = false;
                uint64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0, ih2 = 0, ih3 = 0;   // ih2 = i*h2, ih3 = i*h3
                uint32_t fm = 1, fih = 0, fil = 0, fk = 0, i = 0, pos = 0;
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(data);
                while (__any_sync(0xffffffffu, alive || w != 0u)) {
                    if (!alive && w != 0u) {                              // next survivor of this lane
                        const uint32_t b = __ffs(w) - 1;
                        w &= w - 1;
                        const uint32_t slot = lane * 32 + b;
                        const uint32_t si = s_slot[slot];
                        pos = si & 0x3ffu;
                        const uint4 f = *reinterpret_cast<const uint4*>(desc + (si >> 14) * 16u);
                        if (f.x == 0 || (f.w >> 16) <= static_cast<uint32_t>(NT)) {
                            // absent filter cannot disqualify (query_exec.go:137-151); k <= NT: every location passed
                            atomicOr(&row[pos >> 5], 1u << (pos & 31u));
                        } else {
                            const ulonglong2 x = htab[2 * slot], y = htab[2 * slot + 1];
                            h0 = x.x; h1 = x.y; h2 = y.x; h3 = y.y;
                            ih2 = static_cast<uint64_t>(NT) * h2;
                            ih3 = static_cast<uint64_t>(NT) * h3;
                            fm = f.x; fih = f.y; fil = f.z; fk = f.w >> 16;
                            w32 = reinterpret_cast<const uint32_t*>(data + ((f.w & 0xffffu) << 4));
                            i = NT;
                            alive = true;
                        }
                    }
                    if (alive) {
                        // two locations per step (independent chains): i and i+1; location(h,i) = h[i&1] + i*h[2 + ...]
                        // i%4 == 0: h0+i*h2, 1: h1+i*h3, 2: h0+i*h3, 3: h1+i*h2
                        const uint64_t a0 = (i & 1u) ? h1 : h0, a1 = (i & 1u) ? h0 : h1;
                        const bool s0 = (((i + (i & 1u)) & 3u) >> 1) != 0u;              // true: h3
                        const bool s1 = ((((i + 1u) + ((i + 1u) & 1u)) & 3u) >> 1) != 0u;
                        const uint64_t l0 = a0 + (s0 ? ih3 : ih2);
                        const uint64_t l1 = a1 + (s1 ? ih3 + h3 : ih2 + h2);
                        const uint32_t b0 = mod_m32(l0, fm, fih, fil), b1 = mod_m32(l1, fm, fih, fil);
                        const uint32_t v0 = w32[b0 >> 5], v1 = w32[b1 >> 5];
                        const uint32_t ok0 = (v0 >> (b0 & 31u)) & 1u;
                        const uint32_t ok1 = ((v1 >> (b1 & 31u)) & 1u) | static_cast<uint32_t>(i + 1u >= fk);
                        i += 2u;
                        ih2 += 2 * h2;
                        ih3 += 2 * h3;
                        if (!(ok0 & ok1)) {
                            alive = false;
                        } else if (i >= fk) {
                            atomicOr(&row[pos >> 5], 1u << (pos & 31u));
                            alive = false;
                        }
                    }
                }
                if (head.z & kTileLastPart) {                            // the unit's row is complete: one coalesced store
                    if (T > 1) team_barrier(1 + team, T * 32); else __syncwarp();
                    if (member == 0) {
                        const uint32_t unit = *reinterpret_cast<const uint32_t*>(st + 16 + 64 + 4 * u);
                        const uint32_t v = ld_volatile_shared_u32(&row[lane]);
                        if (lane < out_words) out_base[static_cast<size_t>(unit) * a.row_words32 + lane] = v;
                        row[lane] = 0;   // reused two units later, after a team barrier this warp joins later
                    }
                    ++units_done;
                    if (TRACE && tr && lane == 0 && member == 0 && 6 + 8 * n < a.trace_slots) tr[6 + 8 * n] = globaltimer_ns();
                }
            }
            // every B warp arrives for every tile; the last one clears the bitmaps and refills the stage
            __syncwarp();
            uint32_t last = 0;
            if (lane == 0) last = atom_add_acq_rel_shared(&done[s], 1u) == NB - 1;
            __syncwarp();
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                for (uint32_t u = 0; u < head.x; ++u)
                    reinterpret_cast<uint32_t*>(st + kTileBitmapOff)[u * 32 + lane] = 0;
                __syncwarp();
                if (lane == 0) {
                    if (TRACE && tr && 5 + 8 * n < a.trace_slots) tr[5 + 8 * n] = globaltimer_ns();
                    done[s] = 0;
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
                        fence_proxy_async();
                        fill_tile(st, &full[s], rec_of(nxt), has_next, has_next ? rec_of(nxt + S) : rec_of(nxt), a.words,
                                  f0, nb16, a.kind_mask, a.hdr_bytes);
                    }
                }
                __syncwarp();
            }
            st += a.stage_bytes;
            if (++s == S) { s = 0; ph ^= 1u; st = stages; }
        }
    }
}

// ---- compiled shapes: <A warps, keys per A thread, A tests, B warps, B team size> ----
#define BSG_TILES_SHAPES(X) \
    X(0, 16, 2, 3, 16, 1) X(1, 16, 2, 2, 16, 1) X(2, 16, 2, 3, 8, 1) X(3, 16, 2, 3, 16, 2) X(4, 16, 2, 3, 16, 4) \
    X(5, 16, 2, 3, 8, 4) X(6, 8, 4, 3, 16, 1) X(7, 16, 2, 2, 16, 4)

int probe_tiles_n_shapes() {
    int n = 0;
#define X(id, na, kpt, nt, nb, t) ++n;
    BSG_TILES_SHAPES(X)
#undef X
    return n;
}
int probe_tiles_teams(int shape) {
    switch (shape) {
#define X(id, na, kpt, nt, nb, t) case id: return nb / t;
        BSG_TILES_SHAPES(X)
#undef X
        default: return 4;
    }
}
const char* probe_tiles_shape_name(int shape) {
    switch (shape) {
#define X(id, na, kpt, nt, nb, t) case id: return "probe_tiles_kernel<" #na "," #kpt "," #nt "," #nb "," #t ">";
        BSG_TILES_SHAPES(X)
#undef X
        default: return "probe_tiles_kernel<?>";
    }
}

cudaError_t probe_tiles_configure(int max_smem_optin) {
    cudaError_t e = cudaSuccess;
#define X(id, na, kpt, nt, nb, t)                                                                                  \
    e = cudaFuncSetAttribute(probe_tiles_kernel<na, kpt, nt, nb, t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             max_smem_optin);                                                                      \
    if (e != cudaSuccess) return e;                                                                                \
    e = cudaFuncSetAttribute(probe_tiles_kernel<na, kpt, nt, nb, t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             max_smem_optin);                                                                      \
    if (e != cudaSuccess) return e;
    BSG_TILES_SHAPES(X)
#undef X
    return e;
}

template <int NA, int KPT, int NT, int NB, int T>
static cudaError_t tiles_launch(const ProbeTilesPlan& plan, const ProbeTilesArgs& args, cudaStream_t s) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.grid);
    cfg.blockDim = dim3((NA + NB) * 32);
    cfg.dynamicSmemBytes = plan.smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = plan.pdl ? 1 : 0;
    if (args.trace) return cudaLaunchKernelEx(&cfg, probe_tiles_kernel<NA, KPT, NT, NB, T, true>, args);
    return cudaLaunchKernelEx(&cfg, probe_tiles_kernel<NA, KPT, NT, NB, T, false>, args);
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
#define X(id, na, kpt, nt, nb, t) case id: return tiles_launch<na, kpt, nt, nb, t>(plan, a, s);
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
