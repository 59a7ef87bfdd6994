// K4/K5 — the MaybeContains probe.
//
// Replaces the reference's per-block loop (query_exec.go:572-615): for every
// unit (data block, or file for the file-level stage query_exec.go:399-406) and
// every query key, TestString on the unit's filter of the key's kind
// (query_exec.go:128-159), i.e. k bit tests at location(h,i) % m with early exit
// on the first clear bit.  The reference re-decodes every filter section and
// re-hashes every key per block; here the corpus is resident in HBM in native
// word order and keys are hashed once per batch (kernels_hash.cu).
//
// Two data paths produce identical bits:
//   staged  — large batches: each unit's bitsets are bulk-copied (TMA 1-D, cp.async.bulk +
//             mbarrier) into a multi-stage shared-memory ring; the warp that frees a stage
//             refills it.  Every bitset byte crosses HBM once per batch: the HBM-roofline
//             regime of SURVEY.md §8(d).  probe_staged2_kernel (default) splits TestString
//             into a branch-free first phase over all keys and a second phase over the
//             compacted survivors; probe_staged_kernel is its one-phase predecessor.
//   gather  — small batches or filters too large to stage: one lane per (unit,key), bit
//             words gathered straight from L2/HBM (the sparse 8*k bytes/probe bound).
// tree_eval turns the (unit x key) bit matrix into the candidate mask with the
// query's AND/OR tree (query_exec.go:89-125) in postfix form.
#include <cstddef>

#include "bsg_device.cuh"
#include "bsg_internal.h"

namespace bsg {

// One membership test against a bitset viewed as 32-bit words.
// LOAD(idx32) returns the 32-bit word idx32 of the filter.
template <typename Load>
__device__ __forceinline__ bool test_bit(uint64_t loc, uint64_t m, uint64_t inv, Load&& load) {
    const uint64_t bit = mod_m(loc, m, inv);
    const uint32_t w = load(bit >> 5);
    return (w >> (static_cast<uint32_t>(bit) & 31u)) & 1u;
}

// TestString with precomputed base hashes: all k locations set?
// location(h,i): i%4 == 0: h0+i*h2, 1: h1+i*h3, 2: h0+i*h3, 3: h1+i*h2.
template <typename Load>
__device__ __forceinline__ bool test_hashes(const uint64_t h0, const uint64_t h1, const uint64_t h2,
                                            const uint64_t h3, uint64_t m, uint64_t inv, uint32_t k,
                                            Load&& load) {
    uint64_t ih2 = 0, ih3 = 0;  // i*h2, i*h3 at i = multiple of 4
    for (uint32_t i = 0; i < k; i += 4) {
        if (!test_bit(h0 + ih2, m, inv, load)) return false;
        if (i + 1 >= k) break;
        if (!test_bit(h1 + ih3 + h3, m, inv, load)) return false;
        if (i + 2 >= k) break;
        if (!test_bit(h0 + ih3 + 2 * h3, m, inv, load)) return false;
        if (i + 3 >= k) break;
        if (!test_bit(h1 + ih2 + 3 * h2, m, inv, load)) return false;
        ih2 += 4 * h2;
        ih3 += 4 * h3;
    }
    return true;
}

// ------------------------------------------------------------ staged path ---
// Shared-memory map: [full mbarriers x16][done counters x16][pad] | stage 0 | stage 1 ...
// stage = [StageRow of the unit, 128 B][head of the unit that will occupy this stage next, 32 B][unit words]
// There is no producer warp: the warp whose release makes a stage free (last of n_warps to
// bump done[s]) refills it at once, using the next unit's head that travelled in with the
// current unit, so a refill never waits on a global load.

__device__ __forceinline__ void fill_stage(uint8_t* st, uint64_t* full_bar, const StageRow* __restrict__ stab,
                                           uint32_t li, bool has_next, uint32_t li_next,
                                           const uint64_t* __restrict__ words, uint64_t word_base, uint32_t nw0,
                                           uint32_t nw1, uint32_t nw2, uint32_t kind_mask,
                                           uint32_t hdr_bytes = kProbeStageHeaderBytes) {
    const uint32_t b0 = (kind_mask & 1u) ? nw0 * 8u : 0u;
    const uint32_t b1 = (kind_mask & 2u) ? nw1 * 8u : 0u;
    const uint32_t b2 = (kind_mask & 4u) ? nw2 * 8u : 0u;
    mbar_arrive_expect_tx(full_bar, kStageRowBytes + (has_next ? kStageHeadBytes : 0u) + b0 + b1 + b2);
    bulk_g2s(st, &stab[li], kStageRowBytes, full_bar);
    if (has_next) bulk_g2s(st + kStageRowBytes, &stab[li_next], kStageHeadBytes, full_bar);
    uint8_t* data = st + hdr_bytes;
    const uint64_t* src = words + word_base;
    if (kind_mask == 7u) {
        if (b0 + b1 + b2) bulk_g2s(data, src, b0 + b1 + b2, full_bar);
    } else {
        if (b0) bulk_g2s(data, src, b0, full_bar);
        if (b1) bulk_g2s(data + nw0 * 8u, src + nw0, b1, full_bar);
        if (b2) bulk_g2s(data + (nw0 + nw1) * 8u, src + nw0 + nw1, b2, full_bar);
    }
}

// Same as test_hashes with the 32-bit modulo and the bitset read from global memory
// (gather path, m < 2^30).
__device__ __forceinline__ bool test_hashes_g32(uint64_t h0, uint64_t h1, uint64_t h2, uint64_t h3, uint32_t m,
                                                uint32_t ih, uint32_t il, uint32_t k,
                                                const uint32_t* __restrict__ w32) {
    auto test = [&](uint64_t loc) {
        const uint32_t bit = mod_m32(loc, m, ih, il);
        return (__ldg(w32 + (bit >> 5)) >> (bit & 31u)) & 1u;
    };
    uint64_t ih2 = 0, ih3 = 0;
    for (uint32_t i = 0; i < k; i += 4) {
        if (!test(h0 + ih2)) return false;
        if (i + 1 >= k) break;
        if (!test(h1 + ih3 + h3)) return false;
        if (i + 2 >= k) break;
        if (!test(h0 + ih3 + 2 * h3)) return false;
        if (i + 3 >= k) break;
        if (!test(h1 + ih2 + 3 * h2)) return false;
        ih2 += 4 * h2;
        ih3 += 4 * h3;
    }
    return true;
}

// TestString with precomputed base hashes on a bitset resident in shared memory, m < 2^30.
// Compile-time unrolled in groups of four locations (i%4 pattern of location()), early exit on
// the first clear bit exactly like BloomFilter.Test.  The first four locations of a key do not
// depend on the unit, so they are computed once per key (l0..l3) and cost no adds per unit.
struct KeyLocs {
    uint64_t l0, l1, l2, l3;  // location(h, 0..3) = h0, h1+h3, h0+2*h3, h1+3*h2
    uint64_t h0, h1, h2, h3;
};

__device__ __forceinline__ bool test_hashes_s32(const KeyLocs& K, uint32_t m, uint32_t ih, uint32_t il, uint32_t k,
                                                const uint32_t* __restrict__ w32) {
    auto test = [&](uint64_t loc) {
        const uint32_t bit = mod_m32(loc, m, ih, il);
        return (w32[bit >> 5] & (1u << (bit & 31u))) != 0u;
    };
    if (k >= 4) {  // the common case (default fpr: k = 10/11)
        if (!test(K.l0)) return false;
        if (!test(K.l1)) return false;
        if (!test(K.l2)) return false;
        if (!test(K.l3)) return false;
        uint64_t ih2 = 4 * K.h2, ih3 = 4 * K.h3;  // i*h2, i*h3 at i = 4
        uint32_t i = 4;
        for (; i + 4 <= k; i += 4) {  // full groups: no per-test bound check
            if (!test(K.h0 + ih2)) return false;
            if (!test(K.h1 + ih3 + K.h3)) return false;
            if (!test(K.h0 + ih3 + 2 * K.h3)) return false;
            if (!test(K.h1 + ih2 + 3 * K.h2)) return false;
            ih2 += 4 * K.h2;
            ih3 += 4 * K.h3;
        }
        if (i < k && !test(K.h0 + ih2)) return false;
        if (i + 1 < k && !test(K.h1 + ih3 + K.h3)) return false;
        if (i + 2 < k && !test(K.h0 + ih3 + 2 * K.h3)) return false;
        return true;
    }
    if (k > 0 && !test(K.l0)) return false;
    if (k > 1 && !test(K.l1)) return false;
    if (k > 2 && !test(K.l2)) return false;
    return true;
}

// The staged probe.  All warps of the CTA work on the same resident unit; thread t owns key
// key_base + t for the whole kernel (hashes in registers) and tests it against every unit this
// CTA streams through its shared-memory ring.  No producer warp: the warp whose release frees
// a stage refills it at once (fill_stage), using the next unit's row head that travelled in
// with the current unit.  (Measured alternatives — per-lane state machines that decouple lanes
// across units or across several keys per lane — executed fewer iterations but 2.5x more
// instructions per iteration and lost; see DESIGN.md.)
template <int MAXT, bool TRACE>
__global__ void __launch_bounds__(MAXT, 1)
probe_staged_kernel(const StageRow* __restrict__ stab, uint32_t n_list_host, const uint32_t* __restrict__ n_list_dev,
                    const uint64_t* __restrict__ words,
                    const uint64_t* __restrict__ hashes, const uint8_t* __restrict__ kinds, uint32_t key_base,
                    uint32_t n_keys, uint32_t kind_mask, uint32_t* __restrict__ matrix32, uint32_t row_words32,
                    uint32_t n_stages, uint32_t stage_bytes, uint32_t stagger_ns, uint64_t* __restrict__ trace,
                    uint32_t trace_slots) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint32_t* done = reinterpret_cast<uint32_t*>(smem + kProbeMaxStages * sizeof(uint64_t));
    uint8_t* stages = smem + kProbeSmemPrefixBytes;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t warp = tid >> 5;
    const uint32_t n_warps = blockDim.x >> 5;
    // optional timeline (profiling only): per CTA [0]=start, [1+2*it]=unit it seen resident by
    // warp 0, [2+2*it]=unit it released by the last warp; globaltimer nanoseconds
    uint64_t* tr = (TRACE && trace) ? trace + static_cast<size_t>(blockIdx.x) * trace_slots : nullptr;
    if (TRACE && tr && tid == 0) tr[0] = globaltimer_ns();
    const uint32_t G = gridDim.x;
    const uint32_t S = n_stages;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            done[s] = 0;
        }
        fence_barrier_init();
    }
    __syncthreads();

    // hierarchical probes compact the surviving units on the device: their count lives there too
    const uint32_t n_list = n_list_dev ? __ldg(n_list_dev) : n_list_host;
    const uint32_t my_count = n_list > blockIdx.x ? (n_list - blockIdx.x + G - 1) / G : 0;

    // ---- prologue: lane l of warp 0 fills stage l with this CTA's l-th unit.  stagger_ns > 0
    //      delays fill l by l*stagger_ns (experiment knob BSG_PROBE_STAGGER; measured: staggering
    //      does not help, the default is 0). ----
    if (warp == 0 && lane < S && lane < my_count) {
        const uint64_t t_start = globaltimer_ns();
        const uint32_t li = blockIdx.x + lane * G;
        const uint4* hp = reinterpret_cast<const uint4*>(&stab[li]);
        const uint4 a = __ldg(hp), b = __ldg(hp + 1);
        const uint64_t word_base = (static_cast<uint64_t>(a.w) << 32) | a.z;
        const uint64_t t_go = t_start + static_cast<uint64_t>(lane) * stagger_ns;
        while (stagger_ns && globaltimer_ns() < t_go) {
        }
        fill_stage(stages + static_cast<size_t>(lane) * stage_bytes, &full[lane], stab, li, lane + S < my_count,
                   li + S * G, words, word_base, b.x, b.y, b.z, kind_mask);
    }

    // ---- this thread's key ----
    const bool valid = tid < n_keys;
    KeyLocs K = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t f_off = 0;  // byte offset of this key's StageFilter inside a stage row
    if (valid) {
        const uint32_t q = key_base + tid;
        const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hashes + 4ull * q);
        const ulonglong2 a = __ldg(hp), b = __ldg(hp + 1);
        K.h0 = a.x; K.h1 = a.y; K.h2 = b.x; K.h3 = b.y;
        K.l0 = K.h0; K.l1 = K.h1 + K.h3; K.l2 = K.h0 + 2 * K.h3; K.l3 = K.h1 + 3 * K.h2;
        f_off = 32u + 32u * __ldg(&kinds[q]);
    }
    const bool warp_has_keys = warp * 32 < n_keys;
    uint32_t* out_base = matrix32 + ((key_base + warp * 32) >> 5);

    uint32_t s = 0, ph = 0;
    const uint8_t* st = stages;
    for (uint32_t it = 0; it < my_count; ++it) {
        mbar_wait(&full[s], ph);
        if (TRACE && tr && tid == 0 && 1 + 2 * it < trace_slots) tr[1 + 2 * it] = globaltimer_ns();
        bool res = false;
        if (valid) {
            const uint4 f = *reinterpret_cast<const uint4*>(st + f_off);  // m, k, ih, il
            if (f.x == 0) {
                res = true;  // absent filter cannot disqualify (query_exec.go:137-151)
            } else {
                const uint32_t rel = *reinterpret_cast<const uint32_t*>(st + f_off + 16);
                res = test_hashes_s32(K, f.x, f.z, f.w, f.y,
                                      reinterpret_cast<const uint32_t*>(st + kProbeStageHeaderBytes + rel));
            }
        }
        const uint32_t bits = __ballot_sync(0xffffffffu, res);
        // ---- release: the last warp out refills this stage with unit it + S ----
        if (lane == 0) {
            const uint32_t unit = *reinterpret_cast<const uint32_t*>(st);
            if (warp_has_keys) out_base[static_cast<size_t>(unit) * row_words32] = bits;
            // Relaxed counter: this warp's reads of the stage were consumed (they decided `bits`)
            // before the add can issue; only the refilling thread needs the fences.
            const uint32_t old = atom_add_relaxed_shared(&done[s], 1u);
            if (old == n_warps - 1) {
                fence_acq_rel_cta();
                done[s] = 0;
                if (TRACE && tr && 2 + 2 * it < trace_slots) tr[2 + 2 * it] = globaltimer_ns();
                const uint32_t nxt = it + S;
                if (nxt < my_count) {
                    const uint4 a = *reinterpret_cast<const uint4*>(st + kStageRowBytes);
                    const uint4 b = *reinterpret_cast<const uint4*>(st + kStageRowBytes + 16);
                    const uint64_t nwb = (static_cast<uint64_t>(a.w) << 32) | a.z;
                    fence_proxy_async();
                    fill_stage(const_cast<uint8_t*>(st), &full[s], stab, blockIdx.x + nxt * G, nxt + S < my_count,
                               blockIdx.x + (nxt + S) * G, words, nwb, b.x, b.y, b.z, kind_mask);
                }
            }
        }
        st += stage_bytes;
        if (++s == S) { s = 0; ph ^= 1u; st = stages; }
    }
}


// ------------------------------------------------- staged path, two phases ---
// probe_staged2: the same ring of bulk-copied units, but TestString is split in two phases so
// that the early exit does not idle lanes (in probe_staged a warp runs until its slowest lane
// is done: ~6.3 of k = 10 tests for 32 absent keys whose mean is 2 -> a third of the lanes work).
//
//   phase A  (warps 0..NA-1, KPT keys per thread, only locations 0..NT-1 of each key kept in
//            registers): the first NT tests of every key, no branch between them (ILP NT*KPT).
//            A key that fails is final (bit 0).  A key whose filter is absent or has k <= NT is
//            final (bit 1).  Every other passing key is a *survivor*: its index is appended to the
//            stage's queue (one shared-memory atomicAdd per warp per unit).  Of the keys absent
//            from a unit 2^-NT survive.
//   phase B  (warps NA..NA+NB-1): wait for the A warps of that unit (mbarrier), then test locations
//            NT..k-1 of the queued survivors, 32 survivors per warp pass (dense lanes), early exit
//            per lane; a survivor that passes everything ORs its bit into the unit's result row in
//            shared memory.  The last B warp out writes the row to HBM with one coalesced store,
//            resets the stage and refills it (same last-arriver refill as probe_staged).
//
// Same bits as probe_staged (TestString is an AND over the k locations; the order in which clear
// bits are discovered does not matter).
constexpr uint32_t kStage2RowBitsOff = kProbeStageHeaderBytes;             // 32 x u32 result row
// u32 survivors in queue: lives in StageRow::pad (always 0 in the row table), so every (re)fill of the stage zeroes it
// through the bulk copy itself — no thread ever has to reset it (racecheck used to report the reset / next-read pair)
constexpr uint32_t kStage2CntOff = 28;
static_assert(offsetof(StageRow, pad) == kStage2CntOff, "the survivor count aliases StageRow::pad");
constexpr uint32_t kStage2QueueOff = kStage2RowBitsOff + 128 + 16;         // 1024 x u16 key indexes
static_assert(kStage2QueueOff + 2 * kProbeMaxKeysPerPass == kProbeStage2HeaderBytes, "stage2 header layout");
static_assert(kProbeStage2HeaderBytes % 16 == 0, "bulk copies need 16-byte aligned destinations");

// Every lane that published into the stage (result-row words, survivor queue entries) arrives on `aready` itself,
// so the publication does not lean on the warp barrier's transitivity (racecheck reported that pair in round 1).
constexpr bool kAllLanesArrive = true;

template <int NA, int KPT, int NT, int NB, int T, bool TRACE>
__global__ void __launch_bounds__((NA + NB) * 32, 1)
probe_staged2_kernel(const StageRow* __restrict__ stab, uint32_t n_list_host, const uint32_t* __restrict__ n_list_dev,
                     const uint64_t* __restrict__ words, const uint64_t* __restrict__ hashes,
                     const uint8_t* __restrict__ kinds, uint32_t key_base, uint32_t n_keys, uint32_t kind_mask,
                     uint32_t* __restrict__ matrix32, uint32_t row_words32, uint32_t n_stages, uint32_t stage_bytes,
                     uint32_t relax_sleep_ns, uint64_t* __restrict__ trace, uint32_t trace_slots,
                     const uint8_t* __restrict__ key_bytes, const uint64_t* __restrict__ key_off,
                     uint64_t* hash_scratch) {
    static_assert(NA * KPT * 32 == static_cast<int>(kProbeMaxKeysPerPass), "A warps must cover one pass of keys");
    static_assert(NA + NB <= 32 && NT >= 1 && NT <= 4 && NB % T == 0, "shape");
    // B warps work in teams of T; team g serves units g, g+NTEAMS, ...  The host guarantees
    // n_stages % NTEAMS == 0: a team then revisits only its own stages, so it can never wait for a
    // fill two phases ahead of a stage's mbarrier (parity waits alias beyond one phase).
    constexpr uint32_t NTEAMS = NB / T;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* aready = full + kProbeMaxStages;
    uint64_t* bdone = aready + kProbeMaxStages;   // every lane of the team's B warps is done reading the stage
    static_assert(3 * kProbeMaxStages * 8 <= kProbe2SmemPrefixBytes, "prefix holds the three mbarrier arrays");
    uint8_t* stages = smem + kProbe2SmemPrefixBytes;

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t warp = tid >> 5;
    const uint32_t G = gridDim.x;
    const uint32_t S = n_stages;
    // optional timeline (profiling only), per CTA: [0] start; per unit it: [1+8it] unit resident (A warp 0),
    // [2+8it] A warp 0 done, [3+8it] all A warps done (seen by the first B warp), [4+8it] released,
    // [5+8it] chunk 0: survivor hashes loaded, [6+8it] chunk 0 tested, [7+8it] B warp 0 arrived on the counter
    uint64_t* tr = (TRACE && trace) ? trace + static_cast<size_t>(blockIdx.x) * trace_slots : nullptr;
    if (TRACE && tr && tid == 0) tr[0] = globaltimer_ns();

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&aready[s], kAllLanesArrive ? NA * 32 : NA);
            mbar_init(&bdone[s], T * 32);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // PDL: the ring's first fills read only the immutable corpus, so they may overlap the tail of the
    // previous kernel in the stream (the hash kernel of this batch, or the previous batch's probe);
    // everything the predecessor wrote (hashes, the compacted unit list) is touched after the wait.
    if (n_list_dev) griddep_wait();
    const uint32_t n_list = n_list_dev ? __ldg(n_list_dev) : n_list_host;
    const uint32_t my_count = n_list > blockIdx.x ? (n_list - blockIdx.x + G - 1) / G : 0;

    if (warp == 0 && lane < S && lane < my_count) {  // prologue: lane l fills stage l
        const uint32_t li = blockIdx.x + lane * G;
        const uint4* hp = reinterpret_cast<const uint4*>(&stab[li]);
        const uint4 a = __ldg(hp), b = __ldg(hp + 1);
        const uint64_t word_base = (static_cast<uint64_t>(a.w) << 32) | a.z;
        fill_stage(stages + static_cast<size_t>(lane) * stage_bytes, &full[lane], stab, li, lane + S < my_count,
                   li + S * G, words, word_base, b.x, b.y, b.z, kind_mask, kProbeStage2HeaderBytes);
    }
    griddep_launch_dependents();
    if (!n_list_dev) griddep_wait();

    // Fused hashing (bsg_probe path): no separate hash kernel — while the ring's first fills are in
    // flight every CTA hashes this pass's keys itself (one key per thread, baseHashes as in
    // kernels_hash.cu) into its own 32 KB slice of a scratch table; phase B reads survivors' hashes
    // from that slice.  The table is written by this kernel: it is read with coherent loads only.
    const uint64_t* htab = hashes;
    uint32_t hbase = 0;
    if (hash_scratch) {
        uint64_t* tab = hash_scratch + static_cast<size_t>(blockIdx.x) * kProbeMaxKeysPerPass * 4;
        for (uint32_t t = tid; t < n_keys; t += blockDim.x) {
            const uint64_t b = __ldg(&key_off[key_base + t]), e = __ldg(&key_off[key_base + t + 1]);
            uint64_t h[4];
            base_hashes(key_bytes + b, static_cast<uint32_t>(e - b), h);
            ulonglong2* out = reinterpret_cast<ulonglong2*>(tab + 4ull * t);
            out[0] = make_ulonglong2(h[0], h[1]);
            out[1] = make_ulonglong2(h[2], h[3]);
        }
        __syncthreads();
        htab = tab;
        hbase = key_base;
    }

    uint32_t s = 0, ph = 0;
    uint8_t* st = stages;
    if (warp < NA) {
        // ------------------------------------------------------------ phase A ---
        uint64_t loc[KPT][NT];  // location(h, 0..NT-1) = h0, h1+h3, h0+2*h3, h1+3*h2
        uint32_t f_off[KPT];    // byte offset of the key's StageFilter in the stage row; 0 = no key
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const uint32_t ql = (warp * KPT + j) * 32 + lane;
#pragma unroll
            for (int t = 0; t < NT; ++t) loc[j][t] = 0;
            f_off[j] = 0;
            if (ql < n_keys) {
                const uint32_t q = key_base + ql;
                const uint64_t* hp = htab + 4ull * (q - hbase);
                const ulonglong2 a = ld_global_u64x2(hp), b = ld_global_u64x2(hp + 2);
                const uint64_t l4[4] = {a.x, a.y + b.y, a.x + 2 * b.y, a.y + 3 * b.x};
#pragma unroll
                for (int t = 0; t < NT; ++t) loc[j][t] = l4[t];
                f_off[j] = 32u + 32u * __ldg(&kinds[q]);
            }
        }
        const uint32_t lt_mask = (1u << lane) - 1u;
        for (uint32_t it = 0; it < my_count; ++it) {
            mbar_wait(&full[s], ph);
            if (TRACE && tr && tid == 0 && 1 + 8 * it < trace_slots) tr[1 + 8 * it] = globaltimer_ns();
            uint32_t fin_bits[KPT], surv_bits[KPT];
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                bool fin = false, surv = false;
                if (f_off[j]) {
                    const uint4 f = *reinterpret_cast<const uint4*>(st + f_off[j]);  // m, k, ih, il
                    if (f.x == 0) {
                        fin = true;  // absent filter cannot disqualify (query_exec.go:137-151)
                    } else {
                        const uint32_t rel = *reinterpret_cast<const uint32_t*>(st + f_off[j] + 16);
                        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(st + kProbeStage2HeaderBytes + rel);
                        uint32_t bit[NT], wv[NT];
#pragma unroll
                        for (int t = 0; t < NT; ++t) bit[t] = mod_m32(loc[j][t], f.x, f.z, f.w);
#pragma unroll
                        for (int t = 0; t < NT; ++t) wv[t] = w32[bit[t] >> 5];
                        bool pass = true;
#pragma unroll
                        for (int t = 0; t < NT; ++t)  // location t exists only when t < k
                            pass &= static_cast<bool>(((wv[t] >> (bit[t] & 31u)) & 1u) | (f.y <= static_cast<uint32_t>(t)));
                        fin = pass & (f.y <= static_cast<uint32_t>(NT));
                        surv = pass & (f.y > static_cast<uint32_t>(NT));
                    }
                }
                fin_bits[j] = __ballot_sync(0xffffffffu, fin);
                surv_bits[j] = __ballot_sync(0xffffffffu, surv);
            }
            uint32_t total = 0;
#pragma unroll
            for (int j = 0; j < KPT; ++j) total += __popc(surv_bits[j]);
            uint32_t base = 0;
            if (lane == 0) {
                uint32_t* r32 = reinterpret_cast<uint32_t*>(st + kStage2RowBitsOff) + warp * KPT;
#pragma unroll
                for (int j = 0; j < KPT; ++j) r32[j] = fin_bits[j];
                if (total) base = atomicAdd(reinterpret_cast<uint32_t*>(st + kStage2CntOff), total);
            }
            __syncwarp();  // reconverge before the shuffle (else it takes the divergent slow path)
            if (total) {
                base = __shfl_sync(0xffffffffu, base, 0);
                uint16_t* queue = reinterpret_cast<uint16_t*>(st + kStage2QueueOff);
#pragma unroll
                for (int j = 0; j < KPT; ++j) {
                    if ((surv_bits[j] >> lane) & 1u)
                        queue[base + __popc(surv_bits[j] & lt_mask)] = static_cast<uint16_t>((warp * KPT + j) * 32 + lane);
                    base += __popc(surv_bits[j]);
                }
            }
            // Publication order: every lane's queue / row stores -> bar.warp.sync (orders memory among the
            // warp's lanes) -> lane 0's mbarrier.arrive (release.cta, cumulative) -> a B warp's try_wait
            // (acquire.cta) -> its loads.  compute-sanitizer racecheck does not credit the warp barrier's
            // transitivity and reports the queue write/read pair; memcheck and synccheck are clean.
            __syncwarp();
            if (kAllLanesArrive || lane == 0) mbar_arrive(&aready[s]);
            if (TRACE && tr && tid == 0 && 2 + 8 * it < trace_slots) tr[2 + 8 * it] = globaltimer_ns();
            st += stage_bytes;
            if (++s == S) { s = 0; ph ^= 1u; st = stages; }
        }
    } else {
        // ------------------------------------------------------------ phase B ---
        const uint32_t wb = warp - NA;
        const uint32_t team = wb / T, member = wb % T;
        // every word of the row that belongs to this pass, pad word included (A warps without keys wrote zeros)
        const uint32_t out_words = min(32u, row_words32 - (key_base >> 5));
        uint32_t* out_base = matrix32 + (key_base >> 5);
        s = team;
        while (s >= S) { s -= S; ph ^= 1u; }
        for (uint32_t it = team; it < my_count; it += NTEAMS) {
            st = stages + static_cast<size_t>(s) * stage_bytes;
            mbar_wait(&full[s], ph);    // the bulk copy's bytes (async proxy) are visible
            // every A warp has published its row words and survivors (usually the long wait of a B warp)
            mbar_wait_relaxed(&aready[s], ph, 1000u, relax_sleep_ns);
            if (TRACE && tr && member == 0 && lane == 0 && 3 + 8 * it < trace_slots) tr[3 + 8 * it] = globaltimer_ns();
            // the count was built with shared-memory atomics by the A warps; it is read the same way (one lane, then a
            // shuffle) so that the tools see an atomic / atomic pair on this word
            uint32_t n = 0;
            if (lane == 0) n = atomicAdd(reinterpret_cast<uint32_t*>(st + kStage2CntOff), 0u);
            n = __shfl_sync(0xffffffffu, n, 0);
            const uint32_t n_chunks = (n + 31) >> 5;
            const uint16_t* queue = reinterpret_cast<const uint16_t*>(st + kStage2QueueOff);
            // rotate the first chunk over the team's warps from unit to unit (even load per warp)
            for (uint32_t c = (member + T - ((it / NTEAMS) % T)) % T; c < n_chunks; c += T) {
                const uint32_t idx = c * 32 + lane;
                if (idx < n) {
                    const uint32_t ql = queue[idx];
                    const uint32_t q = key_base + ql;
                    const uint64_t* hp = htab + 4ull * (q - hbase);
                    const ulonglong2 a = ld_global_u64x2(hp), b = ld_global_u64x2(hp + 2);
                    const uint32_t fo = 32u + 32u * __ldg(&kinds[q]);
                    if (TRACE && tr && c == 0 && lane == 0 && 5 + 8 * it < trace_slots)
                        tr[5 + 8 * it] = globaltimer_ns() + ((a.x ^ b.y ^ fo) == 0x123456789abcull);  // hashes arrived
                    const uint4 f = *reinterpret_cast<const uint4*>(st + fo);
                    const uint32_t rel = *reinterpret_cast<const uint32_t*>(st + fo + 16);
                    if (test_tail_s32<NT>(a.x, a.y, b.x, b.y, f.x, f.z, f.w, f.y,
                                          reinterpret_cast<const uint32_t*>(st + kProbeStage2HeaderBytes + rel)))
                        atomicOr(reinterpret_cast<uint32_t*>(st + kStage2RowBitsOff) + (ql >> 5), 1u << (ql & 31u));
                }
                __syncwarp();
                if (TRACE && tr && c == 0 && lane == 0 && 6 + 8 * it < trace_slots) tr[6 + 8 * it] = globaltimer_ns();  // chunk 0 tested
            }
            // hand-off of the stage: every lane that read it arrives on bdone[s]; the team's first warp waits for the
            // whole team (an mbarrier phase per use of the stage: same parity as full[s]), then emits the row and refills.
            // (Rounds 1-2 elected the last arriver with an acq_rel counter: same order, but a hand-off that
            // compute-sanitizer racecheck does not model — it reported the refill against the team's reads.)
            mbar_arrive(&bdone[s]);
            if (TRACE && tr && member == 0 && lane == 0 && 7 + 8 * it < trace_slots) tr[7 + 8 * it] = globaltimer_ns();
            if (member == 0) {
                mbar_wait(&bdone[s], ph);
                const uint32_t unit = *reinterpret_cast<const uint32_t*>(st);
                if (lane < out_words)
                    out_base[static_cast<size_t>(unit) * row_words32 + lane] =
                        ld_volatile_shared_u32(st + kStage2RowBitsOff + 4 * lane);
                __syncwarp();
                if (lane == 0) {
                    if (TRACE && tr && 4 + 8 * it < trace_slots) tr[4 + 8 * it] = globaltimer_ns();
                    const uint32_t nxt = it + S;
                    if (nxt < my_count) {
                        const uint4 a = *reinterpret_cast<const uint4*>(st + kStageRowBytes);
                        const uint4 b = *reinterpret_cast<const uint4*>(st + kStageRowBytes + 16);
                        const uint64_t nwb = (static_cast<uint64_t>(a.w) << 32) | a.z;
                        fence_proxy_async();
                        fill_stage(st, &full[s], stab, blockIdx.x + nxt * G, nxt + S < my_count,
                                   blockIdx.x + (nxt + S) * G, words, nwb, b.x, b.y, b.z, kind_mask,
                                   kProbeStage2HeaderBytes);
                    }
                }
                __syncwarp();
            }
            s += NTEAMS;
            while (s >= S) { s -= S; ph ^= 1u; }
        }
    }
}

// The shapes kept for measurement (BSG_PROBE_VARIANT): <A warps, keys per A thread, A tests, B warps, B team size>;
// 3 is the default (api.cu), the others are kept for the parity test and for measurement
template <int NA, int KPT, int NT, int NB, int T>
static cudaError_t staged2_configure(int max_smem_optin) {
    cudaError_t e = cudaFuncSetAttribute(probe_staged2_kernel<NA, KPT, NT, NB, T, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(probe_staged2_kernel<NA, KPT, NT, NB, T, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    if constexpr (T != NB) return staged2_configure<NA, KPT, NT, NB, NB>(max_smem_optin);
    return cudaSuccess;
}
template <int NA, int KPT, int NT, int NB, int T>
static void staged2_launch(const ProbeStagedPlan& plan, const StageRow* d_stab, uint32_t n_list,
                           const uint32_t* d_n_list, const uint64_t* d_words, const uint64_t* d_hashes,
                           const uint8_t* d_kinds, uint32_t key_base, uint32_t n_keys, uint32_t kind_mask,
                           uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s, uint64_t* d_trace,
                           uint32_t trace_slots) {
    const uint32_t sb = kProbeStage2HeaderBytes + plan.stage_data_bytes;
    constexpr uint32_t NTEAMS = NB / T;
    uint32_t n_stages = static_cast<uint32_t>(plan.n_stages);
    if constexpr (NTEAMS > 1) {
        if (n_stages < NTEAMS) {  // ring too short for teams (large units): one team of all B warps
            staged2_launch<NA, KPT, NT, NB, NB>(plan, d_stab, n_list, d_n_list, d_words, d_hashes, d_kinds, key_base,
                                                n_keys, kind_mask, d_matrix32, row_words32, s, d_trace, trace_slots);
            return;
        }
        n_stages -= n_stages % NTEAMS;  // a team must own its stages (see the kernel)
    }
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
    uint64_t* tr = d_trace;
    uint32_t ts = d_trace ? trace_slots : 0;
    if (d_trace)
        cudaLaunchKernelEx(&cfg, probe_staged2_kernel<NA, KPT, NT, NB, T, true>, d_stab, n_list, d_n_list, d_words, d_hashes,
                           d_kinds, key_base, n_keys, kind_mask, d_matrix32, row_words32, n_stages, sb,
                           plan.relax_sleep_ns, tr, ts, plan.fuse_keys, plan.fuse_key_off, plan.fuse_scratch);
    else
        cudaLaunchKernelEx(&cfg, probe_staged2_kernel<NA, KPT, NT, NB, T, false>, d_stab, n_list, d_n_list, d_words, d_hashes,
                           d_kinds, key_base, n_keys, kind_mask, d_matrix32, row_words32, n_stages, sb,
                           plan.relax_sleep_ns, tr, ts, plan.fuse_keys, plan.fuse_key_off, plan.fuse_scratch);
}
#define BSG_STAGED2_SHAPES(X)                                                                          \
    X(1, 16, 2, 2, 16, 16) X(2, 16, 2, 3, 16, 16) X(3, 16, 2, 3, 16, 4) X(4, 16, 2, 3, 16, 2) X(5, 16, 2, 4, 16, 4)

cudaError_t probe_staged_configure(int max_smem_optin) {
    cudaError_t e = cudaFuncSetAttribute(probe_staged_kernel<1024, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
#define X(id, na, kpt, nt, nb, t) \
    e = staged2_configure<na, kpt, nt, nb, t>(max_smem_optin); \
    if (e != cudaSuccess) return e;
    BSG_STAGED2_SHAPES(X)
#undef X
    return cudaFuncSetAttribute(probe_staged_kernel<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                max_smem_optin);
}

cudaError_t launch_probe_staged(const ProbeStagedPlan& plan, const StageRow* d_stab, uint32_t n_list,
                                const uint64_t* d_words, const uint64_t* d_hashes, const uint8_t* d_kinds,
                                uint32_t key_base, uint32_t n_keys, uint32_t kind_mask, uint32_t* d_matrix32,
                                uint32_t row_words32, cudaStream_t s, uint64_t* d_trace, uint32_t trace_slots,
                                const uint32_t* d_n_list) {
    if ((n_list == 0 && !d_n_list) || n_keys == 0) return cudaSuccess;
    if (n_keys > kProbeMaxKeysPerPass) return cudaErrorInvalidValue;
    switch (plan.variant) {  // two-phase kernel shapes
#define X(id, na, kpt, nt, nb, t) \
        case id: \
            staged2_launch<na, kpt, nt, nb, t>(plan, d_stab, n_list, d_n_list, d_words, d_hashes, d_kinds, key_base, \
                                            n_keys, kind_mask, d_matrix32, row_words32, s, d_trace, trace_slots); \
            return cudaGetLastError();
        BSG_STAGED2_SHAPES(X)
#undef X
        default: break;
    }
    const uint32_t stage_bytes = kProbeStageHeaderBytes + plan.stage_data_bytes;
    // one key per thread; at least 4 warps so a small batch still has some latency hiding
    uint32_t warps = (n_keys + 31) / 32;
    if (plan.warps > 0 && static_cast<uint32_t>(plan.warps) > warps) warps = plan.warps;
    if (warps < 4) warps = 4;
    if (warps > 32) warps = 32;
    if (d_trace)
        probe_staged_kernel<1024, true><<<dim3(plan.grid), dim3(warps * 32), plan.smem_bytes, s>>>(
            d_stab, n_list, d_n_list, d_words, d_hashes, d_kinds, key_base, n_keys, kind_mask, d_matrix32, row_words32,
            static_cast<uint32_t>(plan.n_stages), stage_bytes, plan.stagger_ns, d_trace, trace_slots);
    else
        probe_staged_kernel<1024, false><<<dim3(plan.grid), dim3(warps * 32), plan.smem_bytes, s>>>(
            d_stab, n_list, d_n_list, d_words, d_hashes, d_kinds, key_base, n_keys, kind_mask, d_matrix32, row_words32,
            static_cast<uint32_t>(plan.n_stages), stage_bytes, plan.stagger_ns, nullptr, 0);
    return cudaGetLastError();
}

// ------------------------------------------------------------ gather path ---
// One lane per (unit, key).  G = lanes per unit (power of two <= 32) when the
// batch has <= 32 keys, so a warp covers 32/G units; otherwise a warp covers
// one 32-key chunk of one unit.
__global__ void __launch_bounds__(256)
probe_gather_kernel(const DevFilter* __restrict__ udesc, const uint64_t* __restrict__ words,
                    const uint32_t* __restrict__ unit_list, uint32_t n_list, const uint64_t* __restrict__ hashes,
                    const uint8_t* __restrict__ kinds, uint32_t n_keys, uint32_t g_log2, uint32_t chunks,
                    uint32_t* __restrict__ matrix32, uint32_t row_words32, const uint32_t* __restrict__ parent,
                    const uint32_t* __restrict__ parent_mask32) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gwarp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    uint32_t list_idx, key, grp = 0;
    const uint32_t G = 1u << g_log2;
    if (chunks == 1) {
        const uint32_t per_warp = 32u >> g_log2;
        grp = lane >> g_log2;
        list_idx = static_cast<uint32_t>(gwarp * per_warp + grp);
        key = lane & (G - 1);
    } else {
        list_idx = static_cast<uint32_t>(gwarp / chunks);
        key = static_cast<uint32_t>(gwarp % chunks) * 32 + lane;
    }
    const bool unit_ok = chunks == 1 ? (gwarp * (32u >> g_log2) + grp < n_list) : (gwarp / chunks < n_list);
    bool res = false;
    uint32_t unit = 0;
    if (unit_ok) {
        unit = unit_list ? __ldg(&unit_list[list_idx]) : list_idx;
        bool alive = true;  // hierarchical probe: a block whose file was disqualified is not read at all
        if (parent) {
            const uint32_t f = __ldg(&parent[unit]);
            alive = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
        }
        if (alive && key < n_keys) {
            const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hashes + 4ull * key);
            const ulonglong2 a = __ldg(hp), b = __ldg(hp + 1);
            const uint32_t kd = __ldg(&kinds[key]);
            const uint4* fp = reinterpret_cast<const uint4*>(&udesc[static_cast<size_t>(unit) * 3 + kd]);
            const uint4 f0 = __ldg(fp), f1 = __ldg(fp + 1);
            const uint64_t word_off = (static_cast<uint64_t>(f0.y) << 32) | f0.x;
            const uint64_t m = (static_cast<uint64_t>(f0.w) << 32) | f0.z;
            const uint64_t inv = (static_cast<uint64_t>(f1.y) << 32) | f1.x;
            const uint32_t k = f1.z;
            if (m == 0) {
                res = true;
            } else {
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(words + word_off);
                if (m < kSmallModLimit)
                    res = test_hashes_g32(a.x, a.y, b.x, b.y, static_cast<uint32_t>(m), static_cast<uint32_t>(inv >> 32),
                                          static_cast<uint32_t>(inv), k, w32);
                else
                    res = test_hashes(a.x, a.y, b.x, b.y, m, inv, k, [&](uint64_t idx) { return __ldg(w32 + idx); });
            }
        }
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, res);
    if (chunks == 1) {
        if (unit_ok && (lane & (G - 1)) == 0) {
            const uint32_t mine = G == 32 ? bits : ((bits >> (grp << g_log2)) & ((1u << G) - 1u));
            matrix32[static_cast<size_t>(unit) * row_words32] = mine;
        }
    } else {
        if (unit_ok && lane == 0)
            matrix32[static_cast<size_t>(unit) * row_words32 + static_cast<uint32_t>(gwarp % chunks)] = bits;
    }
}

// The same probe for callers that want only the candidate mask of a SMALL query (<= 32 keys, one expression):
// evaluateBloomExpression short-circuits (query_exec.go:105-119: an AND stops at its first failing child, an OR at
// its first passing one), so in the reference a key whose sibling already decided the unit is never tested.  The
// batched form of that: the G lanes of a unit test their keys in rounds (locations 0-2, 3-6, the rest); after each
// round the expression is evaluated OPTIMISTICALLY on the unit's bits (a key that has not failed yet counts as
// present).  Bloom tests and AND/OR trees are monotone, so an optimistic "false" is final: every lane of that unit
// stops.  Present keys — the ones that cost all k locations — stop after 3 or 7 locations wherever another branch
// has already disqualified the unit.  The row written is the optimistic one: the tree kernel derives exactly the
// reference's mask from it, but it is not the membership matrix (callers that want the matrix take the plain kernel).
constexpr uint32_t kGatherScMaxOps = 128;
__global__ void __launch_bounds__(256)
probe_gather_sc_kernel(const DevFilter* __restrict__ udesc, const uint64_t* __restrict__ words,
                       const uint32_t* __restrict__ unit_list, uint32_t n_list, const uint64_t* __restrict__ hashes,
                       const uint8_t* __restrict__ kinds, uint32_t n_keys, uint32_t g_log2,
                       const bsg_expr_op* __restrict__ prog, uint32_t prog_len, uint32_t* __restrict__ matrix32,
                       uint32_t row_words32, const uint32_t* __restrict__ parent, const uint32_t* __restrict__ parent_mask32) {
    __shared__ bsg_expr_op sprog[kGatherScMaxOps];
    for (uint32_t i = threadIdx.x; i < prog_len; i += blockDim.x) sprog[i] = prog[i];
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t gwarp = (static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint32_t G = 1u << g_log2;
    const uint32_t per_warp = 32u >> g_log2;
    const uint32_t grp = lane >> g_log2;
    const uint32_t key = lane & (G - 1);
    const uint64_t list_idx = gwarp * per_warp + grp;
    const bool unit_ok = list_idx < n_list;
    uint32_t unit = 0;
    bool pass = false;           // this lane's key has not failed a location yet
    uint32_t i = 0, k = 0;       // next location, number of locations
    uint64_t h0 = 0, h1 = 0, h2 = 0, h3 = 0, m = 0, inv = 0;
    const uint32_t* w32 = nullptr;
    if (unit_ok) {
        unit = unit_list ? __ldg(&unit_list[list_idx]) : static_cast<uint32_t>(list_idx);
        bool alive = true;
        if (parent) {
            const uint32_t f = __ldg(&parent[unit]);
            alive = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
        }
        if (alive && key < n_keys) {
            const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hashes + 4ull * key);
            const ulonglong2 a = __ldg(hp), b = __ldg(hp + 1);
            h0 = a.x; h1 = a.y; h2 = b.x; h3 = b.y;
            const uint32_t kd = __ldg(&kinds[key]);
            const uint4* fp = reinterpret_cast<const uint4*>(&udesc[static_cast<size_t>(unit) * 3 + kd]);
            const uint4 f0 = __ldg(fp), f1 = __ldg(fp + 1);
            m = (static_cast<uint64_t>(f0.w) << 32) | f0.z;
            inv = (static_cast<uint64_t>(f1.y) << 32) | f1.x;
            k = m ? f1.z : 0u;   // absent filter: passes, nothing to test
            w32 = reinterpret_cast<const uint32_t*>(words + ((static_cast<uint64_t>(f0.y) << 32) | f0.x));
            pass = true;
        }
    }
    const uint32_t shift = grp << g_log2;
    const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    uint32_t mine = 0;
    for (uint32_t round = 0;; ++round) {
        // the locations of a round are loaded TOGETHER (independent loads, no exit between them): every one of them is
        // a DRAM round trip on a corpus of many GB, and a warp is as slow as its longest chain of dependent misses —
        // a present key costs 3 round trips this way instead of k
        const uint32_t upto = round == 0 ? 3u : round == 1 ? 7u : i + 4u;
        if (pass && i < k) {
            uint32_t wv[4], sh[4];
#pragma unroll
            for (uint32_t t = 0; t < 4; ++t) {
                const uint32_t ii = i + t;
                const bool on = ii < k && ii < upto;
                const uint64_t a = (ii & 1u) ? h1 : h0;
                const uint64_t b = (((ii + (ii & 1u)) & 3u) >> 1) ? h3 : h2;
                const uint64_t loc = a + static_cast<uint64_t>(ii) * b;
                uint64_t bit;
                if (m < kSmallModLimit)
                    bit = mod_m32(loc, static_cast<uint32_t>(m), static_cast<uint32_t>(inv >> 32), static_cast<uint32_t>(inv));
                else
                    bit = mod_m(loc, m, inv);
                sh[t] = static_cast<uint32_t>(bit) & 31u;
                wv[t] = on ? __ldg(w32 + (bit >> 5)) : 0xffffffffu;
            }
            uint32_t ok = 1u;
#pragma unroll
            for (uint32_t t = 0; t < 4; ++t) ok &= wv[t] >> sh[t];
            pass = ok & 1u;
            i = min(k, upto);
        }
        mine = (__ballot_sync(0xffffffffu, pass) >> shift) & gmask;
        if (__ballot_sync(0xffffffffu, pass && i < k) == 0u) break;   // every key of the warp has failed or finished
        // optimistic value of the expression on this unit's bits
        uint64_t stack = 0;
        for (uint32_t pc = 0; pc < prog_len; ++pc) {
            const uint32_t op = sprog[pc].op, arg = sprog[pc].arg;
            if (op == BSG_OP_LEAF) {
                stack = (stack << 1) | ((mine >> arg) & 1u);
            } else if (op == BSG_OP_TRUE) {
                stack = (stack << 1) | 1ull;
            } else if (op == BSG_OP_FALSE) {
                stack = stack << 1;
            } else {
                const uint64_t msk = arg >= 64 ? ~0ull : ((1ull << arg) - 1ull);
                const uint64_t top = stack & msk;
                const uint64_t v = (op == BSG_OP_AND) ? (top == msk) : (top != 0);
                stack = arg >= 64 ? 0ull : (stack >> arg);
                stack = (stack << 1) | v;
            }
        }
        if (!(stack & 1ull)) i = k;   // the unit is disqualified whatever the remaining locations say
    }
    if (unit_ok && key == 0) matrix32[static_cast<size_t>(unit) * row_words32] = mine;
}

cudaError_t launch_probe_gather(const DevFilter* d_udesc, const uint64_t* d_words, const uint32_t* d_unit_list,
                                uint32_t n_list, const uint64_t* d_hashes, const uint8_t* d_kinds, uint32_t n_keys,
                                uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s, const uint32_t* d_parent,
                                const uint32_t* d_parent_mask32, const bsg_expr_op* d_prog, uint32_t prog_len) {
    if (n_list == 0 || n_keys == 0) return cudaSuccess;
    if (d_prog && prog_len && prog_len <= kGatherScMaxOps && n_keys <= 32) {   // mask-only small query: short-circuit form
        uint32_t g_log2 = 0;
        while ((1u << g_log2) < n_keys) ++g_log2;
        const uint32_t per_warp = 32u >> g_log2;
        const uint64_t n_warps = (static_cast<uint64_t>(n_list) + per_warp - 1) / per_warp;
        const uint64_t n_blocks = (n_warps + 7) / 8;
        if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
        probe_gather_sc_kernel<<<static_cast<uint32_t>(n_blocks), 256, 0, s>>>(
            d_udesc, d_words, d_unit_list, n_list, d_hashes, d_kinds, n_keys, g_log2, d_prog, prog_len, d_matrix32,
            row_words32, d_parent, d_parent_mask32);
        return cudaGetLastError();
    }
    uint32_t g_log2 = 5, chunks = 1;
    uint64_t n_warps;
    if (n_keys <= 32) {
        g_log2 = 0;
        while ((1u << g_log2) < n_keys) ++g_log2;
        const uint32_t per_warp = 32u >> g_log2;
        n_warps = (static_cast<uint64_t>(n_list) + per_warp - 1) / per_warp;
    } else {
        chunks = (n_keys + 31) / 32;
        n_warps = static_cast<uint64_t>(n_list) * chunks;
    }
    const uint32_t warps_per_block = 8;
    const uint64_t n_blocks = (n_warps + warps_per_block - 1) / warps_per_block;
    if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    probe_gather_kernel<<<static_cast<uint32_t>(n_blocks), warps_per_block * 32, 0, s>>>(
        d_udesc, d_words, d_unit_list, n_list, d_hashes, d_kinds, n_keys, g_log2, chunks, d_matrix32, row_words32,
        d_parent, d_parent_mask32);
    return cudaGetLastError();
}

// -------------------------------------------------------------- tree eval ---
// One thread per unit; evaluation stack is a 64-bit bit-stack (BSG_MAX_STACK).
constexpr uint32_t kTreeSmemOps = 4096;   // 32 KB of dynamic shared memory (no opt-in needed)
__global__ void __launch_bounds__(256)
tree_eval_kernel(const uint32_t* __restrict__ matrix32, uint32_t row_words32, uint64_t n_units,
                 const bsg_expr_op* __restrict__ prog, uint32_t prog_len, uint32_t* __restrict__ mask32,
                 const uint32_t* __restrict__ parent, const uint32_t* __restrict__ parent_mask32) {
    extern __shared__ bsg_expr_op sprog_buf[];
    // programs up to kTreeSmemOps ops are staged in shared memory, longer ones are read from global (L1/L2 cached)
    const bsg_expr_op* sprog = prog;
    if (prog_len <= kTreeSmemOps) {
        for (uint32_t i = threadIdx.x; i < prog_len; i += blockDim.x) sprog_buf[i] = prog[i];
        __syncthreads();
        sprog = sprog_buf;
    }
    const uint64_t unit = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool alive = false;
    bool parent_alive = true;
    if (unit < n_units && parent) {
        const uint32_t f = __ldg(&parent[unit]);
        parent_alive = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
    }
    if (unit < n_units && parent_alive) {  // rows of disqualified parents were never written
        const uint32_t* row = matrix32 + unit * row_words32;
        uint64_t stack = 0;
        for (uint32_t pc = 0; pc < prog_len; ++pc) {
            const uint32_t op = sprog[pc].op, arg = sprog[pc].arg;
            if (op == BSG_OP_LEAF) {
                const uint32_t bit = (__ldg(&row[arg >> 5]) >> (arg & 31u)) & 1u;
                stack = (stack << 1) | bit;
            } else if (op == BSG_OP_TRUE) {
                stack = (stack << 1) | 1ull;
            } else if (op == BSG_OP_FALSE) {
                stack = stack << 1;
            } else {
                const uint64_t msk = arg >= 64 ? ~0ull : ((1ull << arg) - 1ull);
                const uint64_t top = stack & msk;
                const uint64_t v = (op == BSG_OP_AND) ? (top == msk) : (top != 0);
                stack = arg >= 64 ? 0ull : (stack >> arg);
                stack = (stack << 1) | v;
            }
        }
        alive = stack & 1ull;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && (unit - (threadIdx.x & 31)) < n_units) mask32[unit >> 5] = bits;
}

cudaError_t launch_tree_eval(const uint32_t* d_matrix32, uint32_t row_words32, uint64_t n_units,
                             const bsg_expr_op* d_prog, uint32_t prog_len, uint32_t* d_mask32, cudaStream_t s,
                             const uint32_t* d_parent, const uint32_t* d_parent_mask32) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_blocks = (n_units + 255) / 256;
    if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    tree_eval_kernel<<<static_cast<uint32_t>(n_blocks), 256, (prog_len <= kTreeSmemOps ? prog_len : 0) * sizeof(bsg_expr_op), s>>>(
        d_matrix32, row_words32, n_units, d_prog, prog_len, d_mask32, d_parent, d_parent_mask32);
    return cudaGetLastError();
}

// Several queries that were probed as ONE batch (bsg_probe_multi: the keys of concurrent queries share a pass over
// the corpus): blockIdx.y = query; its program is prog[prog_begin[q] .. prog_begin[q+1]) with leaf arguments that
// already index the batch's key columns.  An empty program keeps every unit (query_exec.go:81-83).  Units whose
// section failed to parse are cleared here (bad32, nullable) instead of by a second kernel.
__global__ void __launch_bounds__(256)
tree_eval_multi_kernel(const uint32_t* __restrict__ matrix32, uint32_t row_words32, uint64_t n_units,
                       const bsg_expr_op* __restrict__ prog, const uint32_t* __restrict__ prog_begin,
                       uint32_t* __restrict__ masks32, uint64_t mask_words32, const uint32_t* __restrict__ bad32) {
    const uint32_t q = blockIdx.y;
    const uint32_t pb = __ldg(&prog_begin[q]), pe = __ldg(&prog_begin[q + 1]);
    const uint64_t unit = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool alive = false;
    if (unit < n_units) {
        const uint32_t* row = matrix32 + unit * row_words32;
        uint64_t stack = 1ull;   // an empty program leaves "true"
        for (uint32_t pc = pb; pc < pe; ++pc) {
            const uint32_t op = __ldg(&prog[pc].op), arg = __ldg(&prog[pc].arg);
            if (op == BSG_OP_LEAF) {
                const uint32_t bit = (__ldg(&row[arg >> 5]) >> (arg & 31u)) & 1u;
                stack = (stack << 1) | bit;
            } else if (op == BSG_OP_TRUE) {
                stack = (stack << 1) | 1ull;
            } else if (op == BSG_OP_FALSE) {
                stack = stack << 1;
            } else {
                const uint64_t msk = arg >= 64 ? ~0ull : ((1ull << arg) - 1ull);
                const uint64_t top = stack & msk;
                const uint64_t v = (op == BSG_OP_AND) ? (top == msk) : (top != 0);
                stack = arg >= 64 ? 0ull : (stack >> arg);
                stack = (stack << 1) | v;
            }
        }
        alive = stack & 1ull;
    }
    uint32_t bits = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && (unit - (threadIdx.x & 31)) < n_units) {
        if (bad32) bits &= ~__ldg(&bad32[unit >> 5]);
        masks32[static_cast<uint64_t>(q) * mask_words32 + (unit >> 5)] = bits;
    }
}

cudaError_t launch_tree_eval_multi(const uint32_t* d_matrix32, uint32_t row_words32, uint64_t n_units,
                                   const bsg_expr_op* d_prog, const uint32_t* d_prog_begin, uint32_t n_queries,
                                   uint32_t* d_masks32, uint64_t mask_words32, const uint32_t* d_bad32,
                                   cudaStream_t s) {
    if (n_units == 0 || n_queries == 0) return cudaSuccess;
    const uint64_t n_blocks = (n_units + 255) / 256;
    if (n_blocks > 0x7fffffffull || n_queries > 65535u) return cudaErrorInvalidValue;
    tree_eval_multi_kernel<<<dim3(static_cast<uint32_t>(n_blocks), n_queries), 256, 0, s>>>(
        d_matrix32, row_words32, n_units, d_prog, d_prog_begin, d_masks32, mask_words32, d_bad32);
    return cudaGetLastError();
}

__global__ void fill_mask_kernel(uint32_t* __restrict__ mask32, uint64_t n_units) {
    const uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t n_words = (n_units + 31) / 32;
    if (w >= n_words) return;
    const uint64_t rem = n_units - w * 32;
    mask32[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
}

// mask[u] = parent_mask[parent[u]]  (no expression: a block survives iff its file did)
__global__ void __launch_bounds__(256)
parent_mask_kernel(uint32_t* __restrict__ mask32, uint64_t n_units, const uint32_t* __restrict__ parent,
                   const uint32_t* __restrict__ parent_mask32) {
    const uint64_t unit = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool alive = false;
    if (unit < n_units) {
        const uint32_t f = __ldg(&parent[unit]);
        alive = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && (unit - (threadIdx.x & 31)) < n_units) mask32[unit >> 5] = bits;
}

cudaError_t launch_parent_mask(uint32_t* d_mask32, uint64_t n_units, const uint32_t* d_parent,
                               const uint32_t* d_parent_mask32, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    parent_mask_kernel<<<static_cast<uint32_t>((n_units + 255) / 256), 256, 0, s>>>(d_mask32, n_units, d_parent,
                                                                                      d_parent_mask32);
    return cudaGetLastError();
}

// Stream compaction between the two stages of a hierarchical probe: keeps the stage rows of the
// units whose parent (file) survived.  Order is not preserved (no consumer needs it).
__global__ void __launch_bounds__(256)
compact_rows_kernel(const StageRow* __restrict__ stab, uint32_t n_rows, const uint32_t* __restrict__ parent,
                    const uint32_t* __restrict__ parent_mask32, StageRow* __restrict__ out, uint32_t* __restrict__ n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    if (i < n_rows) {
        const uint32_t f = __ldg(&parent[stab[i].unit]);
        keep = (__ldg(&parent_mask32[f >> 5]) >> (f & 31u)) & 1u;
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0 && bal) base = atomicAdd(n_out, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        const uint4* src = reinterpret_cast<const uint4*>(&stab[i]);
        uint4* dst = reinterpret_cast<uint4*>(&out[base + __popc(bal & ((1u << lane) - 1u))]);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = __ldg(src + j);
    }
}

cudaError_t launch_compact_rows(const StageRow* d_stab, uint32_t n_rows, const uint32_t* d_parent,
                                const uint32_t* d_parent_mask32, StageRow* d_out, uint32_t* d_n_out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_n_out, 0, 4, s);
    if (e != cudaSuccess || n_rows == 0) return e;
    compact_rows_kernel<<<(n_rows + 255) / 256, 256, 0, s>>>(d_stab, n_rows, d_parent, d_parent_mask32, d_out, d_n_out);
    return cudaGetLastError();
}

// mask &= ~bad: units whose filter section failed to parse are never candidates (query_exec.go:580-590:
// the reference records the error and `continue`s, the block is not scanned)
__global__ void __launch_bounds__(256)
mask_andnot_kernel(uint32_t* __restrict__ mask32, const uint32_t* __restrict__ bad32, uint64_t n_words32) {
    const uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (w < n_words32) mask32[w] &= ~__ldg(&bad32[w]);
}

cudaError_t launch_mask_andnot(uint32_t* d_mask32, const uint32_t* d_bad32, uint64_t n_units, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_words = (n_units + 31) / 32;
    mask_andnot_kernel<<<static_cast<uint32_t>((n_words + 255) / 256), 256, 0, s>>>(d_mask32, d_bad32, n_words);
    return cudaGetLastError();
}

cudaError_t launch_fill_mask(uint32_t* d_mask32, uint64_t n_units, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_words = (n_units + 31) / 32;
    fill_mask_kernel<<<static_cast<uint32_t>((n_words + 255) / 256), 256, 0, s>>>(d_mask32, n_units);
    return cudaGetLastError();
}

}  // namespace bsg
