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
//   probe_staged  — large batches: each unit's bitsets are bulk-copied (TMA 1-D,
//                   cp.async.bulk + mbarrier) into a multi-stage shared-memory
//                   ring by a producer warp while consumer warps test their keys
//                   against the previous stages.  Every bitset byte crosses HBM
//                   once per batch: the HBM-roofline regime of SURVEY.md §8(d).
//   probe_gather  — small batches or filters too large to stage: one lane per
//                   (unit,key), bit words gathered straight from L2/HBM (the
//                   sparse 8*k bytes/probe bound).
// tree_eval turns the (unit x key) bit matrix into the candidate mask with the
// query's AND/OR tree (query_exec.go:89-125) in postfix form.
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
// Shared-memory map: [full mbarriers x16][empty mbarriers x16] | stage 0 | stage 1 ...
// stage = [128 B header: 3 x DevFilter, then {u32 unit, u32 pad, u64 word_base}] [unit words]
struct StageHdrTail {
    uint32_t unit;
    uint32_t pad;
    uint64_t word_base;
};

template <int KPT>
__global__ void __launch_bounds__(kProbeThreads, 1)
probe_staged_kernel(const DevFilter* __restrict__ udesc, const UnitTab* __restrict__ utab,
                    const uint64_t* __restrict__ words, const uint32_t* __restrict__ unit_list, uint32_t n_list,
                    const uint64_t* __restrict__ hashes, const uint8_t* __restrict__ kinds, uint32_t key_base,
                    uint32_t n_keys, uint32_t kind_mask, uint32_t* __restrict__ matrix32, uint32_t row_words32,
                    int n_stages, uint32_t stage_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kProbeMaxStages;
    uint8_t* stages = smem + 2 * kProbeMaxStages * sizeof(uint64_t);

    const uint32_t tid = threadIdx.x;
    const uint32_t lane = tid & 31;
    const uint32_t warp = tid >> 5;
    const uint32_t n_cwarps = (blockDim.x >> 5) - 1;  // last warp is the producer

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], n_cwarps);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const uint32_t my_count = n_list > blockIdx.x ? (n_list - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == n_cwarps) {
        // ===== producer warp: lanes prefetch 32 UnitTabs at a time, lane 0 issues bulk copies =====
        for (uint32_t it0 = 0; it0 < my_count; it0 += 32) {
            const uint32_t it_l = it0 + lane;
            uint32_t unit_l = 0;
            UnitTab ut_l = {0, {0, 0, 0}, 0, 0};
            if (it_l < my_count) {
                const uint32_t li = blockIdx.x + it_l * gridDim.x;
                unit_l = unit_list ? __ldg(&unit_list[li]) : li;
                const uint4* p = reinterpret_cast<const uint4*>(&utab[unit_l]);
                const uint4 a = __ldg(p);
                const uint4 b = __ldg(p + 1);
                ut_l.word_base = (static_cast<uint64_t>(a.y) << 32) | a.x;
                ut_l.nw[0] = a.z; ut_l.nw[1] = a.w; ut_l.nw[2] = b.x; ut_l.total = b.y;
            }
            const uint32_t nb = min(32u, my_count - it0);
            for (uint32_t j = 0; j < nb; ++j) {
                const uint32_t unit = __shfl_sync(0xffffffffu, unit_l, j);
                const uint32_t wb_lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(ut_l.word_base), j);
                const uint32_t wb_hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(ut_l.word_base >> 32), j);
                const uint32_t nw0 = __shfl_sync(0xffffffffu, ut_l.nw[0], j);
                const uint32_t nw1 = __shfl_sync(0xffffffffu, ut_l.nw[1], j);
                const uint32_t nw2 = __shfl_sync(0xffffffffu, ut_l.nw[2], j);
                const uint32_t it = it0 + j;
                const int s = it % n_stages;
                const uint32_t ph = (it / n_stages) & 1u;
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    uint8_t* st = stages + static_cast<size_t>(s) * stage_bytes;
                    const uint64_t word_base = (static_cast<uint64_t>(wb_hi) << 32) | wb_lo;
                    StageHdrTail* tail = reinterpret_cast<StageHdrTail*>(st + 3 * sizeof(DevFilter));
                    tail->unit = unit;
                    tail->pad = 0;
                    tail->word_base = word_base;
                    const uint32_t b0 = (kind_mask & 1u) ? nw0 * 8u : 0u;
                    const uint32_t b1 = (kind_mask & 2u) ? nw1 * 8u : 0u;
                    const uint32_t b2 = (kind_mask & 4u) ? nw2 * 8u : 0u;
                    mbar_arrive_expect_tx(&full[s], 3u * sizeof(DevFilter) + b0 + b1 + b2);
                    bulk_g2s(st, &udesc[static_cast<size_t>(unit) * 3], 3u * sizeof(DevFilter), &full[s]);
                    uint8_t* data = st + kProbeStageHeaderBytes;
                    const uint64_t* src = words + word_base;
                    if (kind_mask == 7u) {
                        if (b0 + b1 + b2) bulk_g2s(data, src, b0 + b1 + b2, &full[s]);
                    } else {
                        if (b0) bulk_g2s(data, src, b0, &full[s]);
                        if (b1) bulk_g2s(data + nw0 * 8u, src + nw0, b1, &full[s]);
                        if (b2) bulk_g2s(data + (nw0 + nw1) * 8u, src + nw0 + nw1, b2, &full[s]);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ===== consumer warps: thread owns keys key_base + j*(n_cwarps*32) + tid =====
        const uint32_t cthreads = n_cwarps * 32;
        uint64_t h[KPT][4];
        uint32_t kd[KPT];
        bool valid[KPT];
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const uint32_t qrel = j * cthreads + tid;
            valid[j] = qrel < n_keys;
            kd[j] = 0;
            h[j][0] = h[j][1] = h[j][2] = h[j][3] = 0;
            if (valid[j]) {
                const uint32_t q = key_base + qrel;
                const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hashes + 4ull * q);
                const ulonglong2 a = __ldg(hp), b = __ldg(hp + 1);
                h[j][0] = a.x; h[j][1] = a.y; h[j][2] = b.x; h[j][3] = b.y;
                kd[j] = __ldg(&kinds[q]);
            }
        }
        for (uint32_t it = 0; it < my_count; ++it) {
            const int s = it % n_stages;
            const uint32_t ph = (it / n_stages) & 1u;
            mbar_wait(&full[s], ph);
            const uint8_t* st = stages + static_cast<size_t>(s) * stage_bytes;
            const DevFilter* hdr = reinterpret_cast<const DevFilter*>(st);
            const StageHdrTail* tail = reinterpret_cast<const StageHdrTail*>(st + 3 * sizeof(DevFilter));
            const uint32_t unit = tail->unit;
            const uint64_t word_base = tail->word_base;
            const uint32_t* data32 = reinterpret_cast<const uint32_t*>(st + kProbeStageHeaderBytes);
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                // warp-uniform: does this warp's 32-key group hold any key?
                const uint32_t group_first = j * cthreads + warp * 32;
                if (group_first >= n_keys) break;
                bool res = false;
                if (valid[j]) {
                    const DevFilter f = hdr[kd[j]];
                    if (f.m == 0) {
                        res = true;  // absent filter cannot disqualify (query_exec.go:137-151)
                    } else {
                        const uint32_t* w32 = data32 + static_cast<uint32_t>(f.word_off - word_base) * 2u;
                        res = test_hashes(h[j][0], h[j][1], h[j][2], h[j][3], f.m, f.inv, f.k,
                                          [&](uint64_t idx) { return w32[static_cast<uint32_t>(idx)]; });
                    }
                }
                const uint32_t bits = __ballot_sync(0xffffffffu, res);
                if (lane == 0)
                    matrix32[static_cast<size_t>(unit) * row_words32 + ((key_base + group_first) >> 5)] = bits;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
}

static int g_max_smem_optin = 0;  // recorded for diagnostics

cudaError_t probe_staged_configure(int max_smem_optin) {
    g_max_smem_optin = max_smem_optin;
    (void)g_max_smem_optin;
    cudaError_t e;
    e = cudaFuncSetAttribute(probe_staged_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(probe_staged_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(probe_staged_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin);
    return e;
}

cudaError_t launch_probe_staged(const ProbeStagedPlan& plan, const DevFilter* d_udesc, const UnitTab* d_utab,
                                const uint64_t* d_words, const uint32_t* d_unit_list, uint32_t n_list,
                                const uint64_t* d_hashes, const uint8_t* d_kinds, uint32_t key_base, uint32_t n_keys,
                                uint32_t kind_mask, uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s) {
    if (n_list == 0 || n_keys == 0) return cudaSuccess;
    // consumer warps: as few rounds as possible, then as few idle lanes as possible
    const int cw = (plan.consumer_warps > 0 && plan.consumer_warps <= kProbeConsumerWarps) ? plan.consumer_warps
                                                                                          : kProbeConsumerWarps;
    const uint32_t cthreads = cw * 32;
    const uint32_t kpt = (n_keys + cthreads - 1) / cthreads;
    const uint32_t stage_bytes = kProbeStageHeaderBytes + plan.stage_data_bytes;
    const dim3 grid(plan.grid), block((cw + 1) * 32);
#define BSG_LAUNCH(KPT)                                                                                        \
    probe_staged_kernel<KPT><<<grid, block, plan.smem_bytes, s>>>(d_udesc, d_utab, d_words, d_unit_list, n_list, \
                                                                  d_hashes, d_kinds, key_base, n_keys, kind_mask, \
                                                                  d_matrix32, row_words32, plan.n_stages, stage_bytes)
    if (kpt <= 1) BSG_LAUNCH(1);
    else if (kpt <= 2) BSG_LAUNCH(2);
    else if (kpt <= 4) BSG_LAUNCH(4);
    else return cudaErrorInvalidValue;
#undef BSG_LAUNCH
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
                    uint32_t* __restrict__ matrix32, uint32_t row_words32) {
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
        if (key < n_keys) {
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

cudaError_t launch_probe_gather(const DevFilter* d_udesc, const uint64_t* d_words, const uint32_t* d_unit_list,
                                uint32_t n_list, const uint64_t* d_hashes, const uint8_t* d_kinds, uint32_t n_keys,
                                uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s) {
    if (n_list == 0 || n_keys == 0) return cudaSuccess;
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
        d_udesc, d_words, d_unit_list, n_list, d_hashes, d_kinds, n_keys, g_log2, chunks, d_matrix32, row_words32);
    return cudaGetLastError();
}

// -------------------------------------------------------------- tree eval ---
// One thread per unit; evaluation stack is a 64-bit bit-stack (BSG_MAX_STACK).
__global__ void __launch_bounds__(256)
tree_eval_kernel(const uint32_t* __restrict__ matrix32, uint32_t row_words32, uint64_t n_units,
                 const bsg_expr_op* __restrict__ prog, uint32_t prog_len, uint32_t* __restrict__ mask32) {
    extern __shared__ bsg_expr_op sprog[];
    for (uint32_t i = threadIdx.x; i < prog_len; i += blockDim.x) sprog[i] = prog[i];
    __syncthreads();
    const uint64_t unit = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    bool alive = false;
    if (unit < n_units) {
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
                             const bsg_expr_op* d_prog, uint32_t prog_len, uint32_t* d_mask32, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_blocks = (n_units + 255) / 256;
    if (n_blocks > 0x7fffffffull) return cudaErrorInvalidValue;
    tree_eval_kernel<<<static_cast<uint32_t>(n_blocks), 256, prog_len * sizeof(bsg_expr_op), s>>>(
        d_matrix32, row_words32, n_units, d_prog, prog_len, d_mask32);
    return cudaGetLastError();
}

__global__ void fill_mask_kernel(uint32_t* __restrict__ mask32, uint64_t n_units) {
    const uint64_t w = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t n_words = (n_units + 31) / 32;
    if (w >= n_words) return;
    const uint64_t rem = n_units - w * 32;
    mask32[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
}

cudaError_t launch_fill_mask(uint32_t* d_mask32, uint64_t n_units, cudaStream_t s) {
    if (n_units == 0) return cudaSuccess;
    const uint64_t n_words = (n_units + 31) / 32;
    fill_mask_kernel<<<static_cast<uint32_t>((n_words + 255) / 256), 256, 0, s>>>(d_mask32, n_units);
    return cudaGetLastError();
}

}  // namespace bsg
