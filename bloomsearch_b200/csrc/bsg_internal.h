// Internal (non-ABI) declarations shared by the kernels and the C-ABI layer.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/bloomgpu.h"

namespace bsg {

// Device-side filter descriptor (32 B, 16-byte aligned so it loads as 2x LDG.128
// and can be bulk-copied into a stage header).  m == 0: filter absent.
struct __align__(16) DevFilter {
    uint64_t word_off;  // offset of the filter's first uint64 word in the corpus words array (even)
    uint64_t m;         // bits
    uint64_t inv;       // floor(2^64/m) (2^64-1 for m == 1), see mod_m
    uint32_t k;         // hash functions
    uint32_t nwords;    // ceil(m/64)
};
static_assert(sizeof(DevFilter) == 32, "DevFilter must be 32 bytes");

// Per-unit staging record read by the probe producer warp (32 B).  A unit's three
// filters are stored back to back (field, token, fieldtoken), each padded to an
// even number of words so every filter starts 16-byte aligned (bulk-copy rule).
struct __align__(16) UnitTab {
    uint64_t word_base;  // first word of the unit's words (even)
    uint32_t nw[3];      // padded word count per kind (0 when absent)
    uint32_t total;      // nw[0] + nw[1] + nw[2]
    uint64_t reserved;
};
static_assert(sizeof(UnitTab) == 32, "UnitTab must be 32 bytes");

// One row per staged unit, in staged order: the probe kernel bulk-copies the whole row
// (descriptors + where the words live) into the stage header, and the 32-byte head of
// the row of the unit that will occupy the same stage next.
struct __align__(16) StageFilter {  // staged filters always have m < 2^30 (they fit shared memory)
    uint32_t m;          // bits; 0 = filter absent
    uint32_t k;          // hash functions
    uint32_t ih, il;     // hi/lo halves of floor(2^64/m), see mod_m32
    uint32_t rel_bytes;  // byte offset of the filter's words inside the stage data area
    uint32_t pad[3];
};
static_assert(sizeof(StageFilter) == 32, "StageFilter must be 32 bytes");
struct __align__(16) StageRow {
    uint32_t unit;         // global unit id (matrix row)
    uint32_t total_words;  // nw[0] + nw[1] + nw[2]
    uint64_t word_base;    // first word of the unit's words (even)
    uint32_t nw[3];        // padded word count per kind
    uint32_t pad;
    StageFilter f[3];
};
static_assert(sizeof(StageRow) == 128, "StageRow must be 128 bytes");
constexpr uint32_t kStageRowBytes = 128;
constexpr uint32_t kStageHeadBytes = 32;

// Build-side filter descriptor (over the caller's out_words layout).
struct __align__(16) BuildFilter {
    uint64_t word_off;
    uint64_t m;
    uint64_t inv;
    uint32_t k;
    uint32_t nwords;
};

inline uint64_t reciprocal(uint64_t m) {
    if (m <= 1) return ~0ULL;
    return static_cast<uint64_t>((static_cast<unsigned __int128>(1) << 64) / m);
}

constexpr int kProbeMaxStages = 16;
constexpr int kProbeSmemPrefixBytes = 256;     // 16 mbarriers (128 B) + 16 done counters (64 B) + pad
constexpr int kProbeStageHeaderBytes = 160;    // StageRow (128 B) + next unit's head (32 B)
constexpr uint32_t kProbeMaxKeysPerPass = 1024;  // one key per thread, up to 32 warps per CTA
// two-phase staged kernel (probe_staged2): stage header = StageRow + next head + result row (128 B)
// + survivor count (16 B) + survivor queue (1024 x u16); prefix = full + aready mbarriers + done counters
constexpr int kProbeStage2HeaderBytes = 160 + 128 + 16 + 2048;
constexpr int kProbe2SmemPrefixBytes = 384;

// ---- launch wrappers (defined in the kernels_*.cu files) -------------------
cudaError_t launch_hash_keys(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                             uint64_t* d_hashes, cudaStream_t s);

struct ProbeStagedPlan {
    int n_stages;
    uint32_t stage_data_bytes;  // per-stage capacity for filter words
    size_t smem_bytes;
    int grid;
    int warps;            // 0 = auto
    uint32_t stagger_ns;  // delay between the prologue's stage fills (0 = none)
    // fused hashing (two-phase kernel only): packed key bytes / offsets on the device and a scratch of
    // grid x 1024 x 32 B; nullptr = the hashes come from hash_keys_kernel
    const uint8_t* fuse_keys;
    const uint64_t* fuse_key_off;
    uint64_t* fuse_scratch;
    int pdl;              // launch with programmatic stream serialization (overlap with the predecessor's tail)
    uint32_t relax_sleep_ns;  // sleep between polls of a phase-B warp waiting for phase A (0 = none)
    int variant;          // 0 = probe_staged (one phase), 1..5 = probe_staged2 shapes (kernels_probe.cu)
};
// ---- tile ring (probe_tiles_kernel, the default staged probe) --------------------------------
// The corpus is cut into TILES, the unit of the shared-memory ring.  Two modes, fixed per corpus:
//   UNIT mode (parts = 1): a tile = up to kTileMaxUnits consecutive whole units (small units, e.g. the
//             flush-shaped 1 000-row blocks: one mbarrier wait and one refill serve several units);
//   KIND mode (parts = 2): a tile = one unit's {field, token} filters or its {fieldtoken} filter (large
//             units, e.g. merged 10 000-row blocks: half-size stages give a ring twice as deep, and the
//             batch's keys are sorted by kind so only the warps holding that kind work on the tile).
// One 512-byte TileRec per tile, bulk-copied into the stage header; the 64-byte TileFill of the tile
// that will occupy the same stage next travels with it, so a refill never waits on a global load.
constexpr uint32_t kTileMaxUnits = 8;
struct __align__(16) TileFilter {  // staged filters: m < 2^30, k < 2^8, at most 16 MB into the tile data
    uint32_t m;        // bits
    uint32_t ih, il;   // hi / lo halves of floor(2^64/m), see mod_m32
    uint32_t krel;     // (byte offset of the filter's words from the START OF THE STAGE, header included) << 8 | k
};
// An ABSENT filter (Go nil: cannot disqualify, query_exec.go:137-151) is k = 0 and reads as a one-bit filter
// (m = 1: every location reduces to bit 0) whose only word is TileRec::ones in the stage header, which is all
// ones: the first tests of a key PASS on it with no predicate, no special case and no zero-filled copy of the
// descriptor (round B1 then sets the key's bit because k <= NT); the same record fills the kinds a KIND-mode
// tile does not carry.
constexpr uint32_t kTileMaxK = 255;
#if defined(__CUDACC__)
#define BSG_HD __host__ __device__
#else
#define BSG_HD
#endif
BSG_HD inline uint32_t tile_k(uint32_t krel) { return krel & 0xffu; }
BSG_HD inline uint32_t tile_rel(uint32_t krel) { return krel >> 8; }
constexpr uint32_t kTileOnesOff = 12;   // offsetof(TileRec, ones)
inline TileFilter tile_filter_absent() { return TileFilter{1u, 0xffffffffu, 0xffffffffu, kTileOnesOff << 8}; }
struct __align__(16) TileFill {    // what the thread that (re)fills a stage needs
    uint64_t word_base;            // first word of the tile's data in the corpus words array (even)
    uint32_t data_bytes;           // all filters of the tile, 16-byte multiple
    uint32_t rec_bytes;            // bytes of the TileRec to copy: kTileRecFixedBytes + 48 * n_units
    uint16_t nb16[kTileMaxUnits][3];  // padded bytes / 16 of each filter, in storage order
};
static_assert(sizeof(TileFill) == 64, "TileFill must be 64 bytes");
struct __align__(16) TileRec {
    uint32_t n_units;     // 1..kTileMaxUnits
    uint32_t part_kinds;  // kinds this tile covers (7 in UNIT mode; 3 or 4 in KIND mode)
    uint32_t flags;       // kTileFirstPart | kTileLastPart | kTileSmallK
    uint32_t ones;        // 0xffffffff: the word an absent filter's record points at (tile_filter_absent)
    TileFill fill;
    uint32_t unit[kTileMaxUnits];         // global unit ids (matrix rows)
    TileFilter f[kTileMaxUnits][3];       // [unit in tile][kind]
    uint32_t pad1[4];
};
static_assert(sizeof(TileRec) == 512, "TileRec must be 512 bytes");
static_assert(offsetof(TileRec, ones) == kTileOnesOff, "tile_filter_absent points at TileRec::ones");
constexpr uint32_t kTileRecFixedBytes = 16 + 64 + 32;   // head + fill + unit ids; the descriptors follow
constexpr uint32_t kTileDescOff = kTileRecFixedBytes;
constexpr uint32_t kTileFirstPart = 1u, kTileLastPart = 2u, kTileSmallK = 4u;  // SmallK: some filter has k < 4
// stage = [TileRec 512][next tile's TileFill 64][pad 64][tile data]
constexpr uint32_t kTileNextFillOff = 512;
constexpr uint32_t kTileBitmapOff = 640;
inline uint32_t tile_header_bytes(uint32_t /*units_cap*/) { return kTileBitmapOff; }
// fixed shared memory of probe_tiles_kernel in front of the ring: mbarriers + list counters, slot table,
// the tile's result rows, the two survivor lists (u16 per (unit, key): worst case every key survives), hashes
constexpr uint32_t kTilesPrefixBytes = 512;
constexpr uint32_t kTilesSlotInfoBytes = 2 * kProbeMaxKeysPerPass;  // u16 per sorted key slot
inline uint32_t tiles_fixed_smem(uint32_t units_cap, uint32_t n_keys) {
    const uint32_t hash_bytes = ((n_keys + 31u) & ~31u) * 32u;
    return kTilesPrefixBytes + kTilesSlotInfoBytes + units_cap * 128u + 2u * units_cap * 2u * kProbeMaxKeysPerPass + hash_bytes;
}

struct ProbeTilesPlan {
    int n_stages;
    uint32_t stage_data_bytes;   // per-stage capacity for filter words (16-byte multiple)
    uint32_t units_cap;          // units per tile the corpus was cut for (header size)
    uint32_t parts;              // 1 = UNIT mode, 2 = KIND mode
    size_t smem_bytes;
    int grid;
    int shape;                   // index into the compiled shapes (kernels_probe_tiles.cu)
    int pdl;
    const uint8_t* fuse_keys;    // fused hashing: packed key bytes / offsets on the device (else d_hashes)
    const uint64_t* fuse_key_off;
};
cudaError_t probe_tiles_configure(int max_smem_optin);
int probe_tiles_n_shapes();
int probe_tiles_threads(int shape);        // threads per CTA of a compiled shape; 1024 / threads CTAs share an SM
const char* probe_tiles_shape_name(int shape);
cudaError_t launch_probe_tiles(const ProbeTilesPlan& plan, const TileRec* d_tiles, uint32_t n_items,
                               const uint32_t* d_n_items, const uint64_t* d_words, const uint64_t* d_hashes,
                               const uint16_t* d_slotinfo, uint32_t key_base, uint32_t n_keys, uint32_t kind_mask,
                               uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s, uint64_t* d_trace = nullptr,
                               uint32_t trace_slots = 0);
// hierarchical probe: keep the items (tiles, or pairs of tiles in KIND mode) with a unit whose parent survived
cudaError_t launch_compact_tiles(const TileRec* d_tiles, uint32_t n_items, uint32_t parts, const uint32_t* d_parent,
                                 const uint32_t* d_parent_mask32, TileRec* d_out, uint32_t* d_n_out, cudaStream_t s);

cudaError_t probe_staged_configure(int max_smem_optin);
cudaError_t launch_probe_staged(const ProbeStagedPlan& plan, const StageRow* d_stab, uint32_t n_list,
                                const uint64_t* d_words, const uint64_t* d_hashes, const uint8_t* d_kinds,
                                uint32_t key_base, uint32_t n_keys, uint32_t kind_mask, uint32_t* d_matrix32,
                                uint32_t row_words32, cudaStream_t s, uint64_t* d_trace = nullptr,
                                uint32_t trace_slots = 0, const uint32_t* d_n_list = nullptr);
cudaError_t launch_probe_gather(const DevFilter* d_udesc, const uint64_t* d_words, const uint32_t* d_unit_list,
                                uint32_t n_list, const uint64_t* d_hashes, const uint8_t* d_kinds, uint32_t n_keys,
                                uint32_t* d_matrix32, uint32_t row_words32, cudaStream_t s,
                                const uint32_t* d_parent = nullptr, const uint32_t* d_parent_mask32 = nullptr,
                                const bsg_expr_op* d_prog = nullptr, uint32_t prog_len = 0);   // prog: short-circuit form (mask-only callers)
cudaError_t launch_tree_eval(const uint32_t* d_matrix32, uint32_t row_words32, uint64_t n_units,
                             const bsg_expr_op* d_prog, uint32_t prog_len, uint32_t* d_mask32, cudaStream_t s,
                             const uint32_t* d_parent = nullptr, const uint32_t* d_parent_mask32 = nullptr);
cudaError_t launch_tree_eval_multi(const uint32_t* d_matrix32, uint32_t row_words32, uint64_t n_units,
                                   const bsg_expr_op* d_prog, const uint32_t* d_prog_begin, uint32_t n_queries,
                                   uint32_t* d_masks32, uint64_t mask_words32, const uint32_t* d_bad32,
                                   cudaStream_t s);
cudaError_t launch_parent_mask(uint32_t* d_mask32, uint64_t n_units, const uint32_t* d_parent,
                               const uint32_t* d_parent_mask32, cudaStream_t s);
cudaError_t launch_compact_rows(const StageRow* d_stab, uint32_t n_rows, const uint32_t* d_parent,
                                const uint32_t* d_parent_mask32, StageRow* d_out, uint32_t* d_n_out, cudaStream_t s);
cudaError_t launch_fill_mask(uint32_t* d_mask32, uint64_t n_units, cudaStream_t s);
cudaError_t launch_mask_andnot(uint32_t* d_mask32, const uint32_t* d_bad32, uint64_t n_units, cudaStream_t s);

cudaError_t build_configure(int max_smem_optin);
cudaError_t launch_build(const uint8_t* d_keys, const uint64_t* d_key_off, const uint64_t* d_group_begin,
                         uint32_t n_groups, const uint32_t* d_group_filter, const uint32_t* d_group_filter2,
                         const BuildFilter* d_filters, uint64_t* d_out_words, uint32_t smem_cap_bytes,
                         cudaStream_t s);

cudaError_t launch_build_ft(const uint8_t* d_strings, const uint64_t* d_str_off, const uint32_t* d_pair_path,
                            const uint32_t* d_pair_token, const uint64_t* d_group_begin, uint32_t n_groups,
                            const uint32_t* d_group_filter, const uint32_t* d_group_filter2,
                            const BuildFilter* d_filters, uint64_t* d_out_words, uint32_t smem_cap_bytes,
                            cudaStream_t s);

cudaError_t launch_count_distinct(const uint8_t* d_keys, const uint64_t* d_key_off, uint64_t n_keys,
                                  const uint64_t* d_group_begin, uint32_t n_groups, const uint32_t* d_group_parent,
                                  uint32_t n_parents, void* d_scratch, unsigned long long* d_group_counts,
                                  unsigned long long* d_parent_counts, cudaStream_t s);
size_t count_distinct_scratch_bytes(uint64_t n_keys);

cudaError_t launch_repack(const uint64_t* d_src, const uint64_t* d_src_off, const DevFilter* d_udesc,
                          uint64_t n_filters, uint64_t* d_dst, int big_endian, cudaStream_t s);

// Section parsing on the device (file_format.go:392-448): per unit, verify
// CRC32C, parse flags / length prefixes / (m,k,bitlen) headers.
struct SectionInfo {          // written by the parse kernel, one per unit
    int32_t status;           // 0 ok, else negative detail code
    uint32_t present;         // flags byte
    uint64_t m[3], k[3];      // per present filter
    uint64_t words_byte_off[3];  // offset of the first BE word inside `sections`
};
cudaError_t sections_configure();
cudaError_t launch_parse_sections(const uint8_t* d_sections, const uint64_t* d_sec_off, uint64_t n_units,
                                  int verify_crc, SectionInfo* d_info, cudaStream_t s);
cudaError_t launch_repack_sections(const uint8_t* d_sections, const SectionInfo* d_info, const DevFilter* d_udesc,
                                   uint64_t n_units, uint64_t* d_dst, cudaStream_t s);

cudaError_t launch_or_words(uint64_t* d_dst, const uint64_t* d_src, uint64_t n_words, cudaStream_t s);

}  // namespace bsg
