// common.cuh -- shared declarations of the isplib_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/isplib_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "isplib_b200 is written for sm_100a (B200) only"
#endif

namespace isplib {

#define ISPLIB_CUDA_TRY(expr)                                            \
    do {                                                                 \
        cudaError_t _e = (expr);                                         \
        if (_e != cudaSuccess) return ISPLIB_CUDA_ERROR_BASE + (int)_e;  \
    } while (0)

#define ISPLIB_LAUNCH_CHECK()                                            \
    do {                                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return ISPLIB_CUDA_ERROR_BASE + (int)_e;  \
    } while (0)

constexpr int kDefaultSegLen = 512;  // Reddit-shape sweep (gpurun_out kbench_r1_seglen): 512-1024 best
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kMinTileW = 8;  // narrowest K tile any variant uses (sizes the ticket array)

// workspace carving shared by the size query and the launcher
struct WorkspaceLayout {
    size_t off_part_val, off_part_arg, off_ticket, total;
    int ticket_stride, ticket_capacity;
};
static inline size_t align_up_(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline WorkspaceLayout workspace_layout(int64_t num_split_items, int64_t k, bool is_arg) {
    WorkspaceLayout W;
    const size_t part = align_up_((size_t)num_split_items * (size_t)((k + 7) / 8 * 8) * 4, 256);
    size_t o = 0;
    W.off_part_val = o; o += part;
    W.off_part_arg = o; if (is_arg) o += part;
    W.ticket_stride = (int)(num_split_items / 2 + 1);
    const int64_t max_tiles = (k + kMinTileW - 1) / kMinTileW;
    W.ticket_capacity = (int)((int64_t)W.ticket_stride * (max_tiles > 0 ? max_tiles : 1));
    W.off_ticket = o; o += align_up_((size_t)W.ticket_capacity * 4, 256);
    W.total = o + 256;
    return W;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- plan layout inside the caller's device buffer ---------------------------------
// [counters: 8 x int64][seg_off: m+1][part_off: m+1][item_desc: wmax x int4][temp...]
struct PlanLayout {
    size_t off_counters, off_seg_off, off_part_off, off_item_desc, off_temp;
    size_t persistent_bytes;  // everything the SpMM kernels read
    int64_t wmax;
};

static inline int32_t effective_seg_len(int32_t seg_len) {
    if (seg_len <= 0) return kDefaultSegLen;
    // multiple of 32 so a warp's 32-wide index chunks stay aligned to the segment start
    int32_t s = (seg_len + 31) / 32 * 32;
    return s;
}

static inline PlanLayout plan_layout(int64_t m, int64_t nnz, int32_t seg_len) {
    PlanLayout L;
    const int32_t S = effective_seg_len(seg_len);
    L.wmax = m + nnz / S + 1;
    size_t o = 0;
    L.off_counters = o;   o = align_up(o + 8 * sizeof(int64_t), 256);
    L.off_seg_off = o;    o = align_up(o + (size_t)(m + 1) * 4, 256);
    L.off_part_off = o;   o = align_up(o + (size_t)(m + 1) * 4, 256);
    L.off_item_desc = o;  o = align_up(o + (size_t)L.wmax * 16, 256);
    L.persistent_bytes = o;
    L.off_temp = o;
    return L;
}

enum PlanCounter { PC_ITEMS = 0, PC_SPLIT_ROWS = 1, PC_SPLIT_ITEMS = 2, PC_MAX_DEG = 3, PC_EMPTY = 4 };

// ---- fused all-gather of x (row-partitioned multi-GPU forward) -------------------------
// The lean SpMM kernels move the slices of x over NVLink THEMSELVES while they multiply: the first
// `copy_ctas` CTAs of the grid PUSH this rank's own slice into every peer's gathered x (posted
// stores into peer-mapped memory: no round trip per request, unlike pulling) and then bump that
// peer's arrival counter; the other CTAs multiply, each work item starting as soon as the arrival
// counter of ITS group says the rows it gathers have landed.  One launch = all-gather + SpMM, the
// transfer of group g + 1 overlaps the multiply of group g, no NCCL call, no barrier kernel.
//   arrival groups = the K TILES of the launch (tile mode: rows stay whole, any plan), or
//                  = column owners (owner mode: grouped plan, rows split per group).
//   flow control   = credits: a rank announces "I have started step e" to every peer at kernel start;
//                    a peer pushes step e + 1 into this rank's buffer of that parity only after that
//                    (double-buffered x, so step e + 1 never overwrites rows step e still reads).
constexpr int kMaxPeers = 15;           // world <= 16
constexpr int kMaxArrivalGroups = 8;
struct GatherParams {
    float* dst[kMaxPeers];              // peer q's gathered x (this step's parity) at MY slice's rows, peer-mapped
    unsigned* peer_arrive[kMaxPeers];   // peer q's arrival counters of this parity: [kMaxArrivalGroups]
    unsigned* peer_credit[kMaxPeers];   // peer q's credit words: we store `epoch` at [my_rank] at kernel start
    int dst_group[kMaxPeers];           // owner mode: the arrival group my slice belongs to AT that peer
    int dst_rank[kMaxPeers];            // the peer's rank (index into credit_local)
    int group_item_end[kMaxArrivalGroups];      // owner mode: items [end[g-1], end[g]) gather from group g
    unsigned arrive_target[kMaxArrivalGroups];  // value arrive_local[g] reaches when group g of THIS step has landed
    const float* own;                   // my own slice inside my gathered x (what gets pushed)
    unsigned* arrive_local;             // my arrival counters of this parity, bumped by the peers' copy CTAs
    unsigned* credit_local;             // [world]: credit_local[q] >= epoch - 1 <=> peer q's buffer of this parity is free
    unsigned* status;                   // set to 1 if a wait timed out (a peer never arrived)
    long long slice_vec4;               // 16-byte vectors per slice
    long long slice_rows;
    unsigned epoch;                     // 1, 2, 3 ... per step on this buffer set
    int n_dst, n_groups, copy_ctas;     // copy_ctas == 0: plain kernel, nothing here is touched
    int my_rank;
    int tile_vec4;                      // tile mode: 16-byte vectors per row and K tile; 0 = owner mode
    int row_vec4;                       // 16-byte vectors per row of x (ldx / 4)
    int phase;                          // 0 fused; 1 push only; 2 multiply only (still waits) -- 1/2: single-GPU emulation
};

// ---- forward kernel parameter block --------------------------------------------------
struct SpmmParams {
    const int32_t* __restrict__ rowptr;
    const int32_t* __restrict__ col;
    const float* __restrict__ val;      // nullable
    const float* __restrict__ x;
    float* __restrict__ out;
    long long* __restrict__ arg_out;    // nullable
    const int32_t* __restrict__ seg_off;
    const int32_t* __restrict__ part_off;
    const int4* __restrict__ item_desc;   // {row, eb, ee, partial slot | -1} per work item
    float* __restrict__ part_val;       // [num_split_items, k]
    int32_t* __restrict__ part_arg;     // [num_split_items, k] (max/min)
    int* __restrict__ row_ticket;       // [ntiles, ticket_stride] arrival counters of split rows
    const float* __restrict__ row_div;  // nullable
    const int32_t* __restrict__ edge_ids;  // nullable
    long long ldx, ldo, arg_sentinel;
    int m, k, tile_w, num_items, num_split_rows, seg_len;
    int ticket_stride, ticket_capacity;   // ints per K tile / ints available
    int tile_base;  // K tile index of blockIdx.y == 0 (sequential-tile launches)
    int kp;         // row stride of the partial buffers = roundup8(k)
    int vec_store;  // 1: out (and arg_out) rows take aligned 16-byte stores
    int flags;      // ISPLIB_FLAG_*
    int div_mode;   // 0 none, 1 by max(deg,1), 2 by row_div[]
    // fused caller epilogue + auxiliary max/min outputs (isplib_b200_epilogue)
    const float* __restrict__ bias;     // [k] or null
    const float* __restrict__ addend;   // [m, k], row stride ld_addend, or null
    long long ld_addend;
    float addend_scale;
    int has_epilogue;                   // bias || addend || ISPLIB_FLAG_RELU
    int32_t* __restrict__ arg_col;      // [m, k] stride ldo or null: col[arg] (-1 where no entry won)
    float* __restrict__ arg_val;        // [m, k] stride ldo or null: val[arg]
    GatherParams gather;                // fused all-gather of x (copy_ctas == 0: off)
};

enum Op { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };

struct VariantDesc {
    const char* name;
    int method;   // 0 = warp-per-segment register gather
    int warps;    // warps per CTA
    int unroll;   // gathers in flight per lane group
    int kt;       // K tile width in elements, 0 = widest the lane mapping allows
    int seq;      // 1: one launch per K tile (stream-ordered) instead of grid.y tiles
};

int variant_count();
const VariantDesc* variant_desc(int v);

// implemented in spmm_fwd.cu
int launch_spmm(int reduce, const SpmmParams& base, int64_t nnz, int variant, cudaStream_t stream);
bool spmm_variant_supported(int variant, int reduce, int64_t k, int64_t ldx, int64_t ldo,
                            const void* x, const void* out);
int spmm_variant_default(int reduce, int64_t n, int64_t k, int64_t ldx, int64_t ldo, const void* x,
                         const void* out, double avg_degree);

}  // namespace isplib
