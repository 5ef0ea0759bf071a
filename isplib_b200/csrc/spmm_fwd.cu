// spmm_fwd.cu -- CSR SpMM forward kernels (sum / mean / max / min) for sm_100a.
//
// Replaces the body of fusedMM_csr as called at /root/reference/csrc/fusedmm.cpp:198
// (declared csrc/fusedMM.h:77-99; real body un-vendored, configure:2-7).
//
// Design (DESIGN.md "Kernels"): the path is an irregular gather-reduce bound by
// HBM/L2 bandwidth -- no tensor cores.  Work is cut into row-aligned SEGMENTS of at
// most seg_len stored entries (the plan, graph_ops.cu); one warp owns one segment.
//   * col/val are read 32 at a time, coalesced, streaming (ld.global.cs), one chunk
//     prefetched ahead, and broadcast with warp shuffles;
//   * a lane group of G lanes covers one K tile of a dense row with 16-byte (float4)
//     loads, 32/G entries are gathered per warp instruction and U of them are kept in
//     flight per lane group before the first use (memory-level parallelism);
//   * fp32 accumulators (and the running arg-edge for max/min) stay in registers;
//   * rows that fit one segment are finalised in place (mean divide, arg widening to
//     int64, optional merge with a previous block's result); longer rows write one
//     partial per segment and the last segment warp to arrive (an atomic ticket per row)
//     merges them in segment order inside the same launch, so the result is
//     deterministic and max/min/arg stay bit-exact however a row is split;
//   * blockIdx.y walks K tiles slowest, so with a narrow tile the whole grid sweeps
//     one [N, tile] column slab of X at a time and the slab stays L2-resident.
#include "common.cuh"

namespace isplib {

// ------------------------------------------------------------------------------------
// variant table
// ------------------------------------------------------------------------------------
static const VariantDesc kVariants[] = {
    {"seg/w8/u4/kfull", 0, 8, 4, 0, 0},
    {"seg/w4/u4/kfull", 0, 4, 4, 0, 0},
    {"seg/w8/u8/kfull", 0, 8, 8, 0, 0},
    {"seg/w4/u8/kfull", 0, 4, 8, 0, 0},
    {"seg/w8/u4/kt128", 0, 8, 4, 128, 0},
    {"seg/w4/u4/kt128", 0, 4, 4, 128, 0},
    {"seg/w8/u4/kt64", 0, 8, 4, 64, 0},
    {"seg/w4/u4/kt64", 0, 4, 4, 64, 0},
    {"seg/w8/u4/kt32", 0, 8, 4, 32, 0},
    {"seg/w4/u4/kt32", 0, 4, 4, 32, 0},
    {"seg/w8/u8/kt64", 0, 8, 8, 64, 0},
    {"seg/w4/u2/kfull", 0, 4, 2, 0, 0},
    {"seg/w4/u2/kt64", 0, 4, 2, 64, 0},
    {"seg/w8/u2/kt64", 0, 8, 2, 64, 0},
    // method 3: 32-byte gathers (LDG.E.256), rows 32-byte aligned and padded to a multiple of 8.
    // A bare gather loop gains +25 % from 256-bit loads (profiles/r1_l2probe.txt), but in this
    // kernel 4 x 32 B in flight per lane cost 101 registers (20 warps/SM) and U=2 loses the gain:
    // 5.1 ms vs 4.2 ms on Reddit-shape K=128.  Selectable by id / ISPLIB_B200_TUNE_ALL=1 only.
    {"seg256/w4/u4/kfull", 3, 4, 4, 0, 0},
    {"seg256/w4/u2/kfull", 3, 4, 2, 0, 0},
    {"seg256/w4/u4/kt128", 3, 4, 4, 128, 0},
    {"seg256/w4/u4/kt64", 3, 4, 4, 64, 0},
    {"seg256/w4/u2/kt64", 3, 4, 2, 64, 0},
    {"seg256/w8/u2/kfull", 3, 8, 2, 0, 0},
    // method 5: lean kernel with 32-byte gathers inside 64 (sum) / 80 (max, min) registers;
    // full tiles of 32/64/128/256 floats only
    {"lean256/w4/kfull", 5, 4, 4, 0, 0},
    {"lean256/w4/kt128", 5, 4, 4, 128, 0},
    {"lean256/w4/kt64", 5, 4, 4, 64, 0},
    // sequential K tiles: one launch per tile, so only ONE [N, tile] slab of X is live in L2 at a
    // time (with grid.y tiles the tail of tile t overlaps the head of tile t+1)
    {"seg/w4/u4/kt64/seq", 0, 4, 4, 64, 1},
    {"lean256/w4/kt64/seq", 5, 4, 4, 64, 1},
    {"lean256/w4/kt128/seq", 5, 4, 4, 128, 1},
    // method 1: TMA bulk-copy gather through a per-warp shared-memory ring; `unroll` = stages
    // (measured 3x slower than the LDG gather for 256-512 B rows -- profiles/r1_kbench_bulk.txt:
    // the copy engine retires one small bulk request per ~14-30 cycles per SM -- so the on-device
    // selection skips these unless ISPLIB_B200_TUNE_BULK=1; they stay selectable by id)
    {"bulk/w4/s3/kfull", 1, 4, 3, 0, 0},
    {"bulk/w4/s3/kt64", 1, 4, 3, 64, 0},
    {"bulk/w8/s3/kt64", 1, 8, 3, 64, 0},
    // method 6: the lean body with 16-byte gathers (48 / 56 registers, 36-40 warps/SM); new
    // entries go to the end so that variant ids quoted in profiles/ stay valid
    {"lean128/w4/kfull", 6, 4, 4, 0, 0},
    {"lean128/w4/kt64", 6, 4, 4, 64, 0},
    {"lean128/w4/kt64/seq", 6, 4, 4, 64, 1},
    // method 7: REFERENCE-ORDER sum.  The general kernel with scalar lanes (one lane group = the
    // whole warp, so a warp takes ONE entry per step) adds a row's terms in CSR order with one FMA
    // each -- exactly the recurrence of the CPU kernel it replaces (z[i,k] = fma(a_e, y[col_e,k],
    // z[i,k]), e ascending; oracle/fusedmm_oracle.c).  With a plan whose seg_len >= the maximum
    // degree (no row is split) sum / mean are BIT-IDENTICAL to the CPU path, which pins that the
    // fast variants differ from it by re-association only.  4x the load instructions: a parity
    // mode (never auto-selected, not in the autotune set), not a performance variant.
    {"ordered/w4/u4/kfull", 7, 4, 4, 0, 0},
};
int variant_count() { return (int)(sizeof(kVariants) / sizeof(kVariants[0])); }
const VariantDesc* variant_desc(int v) {
    return (v >= 0 && v < variant_count()) ? &kVariants[v] : nullptr;
}

// ------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------
struct TileShape { int vec, g, lpl, tile_w, ntiles; };
typedef void (*SegKernel)(const SpmmParams);
constexpr int kBulkSE = 16;   // rows per stage of the bulk-copy ring (mirrors spmm_kernels.cuh)
// one translation unit per reduction (spmm_inst_*.cu)
SegKernel seg_kernel_sum(const TileShape&, int, bool);
SegKernel seg_kernel_max(const TileShape&, int, bool);
SegKernel seg_kernel_min(const TileShape&, int, bool);
SegKernel bulk_kernel_sum(const TileShape&, int);
SegKernel bulk_kernel_max(const TileShape&, int);
SegKernel bulk_kernel_min(const TileShape&, int);
SegKernel lean256_kernel_sum(int g, bool ragged, bool noval);
SegKernel lean256_kernel_max(int g, bool ragged, bool noval);
SegKernel lean256_kernel_min(int g, bool ragged, bool noval);
SegKernel lean128_kernel_sum(int g, bool ragged, bool noval);
SegKernel lean128_kernel_max(int g, bool ragged, bool noval);
SegKernel lean128_kernel_min(int g, bool ragged, bool noval);

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 16-byte gathers need 16-byte aligned x rows that own their padding up to roundup4(K); the
// OUTPUT side does not matter (stores fall back to scalars per vector, finalize_store)
static int pick_vec(int64_t k, int64_t ldx, const void* x) {
    if (ldx % 4 == 0 && ldx >= ((k + 3) & ~(int64_t)3) && aligned16(x)) return 4;
    return 1;
}
// 32-byte (LDG.E.256) gathers: 32-byte aligned rows that own their padding up to roundup8(K)
static bool vec8_ok(int64_t k, int64_t ldx, const void* x) {
    return ldx % 8 == 0 && ldx >= ((k + 7) & ~(int64_t)7) && (reinterpret_cast<uintptr_t>(x) & 31u) == 0;
}

static TileShape pick_shape(int vec, int64_t k, int kt) {
    TileShape t;
    t.vec = vec;
    const int max_tile = vec == 8 ? 512 : vec * 32 * 4;  // G=32, LPL=4 (LPL=2 for 32-byte vectors)
    if (vec > 1) k = (k + vec - 1) / vec * vec;
    int64_t tw = (kt <= 0 || kt >= k) ? k : kt;
    if (tw > max_tile) tw = max_tile;
    if (vec > 1) tw = (tw + vec - 1) / vec * vec;
    const int tv = (int)((tw + vec - 1) / vec);  // vectors per tile
    if (vec == 8 && tv <= 4)       { t.g = 4;  t.lpl = 1; }
    else if (vec == 8 && tv <= 8)  { t.g = 8;  t.lpl = 1; }
    else if (vec == 8 && tv <= 16) { t.g = 16; t.lpl = 1; }
    else if (vec == 8 && tv <= 32) { t.g = 32; t.lpl = 1; }
    else if (vec == 8)             { t.g = 32; t.lpl = 2; }
    else if (vec == 4 && tv <= 8)  { t.g = 8;  t.lpl = 1; }
    else if (vec == 4 && tv <= 16) { t.g = 16; t.lpl = 1; }
    else if (tv <= 32)             { t.g = 32; t.lpl = 1; }
    else if (tv <= 64)             { t.g = 32; t.lpl = 2; }
    else                           { t.g = 32; t.lpl = 4; }
    t.tile_w = (int)tw;
    t.ntiles = (int)((k + tw - 1) / tw);
    return t;
}

static size_t bulk_smem_bytes(const TileShape& t, int warps, int stages) {
    return (size_t)warps * stages * kBulkSE * (size_t)(t.g * t.lpl * 16) + (size_t)warps * stages * 8 + 16;
}
constexpr size_t kMaxDynSmem = 227 * 1024;

bool spmm_variant_supported(int variant, int reduce, int64_t k, int64_t ldx, int64_t ldo,
                            const void* x, const void* out) {
    const VariantDesc* d = variant_desc(variant);
    if (!d || reduce < 0 || reduce > 3 || k <= 0) return false;
    (void)ldo; (void)out;
    const int vec = pick_vec(k, ldx, x);
    if (d->method == 5) {   // lean 32-byte kernel: whole tiles of 32..256 floats
        if (!vec8_ok(k, ldx, x)) return false;
        const int64_t tw = d->kt > 0 ? d->kt : k;
        if (d->kt > 0 && d->kt >= k) return false;
        // an untiled launch may end inside a vector (K = 47 in rows padded to 48): the loads stay
        // inside the padded row, finalize_store writes only the valid columns
        return tw >= 32 && tw <= 256 && (d->kt == 0 || tw % 8 == 0) && k % tw == 0;
    }
    if (d->method == 6) {   // lean 16-byte kernel: whole tiles of 16..128 floats
        if (vec != 4) return false;
        const int64_t tw = d->kt > 0 ? d->kt : k;
        if (d->kt > 0 && d->kt >= k) return false;
        return tw >= 16 && tw <= 128 && tw % 4 == 0 && k % tw == 0;
    }
    if (d->method == 7) return true;   // scalar lanes: any K, any alignment
    if (d->method == 3) {   // 32-byte gathers
        if (!vec8_ok(k, ldx, x)) return false;
        if (d->kt > 0 && (d->kt >= k || d->kt % 8 != 0)) return false;
        return true;
    }
    if (d->method == 1) {
        // bulk copies need 16-byte aligned rows and sizes, and the ring must fit shared memory
        if (vec != 4) return false;
        const TileShape t = pick_shape(vec, k, d->kt);
        if (bulk_smem_bytes(t, d->warps, d->unroll) > kMaxDynSmem) return false;
    }
    if (d->kt > 0) {
        if (d->kt >= k) return false;            // same as kfull: do not time it twice
        if (vec == 4 && d->kt % 4 != 0) return false;
    }
    return true;
}

static int find_variant(int method, int warps, int unroll, int kt, int seq = 0) {
    for (int v = 0; v < variant_count(); ++v)
        if (kVariants[v].method == method && kVariants[v].seq == seq && kVariants[v].warps == warps &&
            kVariants[v].unroll == unroll && kVariants[v].kt == kt) return v;
    return -1;
}

// Shape-only default (the op layer replaces it by the measured winner when autotuning is on).
// Measured on B200 (profiles/r1_kbench_*.txt): 4 warps/CTA and 4 gathers in flight win
// everywhere; sum/mean on 32-byte-aligned rows of 32..256 floats (any multiple of 8) prefer the
// lean 256-bit kernel, max/min only when x is HBM-resident;
// a K tile pays off only when it makes an [n, tile] slab of x L2-resident (126 MB L2) while x
// itself is not (Reddit-shape K>=128 -> 64-wide); when nothing can be resident
// (products/amazon shapes) 128-wide tiles are marginally ahead for K > 128.
int spmm_variant_default(int reduce, int64_t n, int64_t k, int64_t ldx, int64_t ldo, const void* x,
                         const void* out, double avg_degree) {
    (void)avg_degree;
    const double MB = 1024.0 * 1024.0;
    const double x_bytes = (double)n * (double)k * 4.0;
    const bool slab64 = (double)n * 64.0 * 4.0 <= 64.0 * MB;
    const bool additive = (reduce == ISPLIB_REDUCE_SUM || reduce == ISPLIB_REDUCE_MEAN);
    // max / min: since round 2 the 32-byte lean body is compiled for 32 warps/SM (64 registers) and
    // leads from K = 64 up (Reddit-shape K=128 4.56 vs 5.0 ms, K=256 9.3 vs 10.1 ms for seg/*;
    // profiles/r2_kbench_max_minb.txt); narrow rows (K < 64) stay with seg/* (K=32: 1.34 vs 1.46 ms)
    if (additive || k >= 64 || x_bytes > 512.0 * MB) {
        int v = -1;
        if (x_bytes > 96.0 * MB && k > 64 && slab64) v = find_variant(5, 4, 4, 64, additive || k > 128 ? 1 : 0);   // 64-wide slabs of x, L2-resident
        else v = find_variant(5, 4, 4, 0);
        if (v >= 0 && spmm_variant_supported(v, reduce, k, ldx, ldo, x, out)) return v;
        v = find_variant(5, 4, 4, 0);
        if (v >= 0 && spmm_variant_supported(v, reduce, k, ldx, ldo, x, out)) return v;
        // rows that are only 16-byte aligned (K = 100): the same body with 16-byte gathers,
        // 8 % ahead of seg/* on products-shape K=100, never behind it for sum
        v = find_variant(6, 4, 4, 0);
        if (v >= 0 && spmm_variant_supported(v, reduce, k, ldx, ldo, x, out)) return v;
    }
    int kt = 0;
    if (x_bytes > 96.0 * MB) {
        if (k > 128 && (double)n * 128.0 * 4.0 <= 64.0 * MB) kt = 128;
        else if (k > 64 && slab64) kt = 64;
        else if (k > 128) kt = 128;
    }
    int v = find_variant(0, 4, 4, kt);
    if (v < 0 || !spmm_variant_supported(v, reduce, k, ldx, ldo, x, out)) v = find_variant(0, 4, 4, 0);
    return v < 0 ? 0 : v;
}

int launch_spmm(int reduce, const SpmmParams& base, int64_t nnz, int variant, cudaStream_t stream) {
    (void)nnz;
    const VariantDesc* d = variant_desc(variant);
    if (!d) return ISPLIB_NO_OPT_IMPL;
    if (base.m == 0 || base.k == 0) return ISPLIB_SUCCESS;

    SpmmParams p = base;
    int vec = pick_vec(p.k, p.ldx, p.x);
    if (d->method == 7) vec = 1;
    if (d->method == 3) {
        if (!vec8_ok(p.k, p.ldx, p.x)) return ISPLIB_NO_OPT_IMPL;
        vec = 8;
    }
    TileShape t = pick_shape(vec, p.k, d->kt);
    if (d->method == 5) {
        if (!vec8_ok(p.k, p.ldx, p.x)) return ISPLIB_NO_OPT_IMPL;
        const int tw = d->kt > 0 ? d->kt : p.k;
        int g = 4;
        while (g * 8 < tw) g <<= 1;               // lanes per row: 4, 8, 16 or 32
        t.vec = 8; t.g = g; t.lpl = 1; t.tile_w = tw; t.ntiles = p.k / tw;
        vec = 8;
    }
    if (d->method == 6) {
        if (vec != 4) return ISPLIB_NO_OPT_IMPL;
        const int tw = d->kt > 0 ? d->kt : p.k;
        int g = 4;
        while (g * 4 < tw) g <<= 1;
        t.vec = 4; t.g = g; t.lpl = 1; t.tile_w = tw; t.ntiles = p.k / tw;
    }
    const int keff = vec > 1 ? (p.k + vec - 1) / vec * vec : p.k;
    p.kp = (p.k + 7) & ~7;
    p.vec_store = (p.k % 4 == 0 && p.ldo % 4 == 0 && aligned16(p.out) && (!p.arg_out || aligned16(p.arg_out))) ? 1 : 0;
    p.tile_w = t.tile_w;
    const int op = (reduce == ISPLIB_REDUCE_MAX) ? OP_MAX : (reduce == ISPLIB_REDUCE_MIN ? OP_MIN : OP_SUM);

    // every tile (incl. the last one) fills all G*LPL vector slots of a lane group?
    const bool partial = (t.tile_w != t.g * t.lpl * t.vec) || (keff % t.tile_w != 0);
    SegKernel kern = nullptr;
    size_t smem = 0;
    if (d->method == 5) {
        const bool ragged = (t.g * 8 != t.tile_w);
        const bool noval = (p.val == nullptr);
        kern = op == OP_SUM ? lean256_kernel_sum(t.g, ragged, noval)
                            : (op == OP_MAX ? lean256_kernel_max(t.g, ragged, noval) : lean256_kernel_min(t.g, ragged, noval));
    } else if (d->method == 6) {
        const bool ragged = (t.g * 4 != t.tile_w);
        const bool noval = (p.val == nullptr);
        kern = op == OP_SUM ? lean128_kernel_sum(t.g, ragged, noval)
                            : (op == OP_MAX ? lean128_kernel_max(t.g, ragged, noval) : lean128_kernel_min(t.g, ragged, noval));
    } else if (d->method == 1) {
        if (op == OP_SUM) kern = bulk_kernel_sum(t, d->unroll);
        else if (op == OP_MAX) kern = bulk_kernel_max(t, d->unroll);
        else kern = bulk_kernel_min(t, d->unroll);
        if (!kern) return ISPLIB_NO_OPT_IMPL;
        smem = bulk_smem_bytes(t, d->warps, d->unroll);
        if (smem > kMaxDynSmem) return ISPLIB_NO_OPT_IMPL;
        if (smem > 48 * 1024)
            ISPLIB_CUDA_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        if (op == OP_SUM) kern = seg_kernel_sum(t, d->unroll, partial);
        else if (op == OP_MAX) kern = seg_kernel_max(t, d->unroll, partial);
        else kern = seg_kernel_min(t, d->unroll, partial);
    }
    if (!kern) return ISPLIB_NO_OPT_IMPL;
    // fused all-gather: only the lean kernels carry the copy role, and the grid must be ONE launch
    if (p.gather.copy_ctas > 0 && ((d->method != 5 && d->method != 6) || d->seq)) return ISPLIB_NO_OPT_IMPL;
    if (p.gather.copy_ctas > 0 && p.gather.tile_vec4 < 0) {
        // tile mode: the arrival groups are this variant's K tiles (whole rows when untiled)
        if (t.ntiles > kMaxArrivalGroups) return ISPLIB_NO_OPT_IMPL;
        p.gather.n_groups = t.ntiles;
        p.gather.tile_vec4 = t.ntiles > 1 ? t.tile_w / 4 : p.gather.row_vec4;
        if (t.ntiles > 1 && t.tile_w % 4 != 0) return ISPLIB_NO_OPT_IMPL;
    }

    const int warps = d->warps;
    const dim3 block(warps * 32);
    const dim3 grid((unsigned)((p.num_items + warps - 1) / warps) + (unsigned)max(p.gather.copy_ctas, 0), (unsigned)t.ntiles);
    if (t.ntiles > 65535) return ISPLIB_NO_OPT_IMPL;
    if (p.num_split_rows > 0) {
        // arrival counters of the split rows, one set per K tile (self-resetting, cleared
        // anyway so an aborted launch cannot poison the next one)
        if ((size_t)t.ntiles * (size_t)p.ticket_stride > (size_t)p.ticket_capacity) return ISPLIB_NOT_ENOUGH_MEM;
        ISPLIB_CUDA_TRY(cudaMemsetAsync(p.row_ticket, 0, (size_t)t.ntiles * (size_t)p.ticket_stride * sizeof(int), stream));
    }
    p.tile_base = 0;
    if (d->seq && t.ntiles > 1) {
        const dim3 grid1(grid.x, 1);
        for (int tile = 0; tile < t.ntiles; ++tile) {
            p.tile_base = tile;
            kern<<<grid1, block, smem, stream>>>(p);
            ISPLIB_LAUNCH_CHECK();
        }
        return ISPLIB_SUCCESS;
    }
    kern<<<grid, block, smem, stream>>>(p);
    ISPLIB_LAUNCH_CHECK();

    return ISPLIB_SUCCESS;
}

}  // namespace isplib
