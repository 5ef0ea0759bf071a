/*
 * isplib_b200.h -- C ABI of the B200-native FusedMM CSR SpMM path.
 *
 * This header is the drop-in boundary.  It replaces the single C entry point the
 * reference's hot path bottoms out in,
 *
 *     int fusedMM_csr(imsg, m, n, k, alpha, nnz, rows, cols, val, indx, pntrb,
 *                     pntre, x, ldx, y, ldy, beta, z, ldz, z_arg)
 *         declared  /root/reference/csrc/fusedMM.h:77-99, csrc/fusedmm.cpp:63-85
 *         called    /root/reference/csrc/fusedmm.cpp:198  (from fusedmm_spmm_fw)
 *
 * whose body lives in an un-vendored CPU library (configure:2-7).  Differences
 * from that interface, all deliberate:
 *   - DEVICE pointers, int32 CSR indices, an explicit cudaStream_t; every call
 *     except the plan builder is asynchronous on that stream;
 *   - no allocation inside: derived metadata ("plan") and scratch ("workspace")
 *     live in caller-owned device buffers sized by the *_bytes queries;
 *   - the 5-stage VOP/ROP/SOP/VSC/AOP message is narrowed to the four reductions
 *     the wrapper actually sends (csrc/fusedmm.cpp:168-186), using the wrapper's
 *     own reduction codes (csrc/fusedmm.cpp:147-152);
 *   - val == NULL means "all ones" (the reference materialises a ones vector,
 *     isplib/__init__.py:51-57);
 *   - the status int is meant to be checked (the reference ignores it,
 *     csrc/fusedmm.cpp:198).
 * A literal host-pointer/int64 `fusedMM_csr` replacement with the reference's
 * exact signature is exported as isplib_b200_fusedmm_csr_host for maintainers
 * who want to relink csrc/fusedmm.cpp unchanged.
 *
 * No torch types appear here.  Only <stdint.h>/<stddef.h> are required; the
 * stream is passed as an opaque pointer (a cudaStream_t).
 */
#ifndef ISPLIB_B200_H
#define ISPLIB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISPLIB_B200_ABI_VERSION 2

/* reduction codes == `reduction` of fusedmm_spmm_fw, csrc/fusedmm.cpp:147-186 */
#define ISPLIB_REDUCE_SUM  0
#define ISPLIB_REDUCE_MAX  1
#define ISPLIB_REDUCE_MIN  2
#define ISPLIB_REDUCE_MEAN 3

/* status codes: 0/1/-1/128 keep the meaning of csrc/fusedMM.h:105-114 */
#define ISPLIB_SUCCESS          0
#define ISPLIB_FAIL             1     /* FUSEDMM_FAIL_RETURN */
#define ISPLIB_NOT_ENOUGH_MEM  (-1)   /* FUSEDMM_NOT_ENOUGH_MEM: plan/workspace too small */
#define ISPLIB_NO_OPT_IMPL      128   /* FUSEDMM_NO_OPT_IMPL: unsupported message/variant */
#define ISPLIB_INVALID_ARG      256   /* null/negative/misaligned/overflowing argument */
#define ISPLIB_CUDA_ERROR_BASE  1000  /* 1000 + cudaError_t */

/* flags for isplib_b200_spmm_csr_ex */
#define ISPLIB_FLAG_ACCUMULATE  0x1   /* merge into the existing out/arg_out instead of overwriting */
#define ISPLIB_FLAG_EMPTY_ZERO  0x2   /* max/min: rows with no entry produce 0 (torch_sparse
                                         convention) instead of keeping lowest()/max()
                                         (what csrc/fusedmm.cpp:147-150 leaves behind) */

#define ISPLIB_FLAG_RELU        0x4   /* out = max(out, 0) after everything else (isplib_b200_spmm_csr_fused) */

#define ISPLIB_VARIANT_AUTO (-1)

typedef void* isplib_stream_t; /* cudaStream_t */

/* Host-side description of a built plan (POD; filled by isplib_b200_plan_build). */
typedef struct isplib_b200_plan_info {
    int64_t m;                /* rows */
    int64_t nnz;              /* stored entries */
    int32_t seg_len;          /* max entries one warp work-item covers */
    int32_t reserved;
    int64_t num_items;        /* work items = sum_i max(1, ceil(deg_i / seg_len)) */
    int64_t num_split_rows;   /* rows covered by more than one item */
    int64_t num_split_items;  /* items that belong to split rows (= partial slots) */
    int64_t max_degree;
    int64_t num_empty_rows;
    uint64_t plan_bytes;      /* bytes of the device plan buffer actually used */
} isplib_b200_plan_info;

/* ---- library info ---------------------------------------------------------- */
int         isplib_b200_abi_version(void);
const char* isplib_b200_status_string(int status);

/* ---- plan: degree-aware work decomposition (replaces nothing in the reference;
 *      it is the metadata half of the `findbestk.py` replacement) -- */
/* Upper bound on the device plan buffer for (m, nnz, seg_len). seg_len <= 0 = default. */
int isplib_b200_plan_bytes(int64_t m, int64_t nnz, int32_t seg_len, size_t* bytes);
/* Builds the plan into plan_dev.  Synchronises `stream` once (reads back counts). */
int isplib_b200_plan_build(int64_t m, int64_t nnz, const int32_t* rowptr, int32_t seg_len,
                           void* plan_dev, size_t plan_dev_bytes,
                           isplib_b200_plan_info* info, isplib_stream_t stream);
/* Scratch needed by isplib_b200_spmm_csr* for this plan at feature width k. */
int isplib_b200_spmm_workspace_bytes(const isplib_b200_plan_info* info, int64_t k,
                                     int reduce, size_t* bytes);

/* ---- forward: replaces fusedMM_csr as called at csrc/fusedmm.cpp:198 --------
 * out[i,:] = REDUCE_{e in row i} val[e] * x[col[e],:]        (mean: / max(deg_i,1))
 * max/min also write arg_out[i,kk] = winning edge id e, or nnz if row i is empty
 * (csrc/fusedmm.cpp:171).  out (and arg_out) are fully written: no pre-init is
 * needed (the reference pre-fills them, csrc/fusedmm.cpp:147-152,171).
 * x: [n,k] fp32 row-major with row stride ldx; out: [m,k] with row stride ldo;
 * arg_out: [m,k] int64 with row stride ldo, required iff reduce is max/min.      */
int isplib_b200_spmm_csr(int reduce, int64_t m, int64_t n, int64_t k, int64_t nnz,
                         const int32_t* rowptr, const int32_t* col, const float* val,
                         const float* x, int64_t ldx, float* out, int64_t ldo,
                         int64_t* arg_out,
                         const isplib_b200_plan_info* info, const void* plan_dev,
                         void* workspace, size_t workspace_bytes,
                         int variant, isplib_stream_t stream);

/* Extended form used by the row-partitioned multi-GPU path:
 *   flags         ISPLIB_FLAG_*
 *   row_divisor   [m] or NULL: out[i,:] is divided by row_divisor[i] at the end
 *                 (sum with an explicit divisor == mean over a partitioned row)
 *   edge_ids      [nnz] or NULL: arg_out receives edge_ids[e] instead of e (global
 *                 edge ids of a column block); must be increasing within a row
 *   arg_sentinel  value written to arg_out where no entry won (nnz when edge_ids==NULL) */
int isplib_b200_spmm_csr_ex(int reduce, int64_t m, int64_t n, int64_t k, int64_t nnz,
                            const int32_t* rowptr, const int32_t* col, const float* val,
                            const float* x, int64_t ldx, float* out, int64_t ldo,
                            int64_t* arg_out,
                            const isplib_b200_plan_info* info, const void* plan_dev,
                            void* workspace, size_t workspace_bytes,
                            int variant, int flags, const float* row_divisor,
                            const int32_t* edge_ids, int64_t arg_sentinel,
                            isplib_stream_t stream);

/* Fused form: the forward plus what the reference's callers do to its result in separate
 * [M,K] passes, applied when a row is finalised:
 *     out[i,:] = relu?( REDUCE(...)[i,:] + addend_scale * addend[i,:] + bias[:] )
 *   GCN   `+ bias`, ReLU            /root/reference/tests/cpu/gcn-sparse.py:61-68 (GCNConv bias, F.relu)
 *   GIN   (1 + eps) * x_i + sum_j   /root/reference/tests/cpu/gin-sparse.py:73-78 (addend = x, scale = 1 + eps)
 *   SAGE  root term / residual      /root/reference/tests/cpu/graphSAGE-sparse.py:71-78
 * and, for max/min, two auxiliary outputs the backward scatter then streams instead of
 * gathering col[arg] / val[arg] (csrc/fusedmm.cpp:432-441): arg_col[i,kk] = col[arg] (-1 where
 * no entry won), arg_val[i,kk] = val[arg].  Every member is optional (NULL / 0); epi == NULL is
 * isplib_b200_spmm_csr_ex.  ReLU is requested with ISPLIB_FLAG_RELU in `flags`.  With
 * ISPLIB_FLAG_ACCUMULATE pass the epilogue on the LAST block only; arg_col/arg_val cannot be
 * combined with ISPLIB_FLAG_ACCUMULATE.                                                        */
typedef struct isplib_b200_epilogue {
    const float* bias;         /* [k] */
    const float* addend;       /* [m,k], row stride ld_addend */
    int64_t      ld_addend;
    float        addend_scale;
    int32_t      reserved;
    int32_t*     arg_col;      /* [m,k], row stride ldo (max/min only) */
    float*       arg_val;      /* [m,k], row stride ldo (max/min only, needs arg_col) */
} isplib_b200_epilogue;
int isplib_b200_spmm_csr_fused(int reduce, int64_t m, int64_t n, int64_t k, int64_t nnz,
                               const int32_t* rowptr, const int32_t* col, const float* val,
                               const float* x, int64_t ldx, float* out, int64_t ldo,
                               int64_t* arg_out,
                               const isplib_b200_plan_info* info, const void* plan_dev,
                               void* workspace, size_t workspace_bytes,
                               int variant, int flags, const float* row_divisor,
                               const int32_t* edge_ids, int64_t arg_sentinel,
                               const isplib_b200_epilogue* epi, isplib_stream_t stream);

/* ---- row-partitioned multi-GPU forward: gather of X fused INTO the SpMM -------------------
 * New functionality (the reference is single-process).  Rank r owns a
 * contiguous block of rows of A and the matching slice of X; its block's columns address the
 * owner-major gathered matrix [world * slice_rows, k].  Instead of an all-gather collective
 * followed by the SpMM, ONE kernel does both over NVLink peer memory: its first `copy_ctas` CTAs
 * PUSH the rank's own slice into every peer's buffer (posted stores, then a release-add on the peer's
 * arrival counter), the rest multiply, each work item starting as soon as the rows of ITS arrival
 * group have landed.  Flow control is by credits (a rank tells its peers when it has started a step),
 * x is double-buffered by step parity; no collective, no barrier kernel.
 *
 * isplib_b200_plan_build_grouped: like isplib_b200_plan_build, but segments never cross the
 * `n_runs` column runs [run_start[j], run_start[j+1]) and the work items are ordered by
 * run_group[j] (0 .. n_groups-1); group_item_end[g] receives the end of group g's items.
 * Usable with every forward entry point (the item order is free).                              */
int isplib_b200_plan_grouped_bytes(int64_t m, int64_t nnz, int32_t seg_len, int32_t n_runs, size_t* bytes);
int isplib_b200_plan_build_grouped(int64_t m, int64_t nnz, const int32_t* rowptr, const int32_t* col,
                                   int32_t seg_len, int32_t n_runs, const int32_t* run_start,
                                   const int32_t* run_group, int32_t n_groups,
                                   void* plan_dev, size_t plan_dev_bytes, isplib_b200_plan_info* info,
                                   int64_t* group_item_end, isplib_stream_t stream);

typedef struct isplib_b200_gather_desc {
    int32_t world, rank;
    int32_t n_groups;               /* owner mode: arrival groups incl. group 0 (the rank's own slice) */
    int32_t copy_ctas;              /* CTAs that push over NVLink; 0 = default (64); the same on every rank */
    void* const* peer_x;            /* [world] host array: rank q's gathered-x buffer OF THIS STEP'S PARITY as mapped
                                       into this process (symmetric memory); slice q = rows [q*slice_rows, ...) of
                                       EVERY buffer; peer_x[rank] == x */
    void* const* peer_arrive;       /* [world] host array: rank q's arrival counters of this parity, uint32[8] */
    void* const* peer_credit;       /* [world] host array: rank q's credit words, uint32[world] */
    const int32_t* owner_group;     /* owner mode, [world]: arrival group of each owner's slice HERE; [rank] == 0 */
    const int32_t* my_group_at_peer;/* owner mode, [world]: the group THIS rank's slice belongs to at peer q */
    int64_t slice_rows;             /* n == world * slice_rows */
    uint32_t* status;               /* device uint32, zeroed once: 1 after a wait timed out (4 s) */
    uint32_t epoch;                 /* 1, 2, 3, ... one per step on this buffer set; parity = epoch & 1 selects
                                       which of the two x buffers / arrival-counter sets the step uses */
    uint32_t tile_mode;             /* 0: arrival groups = column owners (grouped plan, rows split per group).
                                       1: arrival groups = the K TILES of the launch: the slice is pushed one K tile
                                          at a time, the items of tile t wait for tile t of every peer; rows stay
                                          whole, any plan works; the owner-mode members are ignored */
    const int64_t* group_item_end;  /* owner mode: [n_groups] host array from isplib_b200_plan_build_grouped */
    uint32_t parity_launch;         /* steps that have used this parity's arrival counters, including this one */
    uint32_t phase;                 /* 0: push + multiply (the product path).  1: push only, 2: multiply only
                                       (still waits for arrivals): the two halves of a step when several ranks are
                                       emulated on ONE GPU, where a rank's kernel cannot wait for kernels that have
                                       not been launched yet */
} isplib_b200_gather_desc;
/* x: the LOCAL gathered buffer [n = world * slice_rows, k] (row stride ldx, ldx % 4 == 0); its own
 * slice must hold this step's rows before the call (stream order); the peers write the other slices.
 * Everything else as isplib_b200_spmm_csr_fused.  world == 1 degenerates to the plain kernel. */
int isplib_b200_spmm_csr_gather(int reduce, int64_t m, int64_t n, int64_t k, int64_t nnz,
                                const int32_t* rowptr, const int32_t* col, const float* val,
                                float* x, int64_t ldx, float* out, int64_t ldo,
                                int64_t* arg_out,
                                const isplib_b200_plan_info* info, const void* plan_dev,
                                void* workspace, size_t workspace_bytes,
                                int variant, int flags, const float* row_divisor,
                                const int32_t* edge_ids, int64_t arg_sentinel,
                                const isplib_b200_epilogue* epi,
                                const isplib_b200_gather_desc* gd, isplib_stream_t stream);

/* ---- kernel variants + on-device selection: replaces autotuner/findbestk.py:29-41
 *      (an offline K sweep) by timing the eligible row-split / K-tile / unroll
 *      variants on the device, on the caller's own graph and buffers ------------ */
int         isplib_b200_variant_count(void);
const char* isplib_b200_variant_name(int variant);
/* 1 if `variant` can run this problem, else 0 */
int         isplib_b200_variant_supported(int variant, int reduce, int64_t k, int64_t ldx,
                                          int64_t ldo, const void* x, const void* out);
/* Heuristic default from the shape alone (n = rows of x): narrow K tiles only when they
 * make an [n, tile] slab of x L2-resident while the whole x is not. */
int         isplib_b200_variant_default(int reduce, int64_t n, int64_t k, int64_t ldx, int64_t ldo,
                                        const void* x, const void* out, double avg_degree);
/* Times every supported variant (1 warm-up + iters timed launches each, CUDA events
 * on `stream`), writes ms per launch to times_ms[variant] (negative = unsupported),
 * returns the fastest in *best_variant.  Overwrites out/arg_out.  Synchronises. */
int isplib_b200_spmm_autotune(int reduce, int64_t m, int64_t n, int64_t k, int64_t nnz,
                              const int32_t* rowptr, const int32_t* col, const float* val,
                              const float* x, int64_t ldx, float* out, int64_t ldo,
                              int64_t* arg_out,
                              const isplib_b200_plan_info* info, const void* plan_dev,
                              void* workspace, size_t workspace_bytes,
                              int iters, int* best_variant, float* times_ms,
                              isplib_stream_t stream);

/* ---- backward ----------------------------------------------------------------
 * sum/mean backward is the forward over the CSC view (csrc/fusedmm.cpp:285,375):
 * these two build that view once per graph on the device, replacing the
 * torch_sparse csr2csc()/colptr() argsort and the two index_selects the plugin
 * caches (isplib/__init__.py:69-99).                                             */
int isplib_b200_csr_transpose_workspace_bytes(int64_t m, int64_t n, int64_t nnz, size_t* bytes);
/* colptr[n+1], row_t[nnz] = row[csr2csc], csr2csc[nnz]: stable by-column order. */
int isplib_b200_csr_transpose(int64_t m, int64_t n, int64_t nnz,
                              const int32_t* rowptr, const int32_t* col,
                              int32_t* colptr, int32_t* row_t, int32_t* csr2csc,
                              void* workspace, size_t workspace_bytes,
                              isplib_stream_t stream);
/* val_t[p] = val[csr2csc[p]]                       (mean_weights == 0; isplib/__init__.py:79)
 * val_t[p] = val[csr2csc[p]] / max(deg(row_t[p]),1) (mean_weights != 0; isplib/__init__.py:86-93)
 * val == NULL stands for all ones.  rowptr is the ORIGINAL (untransposed) rowptr. */
int isplib_b200_permute_values(int64_t nnz, const float* val, const int32_t* csr2csc,
                               const int32_t* row_t, const int32_t* rowptr,
                               int mean_weights, float* val_t, isplib_stream_t stream);

/* max/min backward, fused: replaces the 5-8 ATen ops at csrc/fusedmm.cpp:417-446
 * (and :484-513).  For every (i,kk) with e = arg[i,kk] != arg_sentinel:
 *     grad_x[col[e], kk]  += (val ? val[e] : 1) * grad_out[i,kk]     (if grad_x)
 *     grad_val[e]         += x[col[e], kk] * grad_out[i,kk]          (if grad_val)
 * zero_init != 0 clears grad_x ([n,k], stride ldgx) / grad_val ([nnz]) first.      */
int isplib_b200_spmm_arg_backward(int64_t m, int64_t n, int64_t k, int64_t nnz,
                                  const int32_t* col, const float* val,
                                  const float* x, int64_t ldx,
                                  const int64_t* arg, int64_t ld_arg, int64_t arg_sentinel,
                                  const float* grad_out, int64_t ldgo,
                                  float* grad_x, int64_t ldgx, float* grad_val,
                                  int zero_init, isplib_stream_t stream);
/* The same grad_x from the forward's auxiliary outputs (isplib_b200_epilogue.arg_col / arg_val,
 * row stride ld_aux): two coalesced 4-byte streams replace the int64 arg read and the
 * col[arg] / val[arg] sector gathers; only the adds into grad_x stay random.
 *     grad_x[arg_col[i,kk], kk] += (arg_val ? arg_val[i,kk] : 1) * grad_out[i,kk]   where arg_col >= 0 */
int isplib_b200_spmm_arg_backward_aux(int64_t m, int64_t n, int64_t k,
                                      const int32_t* arg_col, const float* arg_val, int64_t ld_aux,
                                      const float* grad_out, int64_t ldgo,
                                      float* grad_x, int64_t ldgx,
                                      int zero_init, isplib_stream_t stream);

/* The same again for a grad_x far beyond L2 (Amazon-shape K=200: 1.25 GB), where every RED of the call
 * above is a random DRAM read-modify-write: the (target, value) pairs are first partitioned by
 * target-row range into 8-byte records (workspace: 8 * m * k bytes + 8 KB), then applied range after
 * range into an L2-resident slab of grad_x.  Same result up to the order of the float adds.        */
int isplib_b200_spmm_arg_backward_binned_workspace_bytes(int64_t m, int64_t n, int64_t k, size_t* bytes);
int isplib_b200_spmm_arg_backward_binned(int64_t m, int64_t n, int64_t k,
                                         const int32_t* arg_col, const float* arg_val, int64_t ld_aux,
                                         const float* grad_out, int64_t ldgo,
                                         float* grad_x, int64_t ldgx, int zero_init,
                                         void* workspace, size_t workspace_bytes, isplib_stream_t stream);

/* Gradient w.r.t. the stored values of the sum / mean SpMM (SDDMM on the CSR pattern):
 *     out_val[e] = < a[row(e),:], x[col[e],:] >   (mean_scale != 0: / max(deg(row(e)),1))
 * with a = grad_out.  The reference never computes it (csrc/fusedmm.cpp:268-272,349-353
 * return an undefined gradient) though its header has the ROP_DOT stage (csrc/fusedMM.h:34);
 * provided as the next op on the same access pattern.  Uses the
 * forward plan of the same graph.  a: [m,k] stride lda, x: [n,k] stride ldx.             */
int isplib_b200_sddmm_csr(int64_t m, int64_t n, int64_t k, int64_t nnz,
                          const int32_t* rowptr, const int32_t* col,
                          const float* a, int64_t lda, const float* x, int64_t ldx,
                          int mean_scale, float* out_val,
                          const isplib_b200_plan_info* info, const void* plan_dev,
                          isplib_stream_t stream);

/* ---- graph ingest: COO (e.g. a PyG edge_index, or a Matrix Market file) -> CSR ----------
 * Replaces the host-side argsort of torch_sparse's SparseTensor constructor that every
 * loader of the reference goes through (tests/cpu/dataset_loader.py:10, T.ToSparseTensor()).
 * Stable sort by (row, col): duplicates keep their input order (the README fixture has one).
 * Outputs: rowptr[m+1], col_out[nnz], perm[nnz] (CSR position -> input position) and, if
 * val != NULL, val_out[nnz].  All int32 / fp32, device pointers.                            */
int isplib_b200_coo_to_csr_workspace_bytes(int64_t m, int64_t n, int64_t nnz, size_t* bytes);
int isplib_b200_coo_to_csr(int64_t m, int64_t n, int64_t nnz,
                           const int32_t* row, const int32_t* col, const float* val,
                           int32_t* rowptr, int32_t* col_out, float* val_out, int32_t* perm,
                           void* workspace, size_t workspace_bytes, isplib_stream_t stream);

/* ---- index helpers ------------------------------------------------------------- */
/* int64 -> int32 narrowing of rowptr/col as the ops receive them
 * (csrc/fusedmm.cpp:128-129 reads int64).  *overflow_flag_dev (device int, may be
 * NULL) is set to 1 if any value does not fit. */
int isplib_b200_narrow_i64_to_i32(int64_t count, const int64_t* src, int32_t* dst,
                                  int32_t* overflow_flag_dev, isplib_stream_t stream);

/* ---- literal replacement of the reference entry point ---------------------------
 * Same signature and HOST-pointer/int64 contract as fusedMM_csr
 * (csrc/fusedMM.h:77-99): z/z_arg pre-initialised by the caller, accumulated into.
 * Uploads, runs the kernels above on the current device, downloads; synchronous.  */
int isplib_b200_fusedmm_csr_host(int32_t imessage, int64_t m, int64_t n, int64_t k,
                                 float alpha, int64_t nnz, int64_t rows, int64_t cols,
                                 const float* val, const int64_t* indx,
                                 const int64_t* pntrb, const int64_t* pntre,
                                 const float* x, int64_t ldx, const float* y, int64_t ldy,
                                 float beta, float* z, int64_t ldz, int64_t* z_arg);

#ifdef __cplusplus
}
#endif
#endif /* ISPLIB_B200_H */
