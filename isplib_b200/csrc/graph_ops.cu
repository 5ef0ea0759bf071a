// graph_ops.cu -- one-time, per-graph device work: the segment plan, the CSC view for
// the sum/mean backward, edge-value permutation and int64->int32 index narrowing.
//
// Reference behaviour being replaced:
//   * csr2csc()/colptr() of torch_sparse storage (an argsort over col*M+row) and the two
//     cached index_selects of the plugin            /root/reference/isplib/__init__.py:69-99
//   * nothing for the plan: the reference has no load balancing beyond OpenMP over rows.
// None of this is on the per-step hot path; CUB primitives (scan, radix sort) are used
// for the bulk sorting/scanning, the graph-specific kernels are hand-written.
#include "common.cuh"
#include <cub/cub.cuh>

namespace isplib {

// ---------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------
__global__ void plan_count_kernel(int m, int seg_len, const int32_t* __restrict__ rowptr,
                                  int32_t* __restrict__ seg_cnt, int32_t* __restrict__ part_cnt,
                                  unsigned long long* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int deg = 0, c = 0, pc = 0;
    if (i < m) {
        deg = rowptr[i + 1] - rowptr[i];
        c = deg <= seg_len ? 1 : (deg + seg_len - 1) / seg_len;
        pc = c > 1 ? c : 0;
        seg_cnt[i] = c;
        part_cnt[i] = pc;
    } else if (i == m) {
        seg_cnt[i] = 0;   // so the exclusive scan over m+1 items leaves the total at [m]
        part_cnt[i] = 0;
    }
    // block-level stats, then one atomic per block
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int max_deg = BR(tmp).Reduce(deg, cub::Max());
    __syncthreads();
    const int n_split = BR(tmp).Sum(pc > 0 ? 1 : 0);
    __syncthreads();
    const int n_empty = BR(tmp).Sum((i < m && deg == 0) ? 1 : 0);
    if (threadIdx.x == 0) {
        atomicMax(&counters[PC_MAX_DEG], (unsigned long long)max_deg);
        if (n_split) atomicAdd(&counters[PC_SPLIT_ROWS], (unsigned long long)n_split);
        if (n_empty) atomicAdd(&counters[PC_EMPTY], (unsigned long long)n_empty);
    }
}

// one warp per row: write the row id into each of its item slots, collect split rows
__global__ void plan_fill_kernel(int m, int seg_len, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ seg_off,
                                 const int32_t* __restrict__ part_off,
                                 int4* __restrict__ item_desc,
                                 unsigned long long* __restrict__ counters) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const int b = seg_off[row], e = seg_off[row + 1];
    const int rb = rowptr[row], re = rowptr[row + 1];
    const int pb = part_off[row];
    for (int w = b + lane; w < e; w += 32) {
        const int s = w - b;
        const int eb = rb + s * seg_len;
        item_desc[w] = make_int4(row, eb, min(re, eb + seg_len), (e - b > 1) ? pb + s : -1);
    }
    if (lane == 0) {
        if (row == m - 1) {
            counters[PC_ITEMS] = (unsigned long long)e;
            counters[PC_SPLIT_ITEMS] = (unsigned long long)part_off[m];
        }
    }
}

}  // namespace isplib

using namespace isplib;

extern "C" int isplib_b200_plan_bytes(int64_t m, int64_t nnz, int32_t seg_len, size_t* bytes) {
    if (!bytes || m < 0 || nnz < 0) return ISPLIB_INVALID_ARG;
    if (m >= INT32_MAX - 1 || nnz >= INT32_MAX - 64) return ISPLIB_INVALID_ARG;   // kernels index e0 + 63
    const PlanLayout L = plan_layout(m, nnz, seg_len);
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(m + 1));
    // temp: seg_cnt[m+1], part_cnt[m+1], cub scratch
    size_t o = L.off_temp;
    o = align_up(o + (size_t)(m + 1) * 4, 256);
    o = align_up(o + (size_t)(m + 1) * 4, 256);
    o = align_up(o + scan_tmp, 256);
    *bytes = o;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_plan_build(int64_t m, int64_t nnz, const int32_t* rowptr, int32_t seg_len,
                                      void* plan_dev, size_t plan_dev_bytes,
                                      isplib_b200_plan_info* info, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!info || m < 0 || nnz < 0 || (m > 0 && !rowptr)) return ISPLIB_INVALID_ARG;
    size_t need = 0;
    int st = isplib_b200_plan_bytes(m, nnz, seg_len, &need);
    if (st) return st;
    if (!plan_dev || plan_dev_bytes < need) return ISPLIB_NOT_ENOUGH_MEM;
    if ((reinterpret_cast<uintptr_t>(plan_dev) & 255u) != 0) return ISPLIB_INVALID_ARG;

    const int32_t S = effective_seg_len(seg_len);
    const PlanLayout L = plan_layout(m, nnz, seg_len);
    char* base = (char*)plan_dev;
    unsigned long long* counters = (unsigned long long*)(base + L.off_counters);
    int32_t* seg_off = (int32_t*)(base + L.off_seg_off);
    int32_t* part_off = (int32_t*)(base + L.off_part_off);
    int4* item_desc = (int4*)(base + L.off_item_desc);
    size_t o = L.off_temp;
    int32_t* seg_cnt = (int32_t*)(base + o);  o = align_up(o + (size_t)(m + 1) * 4, 256);
    int32_t* part_cnt = (int32_t*)(base + o); o = align_up(o + (size_t)(m + 1) * 4, 256);
    void* scan_tmp = base + o;
    size_t scan_bytes = plan_dev_bytes - o;

    ISPLIB_CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(int64_t), stream));

    const int threads = 256;
    const int blocks = (int)((m + 1 + threads - 1) / threads);
    plan_count_kernel<<<blocks, threads, 0, stream>>>((int)m, S, rowptr, seg_cnt, part_cnt, counters);
    ISPLIB_LAUNCH_CHECK();
    ISPLIB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, seg_cnt, seg_off, (int)(m + 1), stream));
    ISPLIB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, part_cnt, part_off, (int)(m + 1), stream));
    if (m > 0) {
        const int wpb = 8;
        plan_fill_kernel<<<(int)((m + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
            (int)m, S, rowptr, seg_off, part_off, item_desc, counters);
        ISPLIB_LAUNCH_CHECK();
    }
    unsigned long long h[8] = {0};
    ISPLIB_CUDA_TRY(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    ISPLIB_CUDA_TRY(cudaStreamSynchronize(stream));

    info->m = m;
    info->nnz = nnz;
    info->seg_len = S;
    info->reserved = 0;
    info->num_items = (int64_t)h[PC_ITEMS];
    info->num_split_rows = (int64_t)h[PC_SPLIT_ROWS];
    info->num_split_items = (int64_t)h[PC_SPLIT_ITEMS];
    info->max_degree = (int64_t)h[PC_MAX_DEG];
    info->num_empty_rows = (int64_t)h[PC_EMPTY];
    info->plan_bytes = (uint64_t)L.persistent_bytes;
    if (info->num_items > L.wmax) return ISPLIB_FAIL;  // rowptr not monotone / not matching nnz
    return ISPLIB_SUCCESS;
}

// ---------------------------------------------------------------------------------
// grouped plan (row-partitioned multi-GPU forward, fused with the gather of X)
// ---------------------------------------------------------------------------------
// The columns of a rank's row block are laid out owner-major ([P * Rc] gathered rows of X).  Owner
// ranges ("runs") are assigned to ARRIVAL GROUPS: group 0 = the rank's own slice, group g > 0 =
// the peers whose slices land g-th.  Segments never cross a run boundary, and the work items are
// ordered by group (all items of group 0 first, ...), so the kernel can start on a group as soon as
// its slices have landed while later groups are still in flight.  Partial slots of split rows stay in
// edge order, so the in-kernel merge (finish_item) is unchanged and max/min/arg stay bit-exact.
namespace isplib {

constexpr int kMaxRuns = 20;     // world <= 16: at most world + 1 contiguous owner runs
constexpr int kMaxGroups = 8;
struct RunTable { int n_runs; int n_groups; int start[kMaxRuns]; int group[kMaxRuns]; };

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// one thread per row: per-group item counts (group-major matrix cnt[g * m + row]), the row's total
__global__ void plan_grouped_count_kernel(int m, int seg_len, const int32_t* __restrict__ rowptr,
                                          const int32_t* __restrict__ col, const RunTable rt,
                                          int32_t* __restrict__ grp_cnt, int32_t* __restrict__ seg_cnt,
                                          int32_t* __restrict__ part_cnt, unsigned long long* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int deg = 0, pc = 0;
    if (i < m) {
        const int rb = rowptr[i], re = rowptr[i + 1];
        deg = re - rb;
        int cnt[kMaxGroups];
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) cnt[g] = 0;
        int total = 0, lo = rb;
        for (int j = 0; j < rt.n_runs; ++j) {
            const int hi = (j + 1 < rt.n_runs) ? lower_bound_i32(col, lo, re, rt.start[j + 1]) : re;
            const int c = (hi - lo + seg_len - 1) / seg_len;
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g) if (g == rt.group[j]) cnt[g] += c;
            total += c;
            lo = hi;
        }
        if (total == 0) { cnt[0] = 1; total = 1; }     // an empty row still owns one item: it writes the row
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) if (g < rt.n_groups) grp_cnt[(size_t)g * m + i] = cnt[g];
        seg_cnt[i] = total;
        pc = total > 1 ? total : 0;
        part_cnt[i] = pc;
    } else if (i == m) {
        seg_cnt[i] = 0;
        part_cnt[i] = 0;
        grp_cnt[(size_t)rt.n_groups * m] = 0;          // scan sentinel: total item count lands here
    }
    typedef cub::BlockReduce<int, 256> BR;
    __shared__ typename BR::TempStorage tmp;
    const int max_deg = BR(tmp).Reduce(deg, cub::Max());
    __syncthreads();
    const int n_split = BR(tmp).Sum(pc > 0 ? 1 : 0);
    __syncthreads();
    const int n_empty = BR(tmp).Sum((i < m && deg == 0) ? 1 : 0);
    if (threadIdx.x == 0) {
        atomicMax(&counters[PC_MAX_DEG], (unsigned long long)max_deg);
        if (n_split) atomicAdd(&counters[PC_SPLIT_ROWS], (unsigned long long)n_split);
        if (n_empty) atomicAdd(&counters[PC_EMPTY], (unsigned long long)n_empty);
    }
}

// one thread per row: emit the row's items, group by group in the item array, edge order in the slots
__global__ void plan_grouped_fill_kernel(int m, int seg_len, const int32_t* __restrict__ rowptr,
                                         const int32_t* __restrict__ col, const RunTable rt,
                                         const int32_t* __restrict__ grp_off, const int32_t* __restrict__ seg_off,
                                         const int32_t* __restrict__ part_off, int4* __restrict__ item_desc,
                                         unsigned long long* __restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int rb = rowptr[i], re = rowptr[i + 1];
    const int total = seg_off[i + 1] - seg_off[i];
    const int pb = part_off[i];
    int pos[kMaxGroups];
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) pos[g] = (g < rt.n_groups) ? grp_off[(size_t)g * m + i] : 0;
    if (re == rb) {
        item_desc[pos[0]] = make_int4(i, rb, rb, -1);
    } else {
        int s = 0, lo = rb;
        for (int j = 0; j < rt.n_runs; ++j) {
            const int hi = (j + 1 < rt.n_runs) ? lower_bound_i32(col, lo, re, rt.start[j + 1]) : re;
            const int g = rt.group[j];
            for (int eb = lo; eb < hi; eb += seg_len) {
                int w = 0;
#pragma unroll
                for (int q = 0; q < kMaxGroups; ++q) if (q == g) w = pos[q]++;
                item_desc[w] = make_int4(i, eb, min(hi, eb + seg_len), total > 1 ? pb + s : -1);
                ++s;
            }
            lo = hi;
        }
    }
    if (i == m - 1) {
        counters[PC_ITEMS] = (unsigned long long)grp_off[(size_t)rt.n_groups * m];
        counters[PC_SPLIT_ITEMS] = (unsigned long long)part_off[m];
    }
}

static inline int64_t grouped_wmax(int64_t m, int64_t nnz, int32_t S, int n_runs) {
    return m * (int64_t)n_runs + nnz / S + 1;
}

}  // namespace isplib

using namespace isplib;

extern "C" int isplib_b200_plan_grouped_bytes(int64_t m, int64_t nnz, int32_t seg_len, int32_t n_runs, size_t* bytes) {
    if (!bytes || m < 0 || nnz < 0 || n_runs < 1 || n_runs > kMaxRuns) return ISPLIB_INVALID_ARG;
    if (m >= INT32_MAX - 1 || nnz >= INT32_MAX - 64) return ISPLIB_INVALID_ARG;
    const int32_t S = effective_seg_len(seg_len);
    const PlanLayout L = plan_layout(m, nnz, seg_len);
    const int64_t wmax = grouped_wmax(m, nnz, S, n_runs);
    if (wmax >= INT32_MAX || (int64_t)kMaxGroups * m + 1 >= INT32_MAX) return ISPLIB_INVALID_ARG;
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_tmp, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(kMaxGroups * m + 1));
    size_t o = align_up(L.off_item_desc + (size_t)wmax * 16, 256);          // item_desc sized for the grouped bound
    o = align_up(o + (size_t)(m + 1) * 4, 256);                             // seg_cnt
    o = align_up(o + (size_t)(m + 1) * 4, 256);                             // part_cnt
    o = align_up(o + (size_t)(kMaxGroups * m + 1) * 4, 256);                // grp_cnt
    o = align_up(o + (size_t)(kMaxGroups * m + 1) * 4, 256);                // grp_off
    o = align_up(o + scan_tmp, 256);
    *bytes = o;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_plan_build_grouped(int64_t m, int64_t nnz, const int32_t* rowptr, const int32_t* col,
                                              int32_t seg_len, int32_t n_runs, const int32_t* run_start,
                                              const int32_t* run_group, int32_t n_groups,
                                              void* plan_dev, size_t plan_dev_bytes, isplib_b200_plan_info* info,
                                              int64_t* group_item_end, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!info || !run_start || !run_group || !group_item_end || m < 0 || nnz < 0 || (m > 0 && !rowptr) || (nnz > 0 && !col))
        return ISPLIB_INVALID_ARG;
    if (n_groups < 1 || n_groups > kMaxGroups) return ISPLIB_INVALID_ARG;
    size_t need = 0;
    int st = isplib_b200_plan_grouped_bytes(m, nnz, seg_len, n_runs, &need);
    if (st) return st;
    if (!plan_dev || plan_dev_bytes < need) return ISPLIB_NOT_ENOUGH_MEM;
    if ((reinterpret_cast<uintptr_t>(plan_dev) & 255u) != 0) return ISPLIB_INVALID_ARG;
    RunTable rt;
    rt.n_runs = n_runs;
    rt.n_groups = n_groups;
    for (int j = 0; j < n_runs; ++j) {
        if (run_group[j] < 0 || run_group[j] >= n_groups) return ISPLIB_INVALID_ARG;
        if (j > 0 && run_start[j] < run_start[j - 1]) return ISPLIB_INVALID_ARG;
        rt.start[j] = run_start[j];
        rt.group[j] = run_group[j];
    }
    const int32_t S = effective_seg_len(seg_len);
    const PlanLayout L = plan_layout(m, nnz, seg_len);
    const int64_t wmax = grouped_wmax(m, nnz, S, n_runs);
    char* base = (char*)plan_dev;
    unsigned long long* counters = (unsigned long long*)(base + L.off_counters);
    int32_t* seg_off = (int32_t*)(base + L.off_seg_off);
    int32_t* part_off = (int32_t*)(base + L.off_part_off);
    int4* item_desc = (int4*)(base + L.off_item_desc);
    size_t o = align_up(L.off_item_desc + (size_t)wmax * 16, 256);
    const size_t persistent = o;
    int32_t* seg_cnt = (int32_t*)(base + o);  o = align_up(o + (size_t)(m + 1) * 4, 256);
    int32_t* part_cnt = (int32_t*)(base + o); o = align_up(o + (size_t)(m + 1) * 4, 256);
    int32_t* grp_cnt = (int32_t*)(base + o);  o = align_up(o + (size_t)(kMaxGroups * m + 1) * 4, 256);
    int32_t* grp_off = (int32_t*)(base + o);  o = align_up(o + (size_t)(kMaxGroups * m + 1) * 4, 256);
    void* scan_tmp = base + o;
    size_t scan_bytes = plan_dev_bytes - o;

    ISPLIB_CUDA_TRY(cudaMemsetAsync(counters, 0, 8 * sizeof(int64_t), stream));
    const int threads = 256;
    plan_grouped_count_kernel<<<(int)((m + 1 + threads - 1) / threads), threads, 0, stream>>>(
        (int)m, S, rowptr, col, rt, grp_cnt, seg_cnt, part_cnt, counters);
    ISPLIB_LAUNCH_CHECK();
    const int n_grp = (int)((int64_t)n_groups * m + 1);
    ISPLIB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, grp_cnt, grp_off, n_grp, stream));
    ISPLIB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, seg_cnt, seg_off, (int)(m + 1), stream));
    ISPLIB_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, part_cnt, part_off, (int)(m + 1), stream));
    if (m > 0) {
        plan_grouped_fill_kernel<<<(int)((m + threads - 1) / threads), threads, 0, stream>>>(
            (int)m, S, rowptr, col, rt, grp_off, seg_off, part_off, item_desc, counters);
        ISPLIB_LAUNCH_CHECK();
    }
    unsigned long long h[8] = {0};
    int32_t ends[kMaxGroups] = {0};
    ISPLIB_CUDA_TRY(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, stream));
    for (int g = 0; g < n_groups; ++g)
        ISPLIB_CUDA_TRY(cudaMemcpyAsync(&ends[g], grp_off + (size_t)(g + 1) * m, 4, cudaMemcpyDeviceToHost, stream));
    ISPLIB_CUDA_TRY(cudaStreamSynchronize(stream));
    for (int g = 0; g < n_groups; ++g) group_item_end[g] = m > 0 ? (int64_t)ends[g] : 0;

    info->m = m;
    info->nnz = nnz;
    info->seg_len = S;
    info->reserved = 0;
    info->num_items = m > 0 ? (int64_t)h[PC_ITEMS] : 0;
    info->num_split_rows = (int64_t)h[PC_SPLIT_ROWS];
    info->num_split_items = (int64_t)h[PC_SPLIT_ITEMS];
    info->max_degree = (int64_t)h[PC_MAX_DEG];
    info->num_empty_rows = (int64_t)h[PC_EMPTY];
    info->plan_bytes = (uint64_t)persistent;
    if (info->num_items > wmax) return ISPLIB_FAIL;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_spmm_workspace_bytes(const isplib_b200_plan_info* info, int64_t k,
                                                int reduce, size_t* bytes) {
    if (reduce < 0 || reduce > 3) return ISPLIB_NO_OPT_IMPL;
    if (!info || !bytes || k < 0) return ISPLIB_INVALID_ARG;
    if (info->num_split_items == 0) { *bytes = 256; return ISPLIB_SUCCESS; }
    if ((int64_t)(info->num_split_items / 2 + 1) * ((k + kMinTileW - 1) / kMinTileW) >= INT32_MAX) return ISPLIB_INVALID_ARG;
    *bytes = workspace_layout(info->num_split_items, k, reduce == ISPLIB_REDUCE_MAX || reduce == ISPLIB_REDUCE_MIN).total;
    return ISPLIB_SUCCESS;
}

// ---------------------------------------------------------------------------------
// CSC view
// ---------------------------------------------------------------------------------
namespace isplib {

__global__ void iota_rows_kernel(int m, const int32_t* __restrict__ rowptr,
                                 int32_t* __restrict__ edge_id, int32_t* __restrict__ edge_row) {
    // one warp per row
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const int b = rowptr[row], e = rowptr[row + 1];
    for (int j = b + lane; j < e; j += 32) { edge_id[j] = j; edge_row[j] = row; }
}

// sorted column keys -> colptr (handles empty columns)
__global__ void colptr_from_sorted_kernel(int nnz, int n, const int32_t* __restrict__ keys,
                                          int32_t* __restrict__ colptr) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > nnz) return;
    const int prev = (p == 0) ? -1 : keys[p - 1];
    const int cur = (p == nnz) ? n : keys[p];
    for (int c = prev + 1; c <= cur; ++c) colptr[c] = p;
}

__global__ void gather_i32_kernel(int n, const int32_t* __restrict__ idx,
                                  const int32_t* __restrict__ src, int32_t* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void permute_values_kernel(int nnz, const float* __restrict__ val,
                                      const int32_t* __restrict__ csr2csc,
                                      const int32_t* __restrict__ row_t,
                                      const int32_t* __restrict__ rowptr, int mean_weights,
                                      float* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    float v = val ? val[csr2csc[p]] : 1.f;
    if (mean_weights) {
        const int r = row_t[p];
        const int deg = rowptr[r + 1] - rowptr[r];
        v = __fdiv_rn(v, (float)max(deg, 1));
    }
    out[p] = v;
}

__global__ void narrow_kernel(long long count, const long long* __restrict__ src,
                              int32_t* __restrict__ dst, int32_t* __restrict__ overflow) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long v = src[i];
    if (overflow && (v < 0 || v > (long long)INT32_MAX)) *overflow = 1;
    dst[i] = (int32_t)v;
}

static int key_bits(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace isplib

extern "C" int isplib_b200_csr_transpose_workspace_bytes(int64_t m, int64_t n, int64_t nnz, size_t* bytes) {
    if (!bytes || m < 0 || n < 0 || nnz < 0 || nnz >= INT32_MAX || n >= INT32_MAX - 1 || m >= INT32_MAX - 1)
        return ISPLIB_INVALID_ARG;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, key_bits(n));
    size_t o = 0;
    o = align_up(o + (size_t)nnz * 4, 256);  // sorted keys
    o = align_up(o + (size_t)nnz * 4, 256);  // edge ids
    o = align_up(o + (size_t)nnz * 4, 256);  // edge rows
    o = align_up(o + sort_tmp, 256);
    *bytes = o + 256;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_csr_transpose(int64_t m, int64_t n, int64_t nnz,
                                         const int32_t* rowptr, const int32_t* col,
                                         int32_t* colptr, int32_t* row_t, int32_t* csr2csc,
                                         void* workspace, size_t workspace_bytes,
                                         isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t need = 0;
    int st = isplib_b200_csr_transpose_workspace_bytes(m, n, nnz, &need);
    if (st) return st;
    if (!colptr || (nnz > 0 && (!rowptr || !col || !row_t || !csr2csc))) return ISPLIB_INVALID_ARG;
    if (!workspace || workspace_bytes < need) return ISPLIB_NOT_ENOUGH_MEM;
    char* base = (char*)workspace;
    base = (char*)align_up((size_t)(uintptr_t)base, 256);
    size_t o = 0;
    int32_t* keys_sorted = (int32_t*)(base + o); o = align_up(o + (size_t)nnz * 4, 256);
    int32_t* edge_id = (int32_t*)(base + o);     o = align_up(o + (size_t)nnz * 4, 256);
    int32_t* edge_row = (int32_t*)(base + o);    o = align_up(o + (size_t)nnz * 4, 256);
    void* sort_tmp = base + o;
    size_t sort_bytes = workspace_bytes - o - 256;

    if (nnz > 0) {
        const int wpb = 8;
        iota_rows_kernel<<<(int)((m + wpb - 1) / wpb), wpb * 32, 0, stream>>>((int)m, rowptr, edge_id, edge_row);
        ISPLIB_LAUNCH_CHECK();
        // stable LSD radix sort by column: rows stay ascending inside each column, which is
        // exactly argsort(col * M + row) for a CSR-ordered edge list
        ISPLIB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, col, keys_sorted, edge_id, csr2csc,
                                                        (int)nnz, 0, key_bits(n), stream));
        gather_i32_kernel<<<(int)((nnz + 255) / 256), 256, 0, stream>>>((int)nnz, csr2csc, edge_row, row_t);
        ISPLIB_LAUNCH_CHECK();
    }
    colptr_from_sorted_kernel<<<(int)((nnz + 1 + 255) / 256), 256, 0, stream>>>((int)nnz, (int)n, keys_sorted, colptr);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_permute_values(int64_t nnz, const float* val, const int32_t* csr2csc,
                                          const int32_t* row_t, const int32_t* rowptr,
                                          int mean_weights, float* val_t, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nnz < 0 || nnz >= INT32_MAX) return ISPLIB_INVALID_ARG;
    if (nnz == 0) return ISPLIB_SUCCESS;
    if (!csr2csc || !val_t || (mean_weights && (!row_t || !rowptr))) return ISPLIB_INVALID_ARG;
    permute_values_kernel<<<(int)((nnz + 255) / 256), 256, 0, stream>>>((int)nnz, val, csr2csc, row_t, rowptr,
                                                                         mean_weights, val_t);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_narrow_i64_to_i32(int64_t count, const int64_t* src, int32_t* dst,
                                             int32_t* overflow_flag_dev, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (count < 0) return ISPLIB_INVALID_ARG;
    if (count == 0) return ISPLIB_SUCCESS;
    if (!src || !dst) return ISPLIB_INVALID_ARG;
    const long long blocks = (count + 255) / 256;
    if (blocks > INT32_MAX) return ISPLIB_INVALID_ARG;
    narrow_kernel<<<(unsigned)blocks, 256, 0, stream>>>((long long)count, (const long long*)src, dst, overflow_flag_dev);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}


// ---------------------------------------------------------------------------------
// COO -> CSR (graph ingest)
// ---------------------------------------------------------------------------------
namespace isplib {

__global__ void coo_keys_kernel(int nnz, long long n, const int32_t* __restrict__ row,
                                const int32_t* __restrict__ col, unsigned long long* __restrict__ keys,
                                int32_t* __restrict__ idx, int32_t* __restrict__ bad, int m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int r = row[i], c = col[i];
    if (r < 0 || r >= m || c < 0 || c >= n) *bad = 1;
    keys[i] = (unsigned long long)r * (unsigned long long)n + (unsigned long long)c;
    idx[i] = i;
}

__global__ void coo_scatter_kernel(int nnz, long long n, const unsigned long long* __restrict__ keys,
                                   const int32_t* __restrict__ perm, const float* __restrict__ val,
                                   int32_t* __restrict__ row_sorted, int32_t* __restrict__ col_out,
                                   float* __restrict__ val_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const unsigned long long k = keys[i];
    const unsigned long long r = k / (unsigned long long)n;
    row_sorted[i] = (int32_t)r;
    col_out[i] = (int32_t)(k - r * (unsigned long long)n);
    if (val) val_out[i] = val[perm[i]];
}

}  // namespace isplib

static int key_bits64(int64_t m, int64_t n) {
    const unsigned long long mx = (unsigned long long)(m > 0 ? m : 1) * (unsigned long long)(n > 0 ? n : 1);
    int b = 1;
    while (b < 64 && (1ull << b) < mx) ++b;
    return b;
}

extern "C" int isplib_b200_coo_to_csr_workspace_bytes(int64_t m, int64_t n, int64_t nnz, size_t* bytes) {
    if (!bytes || m < 0 || n < 0 || nnz < 0 || nnz >= INT32_MAX || m >= INT32_MAX - 1 || n >= INT32_MAX) return ISPLIB_INVALID_ARG;
    size_t sort_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const int32_t*)nullptr, (int32_t*)nullptr, (int)nnz, 0, key_bits64(m, n));
    size_t o = 0;
    o = align_up(o + (size_t)nnz * 8, 256);   // keys in
    o = align_up(o + (size_t)nnz * 8, 256);   // keys sorted
    o = align_up(o + (size_t)nnz * 4, 256);   // iota
    o = align_up(o + (size_t)nnz * 4, 256);   // sorted rows
    o = align_up(o + 256, 256);               // error flag
    o = align_up(o + sort_tmp, 256);
    *bytes = o + 256;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_coo_to_csr(int64_t m, int64_t n, int64_t nnz,
                                      const int32_t* row, const int32_t* col, const float* val,
                                      int32_t* rowptr, int32_t* col_out, float* val_out, int32_t* perm,
                                      void* workspace, size_t workspace_bytes, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    size_t need = 0;
    int st = isplib_b200_coo_to_csr_workspace_bytes(m, n, nnz, &need);
    if (st) return st;
    if (!rowptr || (nnz > 0 && (!row || !col || !col_out || !perm)) || (val && !val_out)) return ISPLIB_INVALID_ARG;
    if (!workspace || workspace_bytes < need) return ISPLIB_NOT_ENOUGH_MEM;
    char* base = (char*)align_up((size_t)(uintptr_t)workspace, 256);
    size_t o = 0;
    unsigned long long* keys = (unsigned long long*)(base + o);        o = align_up(o + (size_t)nnz * 8, 256);
    unsigned long long* keys_sorted = (unsigned long long*)(base + o); o = align_up(o + (size_t)nnz * 8, 256);
    int32_t* iota = (int32_t*)(base + o);                              o = align_up(o + (size_t)nnz * 4, 256);
    int32_t* row_sorted = (int32_t*)(base + o);                        o = align_up(o + (size_t)nnz * 4, 256);
    int32_t* bad = (int32_t*)(base + o);                               o = align_up(o + 256, 256);
    void* sort_tmp = base + o;
    size_t sort_bytes = workspace_bytes - o - 256;
    ISPLIB_CUDA_TRY(cudaMemsetAsync(bad, 0, 4, stream));
    if (nnz > 0) {
        const int blocks = (int)((nnz + 255) / 256);
        coo_keys_kernel<<<blocks, 256, 0, stream>>>((int)nnz, (long long)n, row, col, keys, iota, bad, (int)m);
        ISPLIB_LAUNCH_CHECK();
        ISPLIB_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys, keys_sorted, iota, perm, (int)nnz, 0,
                                                        key_bits64(m, n), stream));
        coo_scatter_kernel<<<blocks, 256, 0, stream>>>((int)nnz, (long long)n, keys_sorted, perm, val, row_sorted, col_out, val_out);
        ISPLIB_LAUNCH_CHECK();
    }
    // rowptr from the sorted rows: same routine as colptr from sorted columns
    colptr_from_sorted_kernel<<<(int)((nnz + 1 + 255) / 256), 256, 0, stream>>>((int)nnz, (int)m, row_sorted, rowptr);
    ISPLIB_LAUNCH_CHECK();
    int32_t h_bad = 0;
    ISPLIB_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, stream));
    ISPLIB_CUDA_TRY(cudaStreamSynchronize(stream));
    return h_bad ? ISPLIB_INVALID_ARG : ISPLIB_SUCCESS;
}
