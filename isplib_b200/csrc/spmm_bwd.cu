// spmm_bwd.cu -- fused max/min backward (arg-driven scatter) for sm_100a.
//
// Replaces the ATen op chain of the reference's FusedMM_SPMMMax/Min::backward:
//   invalid = arg == nnz; arg.masked_fill; value.index_select(arg).mul_(grad_out);
//   masked_fill_; col.index_select(arg); zeros_like(mat).scatter_add_(-2, ind, v)
//                                   /root/reference/csrc/fusedmm.cpp:417-446, :484-513
// (each step there materialises an [M,K] temporary) by a single pass over arg/grad_out.
// sum/mean backward needs no kernel of its own: it is the forward kernel run on the CSC
// view (csrc/fusedmm.cpp:285,375), see graph_ops.cu.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace isplib {

// arg / grad_out are read coalesced along the feature axis (4 features per thread as 2x16 B
// + 16 B when the rows allow it); the scatter target row differs per element (that is the
// operation), so the adds are fp32 RED atomics into L2.  All dependent gathers of a thread's
// 4 elements (col[e], val[e], x[col[e]]) are issued together before the first atomic.
//
// K tiling (blockIdx.y, slowest): the atomics are read-modify-writes of 32-byte sectors of
// grad_x at random rows; if grad_x does not fit L2 every one of them costs a DRAM sector read
// and a write-back.  Sweeping one [n, kt] column slab at a time keeps the slab L2-resident, and
// -- unlike in the forward -- narrow tiles are free here: arg / grad_out are still read exactly
// once, just in kt-wide row pieces (>= 32 bytes).
template <int V>
__global__ void __launch_bounds__(256)
arg_backward_kernel(long long m, int k, int kt, const int32_t* __restrict__ col,
                    const float* __restrict__ val, const float* __restrict__ x, long long ldx,
                    const long long* __restrict__ arg, long long ld_arg, long long sentinel,
                    const float* __restrict__ grad_out, long long ldgo,
                    float* __restrict__ grad_x, long long ldgx, float* __restrict__ grad_val) {
    const int k0 = blockIdx.y * kt;
    const int kw = min(kt, k - k0);                       // width of this tile
    const int kv = (kw + V - 1) / V;                      // V-wide groups per row in this tile
    const long long total = m * (long long)kv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / kv;
        const int kk = k0 + (int)(t - i * kv) * V;
        long long e[V];
        float g[V];
        if constexpr (V == 4) {
            const longlong2 a0 = __ldcs(reinterpret_cast<const longlong2*>(arg + i * ld_arg + kk));
            const longlong2 a1 = __ldcs(reinterpret_cast<const longlong2*>(arg + i * ld_arg + kk) + 1);
            const float4 gg = __ldcs(reinterpret_cast<const float4*>(grad_out + i * ldgo + kk));
            e[0] = a0.x; e[1] = a0.y; e[2] = a1.x; e[3] = a1.y;
            g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
        } else {
            e[0] = __ldcs(arg + i * ld_arg + kk);
            g[0] = __ldcs(grad_out + i * ldgo + kk);
        }
        int c[V];
        float a[V];
        bool ok[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            ok[v] = (e[v] != sentinel);
            c[v] = ok[v] ? __ldg(col + e[v]) : 0;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) a[v] = (val && ok[v]) ? __ldg(val + e[v]) : 1.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (!ok[v]) continue;
            if (grad_x) atomicAdd(grad_x + (long long)c[v] * ldgx + kk + v, val ? __fmul_rn(a[v], g[v]) : g[v]);
            if (grad_val) atomicAdd(grad_val + e[v], __fmul_rn(__ldg(x + (long long)c[v] * ldx + kk + v), g[v]));
        }
    }
}

// The same scatter fed by the forward's auxiliary outputs (isplib_b200_epilogue.arg_col /
// arg_val): the winner's column and value arrive as two coalesced 4-byte streams, so neither the
// int64 arg nor the 32-byte sectors around col[arg] / val[arg] are read; 12 bytes per output
// element instead of 12 + 2 sectors.  Only the RED into grad_x stays random.
template <int V>
__global__ void __launch_bounds__(256)
arg_backward_aux_kernel(long long m, int k, int kt, const int32_t* __restrict__ arg_col,
                        const float* __restrict__ arg_val, long long ld_aux,
                        const float* __restrict__ grad_out, long long ldgo,
                        float* __restrict__ grad_x, long long ldgx) {
    const int k0 = blockIdx.y * kt;
    const int kw = min(kt, k - k0);
    const int kv = (kw + V - 1) / V;
    const long long total = m * (long long)kv;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / kv;
        const int kk = k0 + (int)(t - i * kv) * V;
        int c[V];
        float a[V], g[V];
        if constexpr (V == 4) {
            const int4 cc = __ldcs(reinterpret_cast<const int4*>(arg_col + i * ld_aux + kk));
            const float4 gg = __ldcs(reinterpret_cast<const float4*>(grad_out + i * ldgo + kk));
            c[0] = cc.x; c[1] = cc.y; c[2] = cc.z; c[3] = cc.w;
            g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
            if (arg_val) {
                const float4 aa = __ldcs(reinterpret_cast<const float4*>(arg_val + i * ld_aux + kk));
                a[0] = aa.x; a[1] = aa.y; a[2] = aa.z; a[3] = aa.w;
            }
        } else {
            c[0] = __ldcs(arg_col + i * ld_aux + kk);
            g[0] = __ldcs(grad_out + i * ldgo + kk);
            if (arg_val) a[0] = __ldcs(arg_val + i * ld_aux + kk);
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (c[v] >= 0) atomicAdd(grad_x + (long long)c[v] * ldgx + kk + v, arg_val ? __fmul_rn(a[v], g[v]) : g[v]);
    }
}

// ------------------------------------------------------------------------------------
// binned scatter: for a grad_x far beyond L2 (Amazon-shape K=200: 1.25 GB) every RED above is a
// random 32-byte read-modify-write in DRAM -- 26 GB of traffic for 7.5 GB of algorithmic bytes
// (profiles/r2_full_argaux_amazon.txt).  Here the (target, value) pairs are first PARTITIONED by
// target-row range ("bin": a [rows_per_bin, k] slab of grad_x that fits L2), written as 8-byte
// records, and then applied bin after bin, so the REDs of a bin hit an L2-resident slab that goes
// to DRAM once.  Three passes over coalesced streams instead of one pass of random DRAM RMWs.
// ------------------------------------------------------------------------------------
constexpr int kMaxBins = 256;

// Both passes below are streams over dense [m*k] arrays: 16-byte loads, all of a thread's loads
// issued before the first shared-memory atomic (the first versions loaded 4 bytes at a time behind
// the atomics and ran at 1.4 TB/s / 1.1 TB/s: one load in flight per thread).
__global__ void __launch_bounds__(256)
arg_bin_count_kernel(long long total4, const int4* __restrict__ arg_col4, int rows_per_bin, int nbins,
                     unsigned long long* __restrict__ bin_count) {
    __shared__ unsigned int h[8][kMaxBins];
    for (int i = threadIdx.x; i < 8 * kMaxBins; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    unsigned int* const mine = h[threadIdx.x >> 5];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += 4 * stride) {
        int4 c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long j = t + u * stride;
            c[u] = j < total4 ? __ldcs(arg_col4 + j) : make_int4(-1, -1, -1, -1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (c[u].x >= 0) atomicAdd(&mine[c[u].x / rows_per_bin], 1u);
            if (c[u].y >= 0) atomicAdd(&mine[c[u].y / rows_per_bin], 1u);
            if (c[u].z >= 0) atomicAdd(&mine[c[u].z / rows_per_bin], 1u);
            if (c[u].w >= 0) atomicAdd(&mine[c[u].w / rows_per_bin], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        unsigned int s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += h[w][i];
        if (s) atomicAdd(bin_count + i, (unsigned long long)s);
    }
}

// bin_start = exclusive prefix of bin_count; cursors start at bin_start
__global__ void arg_bin_prefix_kernel(int nbins, const unsigned long long* __restrict__ bin_count,
                                      unsigned long long* __restrict__ bin_start, unsigned long long* __restrict__ cursor) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long s = 0;
        for (int b = 0; b < nbins; ++b) { bin_start[b] = s; cursor[b] = s; s += bin_count[b]; }
        bin_start[nbins] = s;
    }
}

// every CTA takes a tile of 4096 elements (4 x 16 bytes per thread and stream); a warp ranks its
// elements inside its own per-bin counters, the CTA reserves its share of every bin with ONE global
// atomic per bin, and the records are written there (the records of one warp and bin are contiguous:
// L2 merges the 8-byte writes into sectors)
constexpr int kBinTile = 256 * 16;
__global__ void __launch_bounds__(256)
arg_bin_partition_kernel(long long total4, int k, const int4* __restrict__ arg_col4, const float4* __restrict__ arg_val4,
                         const float4* __restrict__ grad_out4, int rows_per_bin, int nbins,
                         unsigned long long* __restrict__ cursor, uint2* __restrict__ records) {
    __shared__ unsigned int cnt[8][kMaxBins];          // per warp and bin: count, then offset inside the CTA's share
    __shared__ unsigned long long base[kMaxBins];      // the CTA's share of every bin
    for (int i = threadIdx.x; i < 8 * kMaxBins; i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5;
    const long long t0 = (long long)blockIdx.x * (kBinTile / 4);
    int c[16];
    float v[16];
    unsigned int rank[16];
    {
        int4 cc[4];
        float4 gg[4], aa[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long j = t0 + (long long)u * 256 + threadIdx.x;
            const bool in = j < total4;
            cc[u] = in ? __ldcs(arg_col4 + j) : make_int4(-1, -1, -1, -1);
            gg[u] = in ? __ldcs(grad_out4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            aa[u] = (in && arg_val4) ? __ldcs(arg_val4 + j) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            c[4 * u] = cc[u].x; c[4 * u + 1] = cc[u].y; c[4 * u + 2] = cc[u].z; c[4 * u + 3] = cc[u].w;
            v[4 * u] = arg_val4 ? __fmul_rn(aa[u].x, gg[u].x) : gg[u].x;
            v[4 * u + 1] = arg_val4 ? __fmul_rn(aa[u].y, gg[u].y) : gg[u].y;
            v[4 * u + 2] = arg_val4 ? __fmul_rn(aa[u].z, gg[u].z) : gg[u].z;
            v[4 * u + 3] = arg_val4 ? __fmul_rn(aa[u].w, gg[u].w) : gg[u].w;
        }
    }
#pragma unroll
    for (int u = 0; u < 16; ++u)
        if (c[u] >= 0) rank[u] = atomicAdd(&cnt[w][c[u] / rows_per_bin], 1u);
    int kk0[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) kk0[u] = (int)(((t0 + (long long)u * 256 + threadIdx.x) * 4) % k);
    __syncthreads();
    for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
        unsigned int s = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) { const unsigned int n = cnt[ww][i]; cnt[ww][i] = s; s += n; }
        base[i] = s ? atomicAdd(cursor + i, (unsigned long long)s) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 16; ++u) {
        if (c[u] >= 0) {
            // element index -> column kk of the dense [m, k] array (one 64-bit modulo per 4-element vector)
            int kk = kk0[u >> 2] + (u & 3);
            if (kk >= k) kk -= k;
            const int b = c[u] / rows_per_bin;
            const unsigned int off = (unsigned int)(c[u] - b * rows_per_bin) * (unsigned int)k + (unsigned int)kk;
            records[base[b] + cnt[w][b] + rank[u]] = make_uint2(off, __float_as_uint(v[u]));
        }
    }
}

// bin after bin (blockIdx.y, dispatched in order): coalesced records in, REDs into the bin's slab
__global__ void __launch_bounds__(256)
arg_bin_apply_kernel(const unsigned long long* __restrict__ bin_start, const uint2* __restrict__ records,
                     int rows_per_bin, int k, long long ldgx, float* __restrict__ grad_x) {
    const int b = blockIdx.y;
    const unsigned long long rb = bin_start[b], re = bin_start[b + 1];
    float* const slab = grad_x + (long long)b * rows_per_bin * ldgx;
    for (unsigned long long r = rb + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < re;
         r += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2 rec = __ldcs(records + r);
        const unsigned int row = rec.x / (unsigned int)k;
        const unsigned int kk = rec.x - row * (unsigned int)k;
        atomicAdd(slab + (long long)row * ldgx + kk, __uint_as_float(rec.y));
    }
}

}  // namespace isplib

using namespace isplib;

// K tile of the scatter target: the widest power-of-two slab (>= 32 floats) whose [n, kt] slice of
// grad_x stays L2-resident; untiled when grad_x fits anyway or when no such slab exists.  Measured
// (B200): Reddit-shape K=256 1.32 -> 0.98 ms with 32-wide tiles; Amazon-shape K=200 would
// need 8-wide tiles and gets SLOWER (8.5 -> 12.0 ms), hence the 32-float floor.
static int scatter_k_tile(int64_t n, int64_t k) {
    int kt = (int)k;
    if ((double)n * (double)k * 4.0 > 64.0 * 1024 * 1024) {
        kt = 256;
        while (kt > 32 && (double)n * (double)kt * 4.0 > 48.0 * 1024 * 1024) kt >>= 1;
        if (kt >= k || (double)n * (double)kt * 4.0 > 48.0 * 1024 * 1024) kt = (int)k;
    }
    return kt;
}

// binned variant: rows_per_bin so that a [rows_per_bin, k] slab of grad_x is ~32 MB
static int binned_layout(int64_t n, int64_t k, int* rows_per_bin, int* nbins) {
    double slab_bytes = 32.0 * 1024 * 1024;
    if (const char* e = getenv("ISPLIB_B200_BIN_BYTES")) { const double v = atof(e); if (v >= 4.0) slab_bytes = v; }   // tests: many bins on small graphs
    int64_t rpb = (int64_t)(slab_bytes / ((double)k * 4.0));
    if (rpb < 1) rpb = 1;
    int64_t nb = (n + rpb - 1) / rpb;
    if (nb > kMaxBins) { nb = kMaxBins; rpb = (n + nb - 1) / nb; nb = (n + rpb - 1) / rpb; }
    if (rpb * k >= (int64_t)UINT32_MAX) return ISPLIB_INVALID_ARG;
    *rows_per_bin = (int)rpb;
    *nbins = (int)nb;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_spmm_arg_backward_binned_workspace_bytes(int64_t m, int64_t n, int64_t k, size_t* bytes) {
    if (!bytes || m < 0 || n < 0 || k < 0) return ISPLIB_INVALID_ARG;
    *bytes = (size_t)m * (size_t)k * 8 + (size_t)(3 * kMaxBins + 8) * 8 + 512;
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_spmm_arg_backward_binned(int64_t m, int64_t n, int64_t k,
                                                    const int32_t* arg_col, const float* arg_val, int64_t ld_aux,
                                                    const float* grad_out, int64_t ldgo,
                                                    float* grad_x, int64_t ldgx, int zero_init,
                                                    void* workspace, size_t workspace_bytes, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 0 || n < 0 || k < 0 || k > INT32_MAX || !grad_x) return ISPLIB_INVALID_ARG;
    if (m > 0 && k > 0 && (!arg_col || !grad_out)) return ISPLIB_INVALID_ARG;
    if (ld_aux < k || ldgo < k || ldgx < k) return ISPLIB_INVALID_ARG;
    size_t need = 0;
    isplib_b200_spmm_arg_backward_binned_workspace_bytes(m, n, k, &need);
    if (!workspace || workspace_bytes < need) return ISPLIB_NOT_ENOUGH_MEM;
    if (zero_init && n > 0 && k > 0) {
        if (ldgx == k) ISPLIB_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)n * (size_t)k * 4, stream));
        else ISPLIB_CUDA_TRY(cudaMemset2DAsync(grad_x, (size_t)ldgx * 4, 0, (size_t)k * 4, (size_t)n, stream));
    }
    if (m == 0 || k == 0 || n == 0) return ISPLIB_SUCCESS;
    // the streams are read as dense 16-byte vectors; anything else takes the direct scatter
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool dense = ld_aux == k && ldgo == k && ((int64_t)m * k) % 4 == 0 && al16(arg_col) && al16(grad_out) &&
                       (!arg_val || al16(arg_val));
    if (!dense)
        return isplib_b200_spmm_arg_backward_aux(m, n, k, arg_col, arg_val, ld_aux, grad_out, ldgo, grad_x, ldgx, 0, stream_);
    int rows_per_bin = 0, nbins = 0;
    int st = binned_layout(n, k, &rows_per_bin, &nbins);
    if (st) return st;
    char* w = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    unsigned long long* bin_count = (unsigned long long*)w;
    unsigned long long* bin_start = bin_count + kMaxBins;
    unsigned long long* cursor = bin_start + kMaxBins + 8;
    uint2* records = (uint2*)(w + (size_t)(3 * kMaxBins + 8) * 8);
    ISPLIB_CUDA_TRY(cudaMemsetAsync(bin_count, 0, (size_t)kMaxBins * 8, stream));
    const long long total = (long long)m * k, total4 = total / 4;
    const int cblocks = (int)std::min<long long>((total4 + 1023) / 1024, (long long)kNumSMs * 8);
    arg_bin_count_kernel<<<cblocks, 256, 0, stream>>>(total4, (const int4*)arg_col, rows_per_bin, nbins, bin_count);
    ISPLIB_LAUNCH_CHECK();
    arg_bin_prefix_kernel<<<1, 32, 0, stream>>>(nbins, bin_count, bin_start, cursor);
    ISPLIB_LAUNCH_CHECK();
    const long long pblocks = (total + kBinTile - 1) / kBinTile;
    if (pblocks > INT32_MAX) return ISPLIB_INVALID_ARG;
    arg_bin_partition_kernel<<<(unsigned)pblocks, 256, 0, stream>>>(total4, (int)k, (const int4*)arg_col, (const float4*)arg_val,
                                                                    (const float4*)grad_out, rows_per_bin, nbins, cursor, records);
    ISPLIB_LAUNCH_CHECK();
    const dim3 agrid((unsigned)(kNumSMs * 16), (unsigned)nbins);
    arg_bin_apply_kernel<<<agrid, 256, 0, stream>>>(bin_start, records, rows_per_bin, (int)k, (long long)ldgx, grad_x);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_spmm_arg_backward_aux(int64_t m, int64_t n, int64_t k,
                                                 const int32_t* arg_col, const float* arg_val, int64_t ld_aux,
                                                 const float* grad_out, int64_t ldgo,
                                                 float* grad_x, int64_t ldgx,
                                                 int zero_init, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 0 || n < 0 || k < 0 || k > INT32_MAX) return ISPLIB_INVALID_ARG;
    if (!grad_x) return ISPLIB_INVALID_ARG;
    if (m > 0 && k > 0 && (!arg_col || !grad_out)) return ISPLIB_INVALID_ARG;
    if (ld_aux < k || ldgo < k || ldgx < k) return ISPLIB_INVALID_ARG;
    if (zero_init && n > 0 && k > 0) {
        if (ldgx == k) ISPLIB_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)n * (size_t)k * 4, stream));
        else ISPLIB_CUDA_TRY(cudaMemset2DAsync(grad_x, (size_t)ldgx * 4, 0, (size_t)k * 4, (size_t)n, stream));
    }
    if (m == 0 || k == 0 || n == 0) return ISPLIB_SUCCESS;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool v4 = (k % 4 == 0) && (ld_aux % 4 == 0) && (ldgo % 4 == 0) && al16(arg_col) && al16(grad_out) &&
                    (!arg_val || al16(arg_val));
    const int kt = scatter_k_tile(n, k);
    const int ntiles = (int)((k + kt - 1) / kt);
    if (ntiles > 65535) return ISPLIB_NO_OPT_IMPL;
    const long long total = (long long)m * (v4 ? (kt + 3) / 4 : kt);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 8 * 32;
    if (blocks > cap) blocks = cap;
    const dim3 grid((unsigned)blocks, (unsigned)ntiles);
    if (v4)
        arg_backward_aux_kernel<4><<<grid, 256, 0, stream>>>((long long)m, (int)k, kt, arg_col, arg_val, (long long)ld_aux,
                                                             grad_out, (long long)ldgo, grad_x, (long long)ldgx);
    else
        arg_backward_aux_kernel<1><<<grid, 256, 0, stream>>>((long long)m, (int)k, kt, arg_col, arg_val, (long long)ld_aux,
                                                             grad_out, (long long)ldgo, grad_x, (long long)ldgx);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}

extern "C" int isplib_b200_spmm_arg_backward(int64_t m, int64_t n, int64_t k, int64_t nnz,
                                             const int32_t* col, const float* val,
                                             const float* x, int64_t ldx,
                                             const int64_t* arg, int64_t ld_arg, int64_t arg_sentinel,
                                             const float* grad_out, int64_t ldgo,
                                             float* grad_x, int64_t ldgx, float* grad_val,
                                             int zero_init, isplib_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (m < 0 || n < 0 || k < 0 || nnz < 0 || k > INT32_MAX) return ISPLIB_INVALID_ARG;
    if (!grad_x && !grad_val) return ISPLIB_SUCCESS;
    if (grad_val && !x) return ISPLIB_INVALID_ARG;
    if (m > 0 && k > 0 && (!arg || !grad_out || (nnz > 0 && !col))) return ISPLIB_INVALID_ARG;
    if (ld_arg < k || ldgo < k || (grad_x && ldgx < k)) return ISPLIB_INVALID_ARG;
    if (zero_init) {
        if (grad_x && n > 0 && k > 0) {
            if (ldgx == k) {
                ISPLIB_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)n * (size_t)k * 4, stream));
            } else {
                ISPLIB_CUDA_TRY(cudaMemset2DAsync(grad_x, (size_t)ldgx * 4, 0, (size_t)k * 4, (size_t)n, stream));
            }
        }
        if (grad_val && nnz > 0) ISPLIB_CUDA_TRY(cudaMemsetAsync(grad_val, 0, (size_t)nnz * 4, stream));
    }
    if (m == 0 || k == 0 || nnz == 0) return ISPLIB_SUCCESS;
    const bool v4 = (k % 4 == 0) && (ld_arg % 2 == 0) && (ldgo % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(arg) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(grad_out) & 15u) == 0);
    const int kt = grad_x ? scatter_k_tile(n, k) : (int)k;
    const int ntiles = (int)((k + kt - 1) / kt);
    if (ntiles > 65535) return ISPLIB_NO_OPT_IMPL;
    const long long total = (long long)m * (v4 ? (kt + 3) / 4 : kt);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 8 * 32;  // grid-stride beyond 32 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    const dim3 grid((unsigned)blocks, (unsigned)ntiles);
    if (v4)
        arg_backward_kernel<4><<<grid, 256, 0, stream>>>(
            (long long)m, (int)k, kt, col, val, x, (long long)ldx, (const long long*)arg, (long long)ld_arg,
            (long long)arg_sentinel, grad_out, (long long)ldgo, grad_x, (long long)ldgx, grad_val);
    else
        arg_backward_kernel<1><<<grid, 256, 0, stream>>>(
            (long long)m, (int)k, kt, col, val, x, (long long)ldx, (const long long*)arg, (long long)ld_arg,
            (long long)arg_sentinel, grad_out, (long long)ldgo, grad_x, (long long)ldgx, grad_val);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}
