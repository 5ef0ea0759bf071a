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
