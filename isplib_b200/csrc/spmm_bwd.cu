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

// One thread per (row, feature).  arg / grad_out are read coalesced along the feature
// axis; the scatter target row differs per element (that is the operation), so the
// adds are fp32 RED atomics into L2.  HBM-bound: 8+4 bytes read, one 4-byte RMW, plus
// the dependent 4-byte col (and val) gathers per element.
__global__ void __launch_bounds__(256)
arg_backward_kernel(long long m, int k, const int32_t* __restrict__ col,
                    const float* __restrict__ val, const float* __restrict__ x, long long ldx,
                    const long long* __restrict__ arg, long long ld_arg, long long sentinel,
                    const float* __restrict__ grad_out, long long ldgo,
                    float* __restrict__ grad_x, long long ldgx, float* __restrict__ grad_val) {
    const long long total = m * (long long)k;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long i = t / k;
        const int kk = (int)(t - i * k);
        const long long e = __ldcs(arg + i * ld_arg + kk);
        if (e == sentinel) continue;
        const float g = __ldcs(grad_out + i * ldgo + kk);
        const int c = __ldg(col + e);
        if (grad_x) {
            const float v = val ? __fmul_rn(__ldg(val + e), g) : g;
            atomicAdd(grad_x + (long long)c * ldgx + kk, v);
        }
        if (grad_val) {
            atomicAdd(grad_val + e, __fmul_rn(__ldg(x + (long long)c * ldx + kk), g));
        }
    }
}

}  // namespace isplib

using namespace isplib;

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
    const long long total = (long long)m * k;
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)kNumSMs * 8 * 16;  // grid-stride beyond 16 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    arg_backward_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
        (long long)m, (int)k, col, val, x, (long long)ldx, (const long long*)arg, (long long)ld_arg,
        (long long)arg_sentinel, grad_out, (long long)ldgo, grad_x, (long long)ldgx, grad_val);
    ISPLIB_LAUNCH_CHECK();
    return ISPLIB_SUCCESS;
}
