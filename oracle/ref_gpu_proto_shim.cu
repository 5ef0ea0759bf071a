// oracle/ref_gpu_proto_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Host shim around the reference's own in-tree CUDA prototype kernel, included
// from where it lies (-I/root/reference/gpu/kernels): `fusedmm_kernel` and its
// launcher `fusedmm_spmm_trusted_kernel` at gpu/kernels/spmm.cuh:3-32.  That
// kernel is the only SpMM arithmetic the reference tree actually contains (sum
// only, int64 indices, accumulates into a pre-zeroed c).  Built into
// oracle/_ref/libref_gpu_proto.so by `make ref`; the GPU tests use it as one more
// pin for the sum path and bench.py can time it as "the reference's GPU kernel".
#include <cstdint>
#include <cuda_runtime.h>
#include "spmm.cuh"

extern "C" int ref_gpu_proto_spmm_sum(int m, int n, int k, int nnz,
                                      const int64_t* indx, const int64_t* ptrb,
                                      const float* val, const float* b, float* c)
{
    // device pointers; c must be zero-filled by the caller (the kernel does +=)
    fusedmm_spmm_trusted_kernel(m, n, k, nnz, indx, ptrb, val, b, c);
    return (int)cudaGetLastError();
}
