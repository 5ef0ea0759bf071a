/* fusedmm_compat.c -- exports the reference's own C symbols on top of isplib_b200.
 *
 *   int  fusedMM_csr(...)        declared /root/reference/csrc/fusedMM.h:77-99,
 *                                called   /root/reference/csrc/fusedmm.cpp:198
 *   void performDummySpMM(int64) declared /root/reference/csrc/fusedmm.cpp:61
 *
 * Built into isplib_b200/libfusedmm_b200_compat.so.  A maintainer who links the
 * reference's unmodified csrc/fusedmm.cpp against this library instead of
 * csrc/fusedmm/fusedmm_cpu.a (setup.py:124-128) gets the B200 kernels behind the
 * reference's CPU-tensor operators: host pointers in, host pointers out, synchronous --
 * exactly the contract of the library it replaces.  See INTEGRATION.md.
 */
#include <stdint.h>
#include "../../include/isplib_b200.h"

void performDummySpMM(int64_t flag) { (void)flag; }

int fusedMM_csr(const int32_t imessage, const int64_t m, const int64_t n, const int64_t k,
                const float alpha, const int64_t nnz, const int64_t rows, const int64_t cols,
                const float* val, const int64_t* indx, const int64_t* pntrb, const int64_t* pntre,
                const float* x, const int64_t ldx, const float* y, const int64_t ldy,
                const float beta, float* z, const int64_t ldz, int64_t* z_arg)
{
    return isplib_b200_fusedmm_csr_host(imessage, m, n, k, alpha, nnz, rows, cols, val, indx, pntrb,
                                        pntre, x, ldx, y, ldy, beta, z, ldz, z_arg);
}
