// spmm_inst_sum.cu -- instantiates the forward kernels of one reduction (OP_SUM).
#include "spmm_kernels.cuh"

namespace isplib {
SegKernel seg_kernel_sum(const TileShape& t, int u, bool partial) { return pick_kernel<OP_SUM>(t, u, partial); }
SegKernel bulk_kernel_sum(const TileShape& t, int stages) { return pick_bulk_kernel<OP_SUM>(t, stages); }
SegKernel lean256_kernel_sum(int g, bool ragged, bool noval) { return pick_lean<OP_SUM, 8>(g, ragged, noval); }
SegKernel lean128_kernel_sum(int g, bool ragged, bool noval) { return pick_lean<OP_SUM, 4>(g, ragged, noval); }
}  // namespace isplib
