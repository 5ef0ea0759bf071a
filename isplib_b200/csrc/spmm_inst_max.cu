// spmm_inst_max.cu -- instantiates the forward kernels of one reduction (OP_MAX).
#include "spmm_kernels.cuh"

namespace isplib {
SegKernel seg_kernel_max(const TileShape& t, int u, bool partial) { return pick_kernel<OP_MAX>(t, u, partial); }
SegKernel bulk_kernel_max(const TileShape& t, int stages) { return pick_bulk_kernel<OP_MAX>(t, stages); }
SegKernel lean256_kernel_max(int g, bool ragged, bool noval) { return pick_lean<OP_MAX, 8>(g, ragged, noval); }
SegKernel lean128_kernel_max(int g, bool ragged, bool noval) { return pick_lean<OP_MAX, 4>(g, ragged, noval); }
}  // namespace isplib
