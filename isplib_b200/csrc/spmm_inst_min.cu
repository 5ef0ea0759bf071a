// spmm_inst_min.cu -- instantiates the forward kernels of one reduction (OP_MIN).
#include "spmm_kernels.cuh"

namespace isplib {
SegKernel seg_kernel_min(const TileShape& t, int u, bool partial) { return pick_kernel<OP_MIN>(t, u, partial); }
SegKernel bulk_kernel_min(const TileShape& t, int stages) { return pick_bulk_kernel<OP_MIN>(t, stages); }
SegKernel lean256_kernel_min(int g, bool ragged, bool noval) { return pick_lean<OP_MIN, 8>(g, ragged, noval); }
SegKernel lean128_kernel_min(int g, bool ragged, bool noval) { return pick_lean<OP_MIN, 4>(g, ragged, noval); }
}  // namespace isplib
