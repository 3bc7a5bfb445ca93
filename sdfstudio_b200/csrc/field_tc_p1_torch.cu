// instantiation of the fused tensor-core field kernel: bf16 (one plane), torch-layout hash table
#include "field_tc_kernel.cuh"

namespace sdfb200 {
int launch_field_tc_p1_torch(const TcArgs& a, int grid, size_t smem, cudaStream_t st) { return launch_field_tc<1, SDFB200_GRID_TORCH>(a, grid, smem, st); }
}  // namespace sdfb200
