// instantiation of the fused tensor-core field kernel: bf16x3 (two bf16 planes), tcnn-layout hash table
#include "field_tc_kernel.cuh"

namespace sdfb200 {
int launch_field_tc_p2_tcnn(const TcArgs& a, int grid, size_t smem, cudaStream_t st) { return launch_field_tc<2, SDFB200_GRID_TCNN>(a, grid, smem, st); }
}  // namespace sdfb200
