// Shared declarations of the fused tensor-core field kernel: constants, launch arguments, launchers (one translation unit per
// (planes, table layout) instantiation: csrc/field_tc_p*_*.cu) and the host side (csrc/field_tc.cu).
#pragma once
#include "field_plan.h"

namespace sdfb200 {


constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGatherWarps = 4;                              // hash-grid part of the encode: one thread per point
#ifndef TCV_ALU_WARPS
#define TCV_ALU_WARPS 2
#endif
constexpr int kAluWarps = TCV_ALU_WARPS;                     // 0: the gather warps also write the PE / x / static colour columns; 2: two extra warps do (two points per thread)
constexpr int kEncWarps = kGatherWarps + kAluWarps;
constexpr int kWarpProducer = kEpiWarps + kEncWarps;         // 14
constexpr int kWarpMma = kWarpProducer + 1;                  // 15
constexpr int kTcThreads = (kWarpMma + 1) * 32;              // 512: register allocation rounds the warp count up to a multiple of 4 anyway
#ifndef TCV_STAGES
#define TCV_STAGES 4
#endif
constexpr int kStages = TCV_STAGES;      // weight ring depth (16 KB per stage and CTA at two planes).  4, not 7: the 48 KB go to L1 (carve-out 228 -> 196 KB), which
                                         // the hash gathers of the coarse levels need more than the MMA issuer needs look-ahead (A/B: 7 -> 1.59 ms, 4 -> 1.47 ms)
constexpr int kKB = 32;           // K per streamed weight block of the 256-row layers (one 16 KB stage per CTA at two planes)
constexpr int kKBMax = 64;        // the 96-row layer streams K blocks of 64 (12 KB): deeper prefetch in bytes for the short layers
constexpr int kMaxGridDim = 32;
constexpr int kMaxPe = 60;        // PE columns (2 * 3 * degree), degree <= 10
constexpr int kPeRows = 64;       // rows reserved for the PE jacobian in the scratch
constexpr int kInK = 96;          // padded K of the two small-K operands (geo input, colour misc input)
constexpr float kHalfPiF = 1.5707963267948966f;
// kernel order of the geo input columns (K = 96): [grid features 0..31 | PE | x(3) | zero padding] -- every group of four hash
// levels is one aligned 16-byte operand chunk.  W0 (columns) and W0^T (rows) are permuted accordingly at pack time.
// per-CTA scratch: softplus'(z1) unorm16 [64 KB] | h2 planes [P x 64 KB] | 2 x input jacobian (PE [64][128] f32 | grid [96][128] f32)
constexpr size_t kJRBytes = (size_t)(kPeRows + kMaxGridDim * 3) * 128 * 4;
__host__ __device__ constexpr size_t kScratchPerCta(int planes) { return 65536 + (size_t)planes * 65536 + 2 * kJRBytes; }

enum { L_G0 = 0, L_G1, L_B1, L_B0, L_C0H, L_C0MISC, L_C1, L_COUNT };

struct TcLayer {
  unsigned long long w_off;  // byte offset of the packed planes inside the blob
  int Np;                    // rows of the weight tile (UMMA N); each CTA of a pair holds Np / 2 of them
  int nkb;                   // number of K blocks
  int kblk;                  // K per block (32 or 64)
};

struct TcArgs {
  sdfb200_grid_t grid;
  TcLayer layer[L_COUNT];
  int use_grid, pe_degree, use_pe, contraction, in_dim, pe_dim, grid_dim, app_dim, use_n_dot_v;
  int mode;  // 0: sdf only (G0, G1)   1: everything
  int n_samples, has_bins, n_tiles, n_tile_pairs;
  long long n_points;
  float rgb_padding, cos_anneal;
  const float *origins, *directions, *bins, *appearance, *variance, *beta, *beta_min;
  const void* table;
  const char* blob;
  // fp32 section offsets (bytes)
  unsigned long long b_g0, b_g1, b_g2, w_g2, b_c0, b_c1, w_c2, b_c2;   // b_c0 = fused bias (bc0 + Wgf b2')
  char* scratch;
  unsigned long long scratch_per_cta;
  sdfb200_field_out_t out;
  TcRender rnd;
};

// one launcher per instantiation; grid = 2 x CTA pairs
int launch_field_tc_p2_torch(const TcArgs& a, int grid, size_t smem, cudaStream_t st);
int launch_field_tc_p2_tcnn(const TcArgs& a, int grid, size_t smem, cudaStream_t st);
int launch_field_tc_p1_torch(const TcArgs& a, int grid, size_t smem, cudaStream_t st);
int launch_field_tc_p1_tcnn(const TcArgs& a, int grid, size_t smem, cudaStream_t st);

}  // namespace sdfb200
