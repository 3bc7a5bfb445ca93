// Fused tcgen05 evaluation of the SDF field (SDFB200_PRECISION_BF16X3 / _BF16) for the neus-facto family of shapes:
// geo MLP in-256-256-(1+256), colour MLP cin-256-256-3, analytic d sdf/dx, 2-feature fp32 hash grid.
//
// One persistent CTA per SM walks 128-point tiles.  Per tile (everything stays on chip except two L2-resident spills):
//   encode   16 epilogue warps: position, contraction, PE, hash gathers (+ jacobian) -> bf16 split planes in smem
//   G0 G1    h = softplus_100(W a + b)       accumulator in TMEM (256 cols), next layer's A operand written to TMEM
//   sdf      fp32 dot of h2 with row 0 of W2 on CUDA cores (exact fp32: the SDF drives NeuS alpha / Laplace density)
//   G2       geo feature (256)               spilled (bf16 planes) to a per-CTA L2-resident scratch
//   B1 B0    reverse sweep: g2 = W2[0,:]*sp'(z2), g1 = (W1^T g2)*sp'(z1), gin = W0^T g1;  sp'(z1) spilled at G0
//   grad     d sdf/dx = gin_x + PE jacobian + grid jacobian / 4      (what autograd computes at sdf_field.py:647-654)
//   C0 C1    relu MLP on [x, dir-enc, grad, geo feature, appearance]; last 256->3 layer as fp32 dots; sigmoid + padding
//   heads    Laplace density, NeuS alpha, occupancy, normals
// MMA = tcgen05.mma kind::f16 (bf16 x bf16 -> fp32), M=128.  bf16x3: a0*w0 + a1*w0 + a0*w1 with a = a0+a1, w = w0+w1
// (error ~2^-16 relative, fp32 accumulate).  Weights stream through a 3-stage shared-memory ring filled by 1-D bulk
// copies (UBLKCP) from a pre-packed image; warp 16 = producer, warp 17 = MMA issuer, warps 0-15 = encode + epilogues.
#include "field_plan.h"
#include "grid.cuh"
#include "tc_common.cuh"

namespace sdfb200 {
using namespace tc;

constexpr int kTcThreads = 576;
constexpr int kEpiThreads = 512;
constexpr int kStages = 3;
constexpr int kKB = 32;           // K per streamed weight block
constexpr int kInK = 96;          // padded K of the two small-K operands (geo input, colour misc input)
constexpr int kMaxGridDim = 32;
constexpr float kHalfPiF = 1.5707963267948966f;

enum { L_G0 = 0, L_G1, L_G2, L_B1, L_B0, L_C0GF, L_C0MISC, L_C1, L_COUNT };

struct TcLayer {
  unsigned long long w_off;  // byte offset of the packed planes inside the blob
  int Np;                    // rows of the weight tile (UMMA N)
  int nkb;                   // K blocks of 32
};

struct TcArgs {
  sdfb200_grid_t grid;
  TcLayer layer[L_COUNT];
  int use_grid, pe_degree, use_pe, contraction, in_dim, pe_dim, grid_dim, cm_dim, app_dim, use_n_dot_v;
  int mode;  // 0: sdf only (G0, G1)   1: everything
  int n_samples, has_bins, n_tiles;
  long long n_points;
  float rgb_padding, cos_anneal;
  const float *origins, *directions, *bins, *appearance, *variance, *beta, *beta_min;
  const void* table;
  const char* blob;
  // fp32 section offsets (bytes)
  unsigned long long b_g0, b_g1, b_g2, w_g2, b_c0, b_c1, w_c2, b_c2;
  char* scratch;
  unsigned long long scratch_per_cta;
  sdfb200_field_out_t out;
};

// pack fp32 W (row n, column k at W[n*ldw + colmap(k)]) into bf16 split planes, K-blocked canonical layout
__global__ void k_tc_pack(const float* __restrict__ W, int ldw, int N, int K, int Np, int nblocks, int planes, int split, int skip,
                          __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nblocks * Np * kKB) return;
  const int kk = idx % kKB;
  const int n = (idx / kKB) % Np;
  const int b = idx / (kKB * Np);
  const int k = b * kKB + kk;
  const int src = k < split ? k : k + skip;
  const float w = (n < N && k < K) ? W[(size_t)n * ldw + src] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(w);
  const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
  const size_t plane_elems = (size_t)Np * kKB;
  const size_t base = (size_t)b * planes * plane_elems;
  const size_t off = (size_t)(kk / 8) * (Np * 8) + (size_t)n * 8 + (kk % 8);
  out[base + off] = hi;
  if (planes > 1) out[base + plane_elems + off] = lo;
}

__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// store one bf16 element (split into P planes) of the small-K smem operand: layout [plane][k/8][row][8]
template <int P>
__device__ __forceinline__ void store_in(uint8_t* inA, int row, int col, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const uint32_t off = (uint32_t)(col >> 3) * 2048u + (uint32_t)row * 16u + (uint32_t)(col & 7) * 2u;
  *reinterpret_cast<__nv_bfloat16*>(inA + off) = hi;
  if (P > 1) *reinterpret_cast<__nv_bfloat16*>(inA + (kInK / 8) * 2048 + off) = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int P>
__global__ void __launch_bounds__(kTcThreads, 1) k_field_tc(const __grid_constant__ TcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr uint32_t kInBytes = (uint32_t)P * (kInK / 8) * 2048;       // small-K operand (all planes)
  constexpr uint32_t kJBytes = kMaxGridDim * 3 * 128 * 4;              // grid jacobian [c][d][row] fp32
  constexpr uint32_t kStageBytes = (uint32_t)P * 256 * kKB * 2;        // one weight K-block, all planes
  uint8_t* inA = smem;
  float* Jbuf = reinterpret_cast<float*>(smem + kInBytes);
  uint8_t* ring = smem + kInBytes + kJBytes;
  float* fbuf = reinterpret_cast<float*>(ring + kStages * kStageBytes);
  float* xbuf = fbuf;                 // [3][128]   contracted position
  float* gradbuf = fbuf + 3 * 128;    // [3][128]
  float* sdfbuf = fbuf + 6 * 128;     // [128]
  float* red = fbuf + 7 * 128;        // [3][4][128] partial sums
  __shared__ uint64_t full[kStages], empty[kStages], dfull;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&dfull, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t d_tmem = tmem;          // accumulator: columns [0,256)
  const uint32_t a_tmem = tmem + 256;    // A planes: plane p at columns 256 + 128 p
  const int nphase_layers = a.mode == 0 ? 2 : L_COUNT;

  if (warp == 16) {
    // ============================== weight producer ==============================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int L = 0; L < nphase_layers; ++L) {
          const TcLayer ly = a.layer[L];
          const uint32_t bytes = (uint32_t)P * ly.Np * kKB * 2;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(a.blob) + ly.w_off;
          for (int kb = 0; kb < ly.nkb; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(&empty[s], ((it / kStages) & 1) ^ 1);
            mbar_arrive_expect_tx(&full[s], bytes);
            bulk_g2s(ring + (size_t)s * kStageBytes, src + (size_t)kb * bytes, bytes, &full[s]);
          }
        }
      }
    }
  } else if (warp == 17) {
    // ============================== MMA issuer ==============================
    uint32_t it = 0;
    const uint32_t in_base = smem_u32(inA);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int L = 0; L < nphase_layers; ++L) {
        const bool continues = (L == L_C0MISC);           // accumulates onto C0GF, no barrier in between
        if (!continues) {
          named_sync(1, kEpiThreads + 32);                // "A operand ready" from the epilogue warps
          tc_fence_after();
        }
        if (lane == 0) {
          const TcLayer ly = a.layer[L];
          const bool a_in_smem = (L == L_G0 || L == L_C0MISC);
          const uint32_t idesc = make_idesc_bf16(128, ly.Np);
          const uint32_t lbo_b = (uint32_t)ly.Np * 16, plane_b = (uint32_t)ly.Np * kKB * 2;
          uint32_t acc = continues ? 1u : 0u;
          for (int kb = 0; kb < ly.nkb; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(&full[s], (it / kStages) & 1);
            tc_fence_after();
            const uint32_t wbase = smem_u32(ring + (size_t)s * kStageBytes);
#pragma unroll
            for (int j = 0; j < kKB / 16; ++j) {
              const int kstep = kb * (kKB / 16) + j;
              const uint64_t b0 = make_smem_desc(wbase + j * 2 * lbo_b, lbo_b, 128);
              const uint64_t b1 = make_smem_desc(wbase + plane_b + j * 2 * lbo_b, lbo_b, 128);
              if (a_in_smem) {
                const uint64_t a0 = make_smem_desc(in_base + kstep * 2 * 2048, 2048, 128);
                mma_ss(d_tmem, a0, b0, idesc, acc);
                acc = 1;
                if (P > 1) {
                  const uint64_t a1 = make_smem_desc(in_base + (kInK / 8) * 2048 + kstep * 2 * 2048, 2048, 128);
                  mma_ss(d_tmem, a1, b0, idesc, 1);
                  mma_ss(d_tmem, a0, b1, idesc, 1);
                }
              } else {
                mma_ts(d_tmem, a_tmem + kstep * 8, b0, idesc, acc);
                acc = 1;
                if (P > 1) {
                  mma_ts(d_tmem, a_tmem + 128 + kstep * 8, b0, idesc, 1);
                  mma_ts(d_tmem, a_tmem + kstep * 8, b1, idesc, 1);
                }
              }
            }
            mma_commit(&empty[s]);
          }
          if (L != L_C0GF) mma_commit(&dfull);            // C0GF is completed by C0MISC
        }
        __syncwarp();
      }
    }
  } else {
    // ============================== encode + epilogues (16 warps) ==============================
    const int row = (warp & 3) * 32 + lane;               // tile row == TMEM lane
    const int q = warp >> 2;                              // column quarter
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const char* blob = a.blob;
    const float* b_g0 = reinterpret_cast<const float*>(blob + a.b_g0);
    const float* b_g1 = reinterpret_cast<const float*>(blob + a.b_g1);
    const float* b_g2 = reinterpret_cast<const float*>(blob + a.b_g2);
    const float* w_g2 = reinterpret_cast<const float*>(blob + a.w_g2);   // row 0 of the last geo layer
    const float* b_c0 = reinterpret_cast<const float*>(blob + a.b_c0);
    const float* b_c1 = reinterpret_cast<const float*>(blob + a.b_c1);
    const float* w_c2 = reinterpret_cast<const float*>(blob + a.w_c2);   // [3 rows][256]
    const float* b_c2 = reinterpret_cast<const float*>(blob + a.b_c2);
    float* sig_s = reinterpret_cast<float*>(a.scratch + (size_t)blockIdx.x * a.scratch_per_cta);        // [64 units][128 rows][4]
    uint8_t* gf_s = reinterpret_cast<uint8_t*>(sig_s) + 131072;                                          // [P][32 units][128][16 B]
    uint32_t dpar = 0;
    const int deg = a.pe_degree;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      const long long p_raw = (long long)tile * 128 + row;
      const bool valid = p_raw < a.n_points;
      const long long p = valid ? p_raw : a.n_points - 1;
      const long long ray = a.has_bins ? p / a.n_samples : p;
      const int smp = a.has_bins ? (int)(p - ray * a.n_samples) : 0;
      // ---------------- position (all four quarter-threads of a row compute it) ----------------
      float px, py, pz, dirx = 0.f, diry = 0.f, dirz = 0.f, delta = 0.f;
      if (a.has_bins) {
        const float t0 = __ldg(a.bins + ray * (a.n_samples + 1) + smp);
        delta = __fsub_rn(__ldg(a.bins + ray * (a.n_samples + 1) + smp + 1), t0);
        dirx = __ldg(a.directions + ray * 3); diry = __ldg(a.directions + ray * 3 + 1); dirz = __ldg(a.directions + ray * 3 + 2);
        px = __fadd_rn(__ldg(a.origins + ray * 3 + 0), __fmul_rn(dirx, t0));
        py = __fadd_rn(__ldg(a.origins + ray * 3 + 1), __fmul_rn(diry, t0));
        pz = __fadd_rn(__ldg(a.origins + ray * 3 + 2), __fmul_rn(dirz, t0));
      } else {
        px = __ldg(a.origins + p * 3); py = __ldg(a.origins + p * 3 + 1); pz = __ldg(a.origins + p * 3 + 2);
        if (a.directions) { dirx = __ldg(a.directions + p * 3); diry = __ldg(a.directions + p * 3 + 1); dirz = __ldg(a.directions + p * 3 + 2); }
      }
      if (a.contraction != SDFB200_CONTRACT_NONE) {
        const float mag = a.contraction == SDFB200_CONTRACT_LINF ? fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz)))
                                                                  : sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
        if (mag >= 1.f) {
          const float k = __fsub_rn(2.f, __fdiv_rn(1.f, mag));
          px = __fmul_rn(k, __fdiv_rn(px, mag)); py = __fmul_rn(k, __fdiv_rn(py, mag)); pz = __fmul_rn(k, __fdiv_rn(pz, mag));
        }
      }
      const float pc[3] = {px, py, pz};
      // ---------------- geo input -> inA (bf16 planes) ----------------
      if (q == 0) {
        store_in<P>(inA, row, 0, px); store_in<P>(inA, row, 1, py); store_in<P>(inA, row, 2, pz);
        if (valid) {
          if (a.out.points_norm) a.out.points_norm[p] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz)));
          if (a.out.points) { a.out.points[p * 3] = px; a.out.points[p * 3 + 1] = py; a.out.points[p * 3 + 2] = pz; }
        }
      }
      if (q == 1)
        for (int c = a.in_dim; c < kInK; ++c) store_in<P>(inA, row, c, 0.f);
      {
        const int half = 3 * deg;
        for (int i = q; i < half; i += 4) {               // PE: sin(x 2^k) | sin(x 2^k + pi/2)   (encodings.py:194-198)
          const int b = i / deg, k = i - b * deg;
          const float sarg = pc[b] * (float)(1 << k);
          store_in<P>(inA, row, 3 + i, a.use_pe ? sinf(sarg) : 0.f);
          store_in<P>(inA, row, 3 + half + i, a.use_pe ? sinf(sarg + kHalfPiF) : 0.f);
        }
      }
      if (a.use_grid) {
        const float x01 = (px + 2.0f) * 0.25f, y01 = (py + 2.0f) * 0.25f, z01 = (pz + 2.0f) * 0.25f;
        for (int l = q; l < a.grid.n_levels; l += 4) {
          float o[2];
          float dj[2][3];
          if (l < a.grid.active_levels) {
            encode_level<float, 2, true>(a.grid, a.table, l, x01, y01, z01, o, dj);
          } else {
            o[0] = o[1] = 0.f;
            dj[0][0] = dj[0][1] = dj[0][2] = dj[1][0] = dj[1][1] = dj[1][2] = 0.f;
          }
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            const int c = l * 2 + f;
            store_in<P>(inA, row, 3 + a.pe_dim + c, o[f]);
            Jbuf[(c * 3 + 0) * 128 + row] = dj[f][0]; Jbuf[(c * 3 + 1) * 128 + row] = dj[f][1]; Jbuf[(c * 3 + 2) * 128 + row] = dj[f][2];
          }
        }
      } else {
        for (int c = q; c < a.grid_dim; c += 4) store_in<P>(inA, row, 3 + a.pe_dim + c, 0.f);
      }
      fence_async_smem();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- E0: h1 = softplus(z1) -> A planes ; softplus'(z1) -> scratch ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = q * 64 + cc * 16;
        uint32_t v[16];
        tmem_ld16(d_tmem + lane_addr + col0, v);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
        float sg[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float h[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float z = __uint_as_float(v[j + u]) + __ldg(b_g0 + col0 + j + u);
            const float t = z * 100.0f;
            const float e = expf(t);
            h[u] = t > 20.0f ? z : log1pf(e) * 0.01f;
            sg[j + u] = t > 20.0f ? 1.0f : __fdividef(e, 1.0f + e);
          }
          split2(h[0], h[1], hi[j >> 1], lo[j >> 1]);
        }
        tmem_st8(a_tmem + lane_addr + (col0 >> 1), hi);
        if (P > 1) tmem_st8(a_tmem + 128 + lane_addr + (col0 >> 1), lo);
        if (a.mode != 0) {
#pragma unroll
          for (int u4 = 0; u4 < 4; ++u4)
            *reinterpret_cast<float4*>(sig_s + ((size_t)((col0 >> 2) + u4) * 128 + row) * 4) = make_float4(sg[4 * u4], sg[4 * u4 + 1], sg[4 * u4 + 2], sg[4 * u4 + 3]);
        }
      }
      tc_wait_st();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- E1: h2 -> A planes ; sdf = W2[0,:] . h2 + b (fp32) ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
      float sdf_part = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = q * 64 + cc * 16;
        uint32_t v[16];
        tmem_ld16(d_tmem + lane_addr + col0, v);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float h[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float z = __uint_as_float(v[j + u]) + __ldg(b_g1 + col0 + j + u);
            const float t = z * 100.0f;
            h[u] = t > 20.0f ? z : log1pf(expf(t)) * 0.01f;
            sdf_part = fmaf(__ldg(w_g2 + col0 + j + u), h[u], sdf_part);
          }
          split2(h[0], h[1], hi[j >> 1], lo[j >> 1]);
        }
        tmem_st8(a_tmem + lane_addr + (col0 >> 1), hi);
        if (P > 1) tmem_st8(a_tmem + 128 + lane_addr + (col0 >> 1), lo);
      }
      tc_wait_st();
      red[q * 128 + row] = sdf_part;
      tc_fence_before();
      if (a.mode != 0) named_arrive(1, kEpiThreads + 32);
      named_sync(2, kEpiThreads);
      float sdf = (red[row] + red[128 + row]) + (red[256 + row] + red[384 + row]) + __ldg(b_g2);
      if (q == 0 && valid && a.out.sdf) a.out.sdf[p] = sdf;
      if (a.mode == 0) {
        named_sync(2, kEpiThreads);  // `red` is reused by the next tile
        continue;
      }

      // ---------------- E2: geo feature -> scratch planes ; g2 = W2[0,:] * softplus'(z2) -> A planes ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = q * 64 + cc * 16;
        uint32_t v[16];
        tmem_ld16(d_tmem + lane_addr + col0, v);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float g0 = __uint_as_float(v[j]) + __ldg(b_g2 + 1 + col0 + j);
          const float g1 = __uint_as_float(v[j + 1]) + __ldg(b_g2 + 1 + col0 + j + 1);
          if (a.out.geo_feature && valid) { a.out.geo_feature[p * 256 + col0 + j] = g0; a.out.geo_feature[p * 256 + col0 + j + 1] = g1; }
          split2(g0, g1, hi[j >> 1], lo[j >> 1]);
        }
#pragma unroll
        for (int u8 = 0; u8 < 2; ++u8) {
          const size_t unit = ((size_t)((col0 >> 3) + u8) * 128 + row) * 16;
          *reinterpret_cast<uint4*>(gf_s + unit) = make_uint4(hi[4 * u8], hi[4 * u8 + 1], hi[4 * u8 + 2], hi[4 * u8 + 3]);
          if (P > 1) *reinterpret_cast<uint4*>(gf_s + 65536 + unit) = make_uint4(lo[4 * u8], lo[4 * u8 + 1], lo[4 * u8 + 2], lo[4 * u8 + 3]);
        }
        // h2 (A planes) -> g2 in place.  softplus'(z) = 1 - exp(-100 h)
        uint32_t h_hi[8], h_lo[8];
        tmem_ld8(a_tmem + lane_addr + (col0 >> 1), h_hi);
        if (P > 1) tmem_ld8(a_tmem + 128 + lane_addr + (col0 >> 1), h_lo);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float ha = bf16lo_to_f32(h_hi[j]), hb = bf16hi_to_f32(h_hi[j]);
          if (P > 1) { ha += bf16lo_to_f32(h_lo[j]); hb += bf16hi_to_f32(h_lo[j]); }
          const float ga = __ldg(w_g2 + col0 + 2 * j) * -expm1f(-100.0f * ha);
          const float gb = __ldg(w_g2 + col0 + 2 * j + 1) * -expm1f(-100.0f * hb);
          split2(ga, gb, hi[j], lo[j]);
        }
        tmem_st8(a_tmem + lane_addr + (col0 >> 1), hi);
        if (P > 1) tmem_st8(a_tmem + 128 + lane_addr + (col0 >> 1), lo);
      }
      tc_wait_st();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- EB1: g1 = (W1^T g2) * softplus'(z1) -> A planes ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = q * 64 + cc * 16;
        uint32_t v[16];
        tmem_ld16(d_tmem + lane_addr + col0, v);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sig_s + ((size_t)((col0 >> 2) + u4) * 128 + row) * 4);
          split2(__uint_as_float(v[4 * u4]) * s4.x, __uint_as_float(v[4 * u4 + 1]) * s4.y, hi[2 * u4], lo[2 * u4]);
          split2(__uint_as_float(v[4 * u4 + 2]) * s4.z, __uint_as_float(v[4 * u4 + 3]) * s4.w, hi[2 * u4 + 1], lo[2 * u4 + 1]);
        }
        tmem_st8(a_tmem + lane_addr + (col0 >> 1), hi);
        if (P > 1) tmem_st8(a_tmem + 128 + lane_addr + (col0 >> 1), lo);
      }
      tc_wait_st();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- EB0: gin (96 cols) -> d sdf / dx ; colour misc input ; reload geo feature ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
      {
        float gx = 0.f, gy = 0.f, gz = 0.f;
        const int half = 3 * deg;
#pragma unroll 1
        for (int c8 = 0; c8 < 3; ++c8) {
          const int c0 = q * 24 + c8 * 8;
          uint32_t v[8];
          tmem_ld8(d_tmem + lane_addr + c0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            const float g = __uint_as_float(v[j]);
            if (c < 3) {
              if (c == 0) gx += g; else if (c == 1) gy += g; else gz += g;
            } else if (c < 3 + a.pe_dim) {
              if (a.use_pe) {
                int i = c - 3;
                const bool second = i >= half;
                if (second) i -= half;
                const int b = i / deg, k = i - b * deg;
                const float fr = (float)(1 << k);
                const float sarg = pc[b] * fr;
                const float dv = fr * g * cosf(second ? sarg + kHalfPiF : sarg);   // autograd of sin on the forward's fp32 argument
                if (b == 0) gx += dv; else if (b == 1) gy += dv; else gz += dv;
              }
            } else if (c < a.in_dim) {
              if (a.use_grid) {
                const int cg = c - 3 - a.pe_dim;
                const float g4 = 0.25f * g;                                        // positions = (x + 2) / 4
                gx = fmaf(g4, Jbuf[(cg * 3 + 0) * 128 + row], gx);
                gy = fmaf(g4, Jbuf[(cg * 3 + 1) * 128 + row], gy);
                gz = fmaf(g4, Jbuf[(cg * 3 + 2) * 128 + row], gz);
              }
            }
          }
        }
        red[(0 * 4 + q) * 128 + row] = gx; red[(1 * 4 + q) * 128 + row] = gy; red[(2 * 4 + q) * 128 + row] = gz;
      }
      named_sync(2, kEpiThreads);
      const float grx = (red[(0 * 4 + 0) * 128 + row] + red[(0 * 4 + 1) * 128 + row]) + (red[(0 * 4 + 2) * 128 + row] + red[(0 * 4 + 3) * 128 + row]);
      const float gry = (red[(1 * 4 + 0) * 128 + row] + red[(1 * 4 + 1) * 128 + row]) + (red[(1 * 4 + 2) * 128 + row] + red[(1 * 4 + 3) * 128 + row]);
      const float grz = (red[(2 * 4 + 0) * 128 + row] + red[(2 * 4 + 1) * 128 + row]) + (red[(2 * 4 + 2) * 128 + row] + red[(2 * 4 + 3) * 128 + row]);
      const float gn = fmaxf(sqrtf(grx * grx + gry * gry + grz * grz), 1e-12f);      // F.normalize eps
      const float nx = grx / gn, ny = gry / gn, nz = grz / gn;
      // colour misc input [x(3) | dir-enc(27) | grad(3) | appearance | n.v | 0...]   (sdf_field.py:572-584)
#pragma unroll 1
      for (int j = 0; j < 24; ++j) {
        const int c = q * 24 + j;
        float val = 0.f;
        if (c < 3) val = pc[c];
        else if (c < 30) {
          const int i = c - 3;
          const float dd[3] = {dirx, diry, dirz};
          if (i < 12) { const int b = i >> 2, k = i & 3; val = sinf(dd[b] * (float)(1 << k)); }
          else if (i < 24) { const int b = (i - 12) >> 2, k = (i - 12) & 3; val = sinf(dd[b] * (float)(1 << k) + kHalfPiF); }
          else val = dd[i - 24];
        } else if (c < 33) val = c == 30 ? grx : (c == 31 ? gry : grz);
        else if (c < 33 + a.app_dim) val = a.appearance ? __ldg(a.appearance + ray * a.app_dim + (c - 33)) : 0.f;
        else if (a.use_n_dot_v && c == 33 + a.app_dim) val = nx * dirx + ny * diry + nz * dirz;
        store_in<P>(inA, row, c, val);
      }
      // geo feature planes back into the A operand
#pragma unroll 1
      for (int u = 0; u < 8; ++u) {
        const size_t unit = ((size_t)(q * 8 + u) * 128 + row) * 16;
        const uint4 h4 = *reinterpret_cast<const uint4*>(gf_s + unit);
        const uint32_t hh[4] = {h4.x, h4.y, h4.z, h4.w};
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_tmem + lane_addr + (q * 8 + u) * 4), "r"(hh[0]), "r"(hh[1]), "r"(hh[2]), "r"(hh[3]) : "memory");
        if (P > 1) {
          const uint4 l4 = *reinterpret_cast<const uint4*>(gf_s + 65536 + unit);
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a_tmem + 128 + lane_addr + (q * 8 + u) * 4), "r"(l4.x), "r"(l4.y), "r"(l4.z), "r"(l4.w) : "memory");
        }
      }
      tc_wait_st();
      fence_async_smem();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- EC0: relu -> A planes ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int col0 = q * 64 + cc * 16;
        uint32_t v[16];
        tmem_ld16(d_tmem + lane_addr + col0, v);
        tc_wait_ld();
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          split2(fmaxf(__uint_as_float(v[j]) + __ldg(b_c0 + col0 + j), 0.f), fmaxf(__uint_as_float(v[j + 1]) + __ldg(b_c0 + col0 + j + 1), 0.f), hi[j >> 1], lo[j >> 1]);
        tmem_st8(a_tmem + lane_addr + (col0 >> 1), hi);
        if (P > 1) tmem_st8(a_tmem + 128 + lane_addr + (col0 >> 1), lo);
      }
      tc_wait_st();
      tc_fence_before();
      named_arrive(1, kEpiThreads + 32);

      // ---------------- EC1: relu, last colour layer (256 -> 3) as fp32 dots ----------------
      mbar_wait(&dfull, dpar); dpar ^= 1; tc_fence_after();
      {
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int col0 = q * 64 + cc * 16;
          uint32_t v[16];
          tmem_ld16(d_tmem + lane_addr + col0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float c1 = fmaxf(__uint_as_float(v[j]) + __ldg(b_c1 + col0 + j), 0.f);
            r0 = fmaf(__ldg(w_c2 + col0 + j), c1, r0);
            r1 = fmaf(__ldg(w_c2 + 256 + col0 + j), c1, r1);
            r2 = fmaf(__ldg(w_c2 + 512 + col0 + j), c1, r2);
          }
        }
        named_sync(2, kEpiThreads);   // everyone has consumed the gradient partials in `red`
        red[(0 * 4 + q) * 128 + row] = r0; red[(1 * 4 + q) * 128 + row] = r1; red[(2 * 4 + q) * 128 + row] = r2;
      }
      tc_fence_before();
      named_sync(2, kEpiThreads);
      if (q == 0 && valid) {
        // ---------------- per-point heads ----------------
        if (a.out.rgb) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float raw = (red[(c * 4 + 0) * 128 + row] + red[(c * 4 + 1) * 128 + row]) + (red[(c * 4 + 2) * 128 + row] + red[(c * 4 + 3) * 128 + row]) + __ldg(b_c2 + c);
            a.out.rgb[p * 3 + c] = sigmoidf_(raw) * (1.f + 2.f * a.rgb_padding) - a.rgb_padding;
          }
        }
        if (a.out.gradients) { a.out.gradients[p * 3] = grx; a.out.gradients[p * 3 + 1] = gry; a.out.gradients[p * 3 + 2] = grz; }
        if (a.out.normals) { a.out.normals[p * 3] = nx; a.out.normals[p * 3 + 1] = ny; a.out.normals[p * 3 + 2] = nz; }
        if (a.out.density) {
          const float beta = fabsf(__ldg(a.beta)) + __ldg(a.beta_min);
          const float sg = sdf > 0.f ? 1.f : (sdf < 0.f ? -1.f : 0.f);
          a.out.density[p] = (1.0f / beta) * (0.5f + 0.5f * sg * expm1f(-fabsf(sdf) / beta));
        }
        if (a.out.occupancy) a.out.occupancy[p] = sigmoidf_(-10.0f * sdf);
        if (a.out.alpha) {
          const float inv_s = fminf(fmaxf(expf(__ldg(a.variance) * 10.0f), 1e-6f), 1e6f);
          const float true_cos = dirx * grx + diry * gry + dirz * grz;
          const float iter_cos = -(fmaxf(-true_cos * 0.5f + 0.5f, 0.f) * (1.0f - a.cos_anneal) + fmaxf(-true_cos, 0.f) * a.cos_anneal);
          const float prev_cdf = sigmoidf_((sdf - iter_cos * delta * 0.5f) * inv_s), next_cdf = sigmoidf_((sdf + iter_cos * delta * 0.5f) * inv_s);
          a.out.alpha[p] = fminf(fmaxf((prev_cdf - next_cdf + 1e-5f) / (prev_cdf + 1e-5f), 0.f), 1.f);
        }
      }
      named_sync(2, kEpiThreads);     // `red` / inA / Jbuf are rewritten by the next tile
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// -----------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------
static size_t tc_layer_bytes(int planes, int Np, int nkb) { return (size_t)nkb * planes * Np * kKB * 2; }

struct TcPlan {
  int planes;
  TcLayer layer[L_COUNT];
  size_t total;
  int cm_dim;
};

static void make_tc_plan(const sdfb200_field_t& f, const FieldPlan& p, TcPlan& t) {
  t.planes = f.precision == SDFB200_PRECISION_BF16X3 ? 2 : 1;
  t.cm_dim = 3 + 27 + 3 + f.appearance_dim + (f.use_n_dot_v ? 1 : 0);
  const int np[L_COUNT] = {256, 256, 256, 256, kInK, 256, 256, 256};
  const int nkb[L_COUNT] = {kInK / kKB, 8, 8, 8, 8, 8, kInK / kKB, 8};
  size_t off = p.tc_off;
  for (int l = 0; l < L_COUNT; ++l) {
    t.layer[l].w_off = off;
    t.layer[l].Np = np[l];
    t.layer[l].nkb = nkb[l];
    off = align_up(off + tc_layer_bytes(t.planes, np[l], nkb[l]), 256);
  }
  t.total = off - p.tc_off;
}

bool field_tc_supported(const sdfb200_field_t& f, const FieldPlan& p) {
  if (f.precision != SDFB200_PRECISION_BF16X3 && f.precision != SDFB200_PRECISION_BF16) return false;
  if (p.n_geo != 3 || p.n_col != 3) return false;
  if (p.geo[0].N != 256 || p.geo[1].N != 256 || p.geo_feat != 256 || p.col[0].N != 256 || p.col[1].N != 256) return false;
  if (f.use_numerical_gradients || f.off_axis || f.use_diffuse_color || f.use_specular_tint || f.use_reflections) return false;
  if (p.in_dim > kInK || p.grid_dim > kMaxGridDim) return false;
  if (f.use_grid_feature && (f.grid.n_features != 2 || f.grid.table_dtype != SDFB200_DT_F32)) return false;
  const int cm = 3 + 27 + 3 + f.appearance_dim + (f.use_n_dot_v ? 1 : 0);
  if (cm > kInK) return false;
  return true;
}

size_t field_tc_packed_bytes(const sdfb200_field_t& f, const FieldPlan& p) {
  TcPlan t;
  make_tc_plan(f, p, t);
  return t.total;
}

constexpr size_t kScratchPerCta(int planes) { return 131072 + (size_t)planes * 65536; }

size_t field_tc_workspace_floats(const sdfb200_field_t& f, const FieldPlan&, int64_t) {
  const int planes = f.precision == SDFB200_PRECISION_BF16X3 ? 2 : 1;
  return kNumSMs * kScratchPerCta(planes) / sizeof(float) + 64;
}

int field_tc_pack(const sdfb200_field_t& f, const FieldPlan& p, char* blob, cudaStream_t st) {
  TcPlan t;
  make_tc_plan(f, p, t);
  auto pack = [&](int L, const float* W, int ldw, int N, int K, int split, int skip) -> int {
    const TcLayer& ly = t.layer[L];
    const int tot = ly.nkb * ly.Np * kKB;
    k_tc_pack<<<(tot + 255) / 256, 256, 0, st>>>(W, ldw, N, K, ly.Np, ly.nkb, t.planes, split, skip, (__nv_bfloat16*)(blob + ly.w_off));
    SDFB_LAUNCHED("k_tc_pack");
    return 0;
  };
  const LayerPlan &g0 = p.geo[0], &g1 = p.geo[1], &g2 = p.geo[2], &c0 = p.col[0], &c1 = p.col[1];
  const int big = 1 << 30;
  int r;
  if ((r = pack(L_G0, (const float*)(blob + g0.w_off), g0.Kp, 256, g0.K, big, 0))) return r;
  if ((r = pack(L_G1, (const float*)(blob + g1.w_off), g1.Kp, 256, 256, big, 0))) return r;
  if ((r = pack(L_G2, (const float*)(blob + g2.w_off) + g2.Kp, g2.Kp, 256, 256, big, 0))) return r;      // rows 1..256 (geo feature)
  if ((r = pack(L_B1, (const float*)(blob + g1.wt_off), g1.Np, 256, 256, big, 0))) return r;              // W1^T: [in][out]
  if ((r = pack(L_B0, (const float*)(blob + g0.wt_off), g0.Np, g0.K, 256, big, 0))) return r;             // W0^T: rows = input index
  // colour layer 0: input = [x(3) dir(27) grad(3) | geo feature(256) | appearance (+ n.v)]
  if ((r = pack(L_C0GF, (const float*)(blob + c0.w_off) + 33, c0.Kp, 256, 256, big, 0))) return r;
  if ((r = pack(L_C0MISC, (const float*)(blob + c0.w_off), c0.Kp, 256, c0.K - 256, 33, 256))) return r;
  if ((r = pack(L_C1, (const float*)(blob + c1.w_off), c1.Kp, 256, 256, big, 0))) return r;
  return 0;
}

int field_tc_forward(const sdfb200_field_t& f, const FieldPlan& p, const char* blob, const void* table, const sdfb200_field_in_t& in,
                     const sdfb200_field_out_t& out, float* ws, size_t ws_floats, cudaStream_t st) {
  const int64_t N = in.n_rays * (int64_t)in.n_samples;
  if (N == 0) return 0;
  TcPlan t;
  make_tc_plan(f, p, t);
  const size_t per_cta = kScratchPerCta(t.planes);
  if (ws_floats * sizeof(float) < (size_t)kNumSMs * per_cta) return fail(SDFB200_EWORKSPACE, "workspace too small for the tensor-core path%s", "", 0);
  const bool sdf_only = out.sdf && !out.geo_feature && !out.gradients && !out.normals && !out.rgb && !out.density && !out.alpha && !out.occupancy;
  if (!sdf_only) {
    if (out.rgb || out.alpha) SDFB_REQUIRE(in.directions != nullptr, "directions required for rgb / alpha");
    if (out.alpha) SDFB_REQUIRE(in.bins != nullptr && in.variance != nullptr, "alpha needs bins and the variance parameter");
    if (out.density) SDFB_REQUIRE(in.beta != nullptr && in.beta_min != nullptr, "density needs beta and beta_min");
  }
  SDFB_REQUIRE(out.sampled_sdf == nullptr, "sampled_sdf is only produced with use_numerical_gradients");
  TcArgs a;
  a.grid = f.grid;
  for (int l = 0; l < L_COUNT; ++l) a.layer[l] = t.layer[l];
  a.use_grid = f.use_grid_feature; a.pe_degree = f.pe_degree; a.use_pe = f.use_position_encoding;
  a.contraction = in.apply_contraction ? f.contraction : SDFB200_CONTRACT_NONE;
  a.in_dim = p.in_dim; a.pe_dim = p.pe_dim; a.grid_dim = p.grid_dim; a.cm_dim = t.cm_dim; a.app_dim = f.appearance_dim; a.use_n_dot_v = f.use_n_dot_v;
  a.mode = sdf_only ? 0 : 1;
  a.n_samples = in.n_samples; a.has_bins = in.bins != nullptr; a.n_points = N; a.n_tiles = (int)ceil_div(N, 128);
  a.rgb_padding = f.rgb_padding; a.cos_anneal = in.cos_anneal_ratio;
  a.origins = in.origins; a.directions = in.directions; a.bins = in.bins; a.appearance = in.appearance; a.variance = in.variance; a.beta = in.beta;
  a.beta_min = in.beta_min; a.table = table; a.blob = blob;
  a.b_g0 = p.geo[0].b_off; a.b_g1 = p.geo[1].b_off; a.b_g2 = p.geo[2].b_off; a.w_g2 = p.geo[2].w_off;
  a.b_c0 = p.col[0].b_off; a.b_c1 = p.col[1].b_off; a.w_c2 = p.col[2].w_off; a.b_c2 = p.col[2].b_off;
  a.scratch = reinterpret_cast<char*>(ws); a.scratch_per_cta = per_cta; a.out = out;
  const int grid = a.n_tiles < kNumSMs ? a.n_tiles : kNumSMs;
  const size_t smem = (size_t)t.planes * (kInK / 8) * 2048 + kMaxGridDim * 3 * 128 * 4 + (size_t)kStages * t.planes * 256 * kKB * 2 + (7 + 12) * 128 * 4 + 1024;
  if (t.planes == 2) {
    SDFB_CUDA(cudaFuncSetAttribute(k_field_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_field_tc<2><<<grid, kTcThreads, smem, st>>>(a);
  } else {
    SDFB_CUDA(cudaFuncSetAttribute(k_field_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_field_tc<1><<<grid, kTcThreads, smem, st>>>(a);
  }
  SDFB_LAUNCHED("k_field_tc");
  return 0;
}

}  // namespace sdfb200
