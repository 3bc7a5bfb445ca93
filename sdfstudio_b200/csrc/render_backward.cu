// Backward of the weights / compositing operators (training path).  The reference gets these from autograd over
// rays.py:131-230 and renderers.py:42-295; here they are two explicit kernels:
//   k_render_bwd   : d(per-ray outputs)/d(weights, rgb samples, normal samples)    -- elementwise given per-ray scalars
//   k_weights_bwd  : d(weights, last transmittance)/d(alphas | density)            -- one warp per ray, prefix scans in double
#include "common.cuh"

namespace sdfb200 {

struct RenderBwdArgs {
  const float *weights, *rgb, *normals, *eu, *bg;
  int bg_mode; int64_t R; int S;
  const float *acc, *depth;
  const float *g_rgb, *g_depth, *g_normal, *g_acc, *g_weights_in;
  float *g_weights, *g_rgb_s, *g_normal_s;
};

__global__ void __launch_bounds__(256) k_render_bwd(const RenderBwdArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R * a.S) return;
  const int64_t r = i / a.S;
  const int s = (int)(i - r * a.S);
  const int S = a.S;
  const float w = a.weights[i];
  float gw = a.g_weights_in ? a.g_weights_in[i] : 0.f;
  if (a.g_rgb) {
    const float g0 = a.g_rgb[r * 3], g1 = a.g_rgb[r * 3 + 1], g2 = a.g_rgb[r * 3 + 2];
    const float* c = a.rgb + i * 3;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (a.bg_mode == SDFB200_BG_COLOR) { b0 = a.bg[0]; b1 = a.bg[1]; b2 = a.bg[2]; }
    else if (a.bg_mode == SDFB200_BG_PER_RAY) { b0 = a.bg[r * 3]; b1 = a.bg[r * 3 + 1]; b2 = a.bg[r * 3 + 2]; }
    else { const float* cl = a.rgb + (r * S + S - 1) * 3; b0 = cl[0]; b1 = cl[1]; b2 = cl[2]; }
    gw += g0 * (c[0] - b0) + g1 * (c[1] - b1) + g2 * (c[2] - b2);   // out = sum w c + bg (1 - sum w)
    if (a.g_rgb_s) {
      float e0 = g0 * w, e1 = g1 * w, e2 = g2 * w;
      if (a.bg_mode == SDFB200_BG_LAST_SAMPLE && s == S - 1) {
        const float rem = 1.0f - a.acc[r];
        e0 += g0 * rem; e1 += g1 * rem; e2 += g2 * rem;
      }
      a.g_rgb_s[i * 3] = e0; a.g_rgb_s[i * 3 + 1] = e1; a.g_rgb_s[i * 3 + 2] = e2;
    }
  } else if (a.g_rgb_s) {
    a.g_rgb_s[i * 3] = 0.f; a.g_rgb_s[i * 3 + 1] = 0.f; a.g_rgb_s[i * 3 + 2] = 0.f;
  }
  if (a.g_normal) {
    const float g0 = a.g_normal[r * 3], g1 = a.g_normal[r * 3 + 1], g2 = a.g_normal[r * 3 + 2];
    const float* n = a.normals + i * 3;
    gw += g0 * n[0] + g1 * n[1] + g2 * n[2];
    if (a.g_normal_s) { a.g_normal_s[i * 3] = g0 * w; a.g_normal_s[i * 3 + 1] = g1 * w; a.g_normal_s[i * 3 + 2] = g2 * w; }
  } else if (a.g_normal_s) {
    a.g_normal_s[i * 3] = 0.f; a.g_normal_s[i * 3 + 1] = 0.f; a.g_normal_s[i * 3 + 2] = 0.f;
  }
  if (a.g_acc) gw += a.g_acc[r];
  if (a.g_depth) {
    // depth = sum(w step) / (acc + 1e-10)   (renderers.py:249-252, before the global clip)
    const float step = (a.eu[r * (S + 1) + s] + a.eu[r * (S + 1) + s + 1]) * 0.5f;
    gw += a.g_depth[r] * (step - a.depth[r]) / (a.acc[r] + 1e-10f);
  }
  a.g_weights[i] = gw;
}

// weights_i = alpha_i T_i,  T_i = prod_{j<i} f_j,  f_j = 1 - alpha_j + 1e-7 (rays.py:204-206)     [mode 0]
// weights_i = (1 - f_i) T_i, f_i = exp(-delta_i sigma_i), T_i = exp(-sum_{j<i} delta_j sigma_j) (rays.py:131-180)  [mode 1]
//   dL/dalpha_i = gw_i T_i - (sum_{k>i} (gw_k w_k + gT_k T_k) + gT_S T_S) / f_i ;      dL/dsigma_i = delta_i f_i dL/dalpha_i
// g_T: gradient of the returned transmittance, gt_cols = 0 (none), 1 (only the last column, [R]: bg_transmittance of
// models/neus.py:101) or the full width ([R,S+1] for alphas, [R,S] for densities: volsdf.py:67-68 reads transmittance[:, -1],
// the transmittance BEFORE the last sample).
__global__ void __launch_bounds__(256) k_weights_bwd(const float* __restrict__ in, const float* __restrict__ eu, int from_density, int64_t R, int S,
                                                     const float* __restrict__ g_weights, const float* __restrict__ g_T, int gt_cols,
                                                     float* __restrict__ g_in) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  double U = 0.0, TS = 1.0;
  for (int pass = 0; pass < 2; ++pass) {
    double carryT = 1.0, carryI = 0.0, carryU = 0.0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const bool on = s < S;
      float f = 1.f, al = 0.f, delta = 0.f, T;
      if (!from_density) {
        al = on ? in[r * S + s] : 0.f;
        f = on ? __fadd_rn(__fsub_rn(1.0f, al), 1e-7f) : 1.f;
        double incl = (double)f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl *= o;
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0;
        T = (float)(carryT * excl);
        carryT *= __shfl_sync(0xffffffffu, incl, 31);
      } else {
        delta = on ? __fsub_rn(eu[r * (S + 1) + s + 1], eu[r * (S + 1) + s]) : 0.f;
        const float dd = on ? __fmul_rn(delta, in[r * S + s]) : 0.f;
        double incl = (double)dd;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 0.0;
        T = expf(-(float)(carryI + excl));
        carryI += __shfl_sync(0xffffffffu, incl, 31);
        f = expf(-dd);
        al = 1.0f - f;
      }
      const float gw = on ? g_weights[r * S + s] : 0.f;
      const float gt = (on && (from_density ? gt_cols >= 1 : gt_cols > 1)) ? g_T[r * gt_cols + s] : 0.f;
      const double u = (double)gw * (double)(al * T) + (double)gt * (double)T;
      double uincl = u;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, uincl, d);
        if (lane >= d) uincl += o;
      }
      if (pass == 1 && on) {
        const double suffix = U - (carryU + uincl);                       // sum_{k>s} gw_k w_k
        const double tail = suffix + ((!from_density && gt_cols >= 1) ? (double)g_T[r * gt_cols + (gt_cols - 1)] * TS : 0.0);
        if (!from_density) g_in[r * S + s] = (float)((double)gw * (double)T - tail / (double)f);
        else g_in[r * S + s] = (float)((double)delta * ((double)gw * (double)T * (double)f - tail));
      }
      carryU += __shfl_sync(0xffffffffu, uincl, 31);
    }
    if (pass == 0) { U = carryU; TS = from_density ? 1.0 : carryT; }
  }
}

}  // namespace sdfb200

using namespace sdfb200;

extern "C" int sdfb200_render_backward(const float* weights, const float* rgb, const float* normals, const float* euclid_bins, const float* bg,
                                       int32_t bg_mode, int64_t n_rays, int32_t n_samples, const float* accumulation, const float* depth,
                                       const float* g_rgb, const float* g_depth, const float* g_normal, const float* g_accumulation,
                                       const float* g_weights_in, float* g_weights, float* g_rgb_samples, float* g_normal_samples, void* stream) {
  SDFB_REQUIRE(n_rays >= 0 && n_samples >= 1, "bad sizes");
  if (n_rays == 0) return 0;
  SDFB_REQUIRE(weights && g_weights, "NULL pointer");
  if (g_rgb) SDFB_REQUIRE(rgb != nullptr && (bg_mode == SDFB200_BG_LAST_SAMPLE || bg != nullptr) && accumulation != nullptr, "g_rgb needs rgb, background, accumulation");
  if (g_normal) SDFB_REQUIRE(normals != nullptr, "g_normal needs normals");
  if (g_depth) SDFB_REQUIRE(euclid_bins != nullptr && depth != nullptr && accumulation != nullptr, "g_depth needs bins, depth, accumulation");
  RenderBwdArgs a;
  a.weights = weights; a.rgb = rgb; a.normals = normals; a.eu = euclid_bins; a.bg = bg; a.bg_mode = bg_mode; a.R = n_rays; a.S = n_samples;
  a.acc = accumulation; a.depth = depth; a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_normal = g_normal; a.g_acc = g_accumulation;
  a.g_weights_in = g_weights_in; a.g_weights = g_weights; a.g_rgb_s = g_rgb_samples; a.g_normal_s = g_normal_samples;
  k_render_bwd<<<(unsigned)ceil_div(n_rays * n_samples, 256), 256, 0, (cudaStream_t)stream>>>(a);
  SDFB_LAUNCHED("k_render_bwd");
  return 0;
}

extern "C" int sdfb200_weights_backward(const float* alphas_or_density, const float* euclid_bins, int32_t from_density, int64_t n_rays,
                                        int32_t n_samples, const float* g_weights, const float* g_transmittance, int32_t g_transmittance_cols,
                                        float* g_in, void* stream) {
  SDFB_REQUIRE(n_rays >= 0 && n_samples >= 1, "bad sizes");
  if (n_rays == 0) return 0;
  SDFB_REQUIRE(alphas_or_density && g_weights && g_in, "NULL pointer");
  if (from_density) SDFB_REQUIRE(euclid_bins != nullptr, "density mode needs bins");
  const int gt_cols = g_transmittance ? g_transmittance_cols : 0;
  SDFB_REQUIRE(gt_cols == 0 || (gt_cols == 1 && !from_density) || gt_cols == n_samples + (from_density ? 0 : 1),
               "g_transmittance_cols must be 1 (alphas: last column only) or the width of the transmittance output");
  k_weights_bwd<<<(unsigned)ceil_div(n_rays, 8), 256, 0, (cudaStream_t)stream>>>(alphas_or_density, euclid_bins, from_density, n_rays, n_samples,
                                                                                 g_weights, g_transmittance, gt_cols, g_in);
  SDFB_LAUNCHED("k_weights_bwd");
  return 0;
}
