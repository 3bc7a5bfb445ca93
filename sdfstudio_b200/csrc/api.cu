// extern "C" entry points of the field (pack / forward) + library bookkeeping.  See include/sdfb200.h.
#include "common.cuh"
#include "field_plan.h"

namespace sdfb200 {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int validate_grid(const sdfb200_grid_t* g);
int field_pack_fp32(const sdfb200_field_t& f, const FieldPlan& p, const sdfb200_field_params_t& prm, char* blob, cudaStream_t st);
int field_forward_fp32(const sdfb200_field_t& f, const FieldPlan& p, const char* blob, const void* table, const sdfb200_field_in_t& in,
                       const sdfb200_field_out_t& out, float* ws, size_t ws_floats, cudaStream_t st, int gemm_planes);
// tensor-core path (field_tc.cu)
size_t field_tc_packed_bytes(const sdfb200_field_t& f, const FieldPlan& p);
bool field_tc_supported(const sdfb200_field_t& f, const FieldPlan& p);
int field_tc_pack(const sdfb200_field_t& f, const FieldPlan& p, char* blob, cudaStream_t st);
int field_tc_forward(const sdfb200_field_t& f, const FieldPlan& p, const char* blob, const void* table, const sdfb200_field_in_t& in,
                     const sdfb200_field_out_t& out, float* ws, size_t ws_floats, cudaStream_t st);
size_t field_tc_workspace_floats(const sdfb200_field_t& f, const FieldPlan& p, int64_t n_points);

static int plan_or_fail(const sdfb200_field_t* f, FieldPlan& p) {
  SDFB_REQUIRE(f != nullptr, "field descriptor is NULL");
  if (f->use_grid_feature) {
    int r = validate_grid(&f->grid);
    if (r) return r;
  } else {
    SDFB_REQUIRE(f->grid.n_levels >= 1 && f->grid.n_levels <= SDFB200_MAX_LEVELS && f->grid.n_features >= 1, "grid dims (needed for the zero feature block)");
  }
  SDFB_REQUIRE(f->pe_degree >= 0 && f->pe_degree <= 16, "pe_degree out of range");
  SDFB_REQUIRE(f->appearance_dim >= 0 && f->appearance_dim <= 256, "appearance_dim out of range");
  const int rc = make_field_plan(*f, p);
  if (rc) return fail(SDFB200_EINVAL, "inconsistent field descriptor%s (plan error %lld)", "", (long long)rc);
  // tensor-core precisions: the fused kernel (field_tc.cu) for the neus-facto shape family, otherwise the generic kernels with the
  // tcgen05 Linear (tc_linear.cu) as their GEMM engine -- no extra packed section for the latter
  if (f->precision != SDFB200_PRECISION_FP32 && field_tc_supported(*f, p)) {
    p.tc_bytes = field_tc_packed_bytes(*f, p);
    p.total_bytes = p.tc_off + p.tc_bytes;
  }
  return 0;
}
}  // namespace sdfb200

using namespace sdfb200;

extern "C" int sdfb200_version(void) { return SDFB200_VERSION; }
extern "C" const char* sdfb200_last_error_string(void) { return g_err; }
extern "C" int64_t sdfb200_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" size_t sdfb200_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(sdfb200_grid_t);
    case 1: return sizeof(sdfb200_field_t);
    case 2: return sizeof(sdfb200_field_params_t);
    case 3: return sizeof(sdfb200_field_in_t);
    case 4: return sizeof(sdfb200_field_out_t);
    case 5: return sizeof(sdfb200_render_out_t);
    default: return 0;
  }
}

extern "C" size_t sdfb200_field_packed_bytes(const sdfb200_field_t* f) {
  FieldPlan p;
  if (plan_or_fail(f, p)) return 0;
  return p.total_bytes;
}

extern "C" int sdfb200_field_pack(const sdfb200_field_t* f, const sdfb200_field_params_t* prm, void* packed, void* stream) {
  FieldPlan p;
  int r = plan_or_fail(f, p);
  if (r) return r;
  SDFB_REQUIRE(prm != nullptr && packed != nullptr, "NULL pointer");
  r = field_pack_fp32(*f, p, *prm, (char*)packed, (cudaStream_t)stream);
  if (r) return r;
  if (f->precision != SDFB200_PRECISION_FP32 && p.tc_bytes > 0) return field_tc_pack(*f, p, (char*)packed, (cudaStream_t)stream);
  return 0;
}

extern "C" size_t sdfb200_field_workspace_bytes(const sdfb200_field_t* f, int64_t n_points) {
  FieldPlan p;
  if (plan_or_fail(f, p) || n_points < 0) return 0;
  if (n_points == 0) return 256;
  FieldWorkspace w;
  make_workspace_plan(*f, p, n_points < kChunkPoints ? n_points : kChunkPoints, w);
  size_t floats = w.floats_per_chunk;
  if (f->precision != SDFB200_PRECISION_FP32 && p.tc_bytes > 0) {
    const size_t t = field_tc_workspace_floats(*f, p, n_points);
    floats = t > floats ? t : floats;
  }
  return floats * sizeof(float) + 256;
}

extern "C" int sdfb200_field_forward(const sdfb200_field_t* f, const void* packed, const void* table, const sdfb200_field_in_t* in,
                                     const sdfb200_field_out_t* out, void* workspace, size_t workspace_bytes, void* stream) {
  FieldPlan p;
  int r = plan_or_fail(f, p);
  if (r) return r;
  SDFB_REQUIRE(packed && in && out, "NULL pointer");
  SDFB_REQUIRE(in->n_rays >= 0 && in->n_samples >= 1, "bad sizes");
  if (in->n_rays == 0) return 0;
  SDFB_REQUIRE(in->origins != nullptr, "origins is NULL");
  SDFB_REQUIRE(in->bins != nullptr || in->n_samples == 1, "point mode requires n_samples == 1");
  SDFB_REQUIRE(in->bins == nullptr || in->directions != nullptr, "ray mode requires directions");
  SDFB_REQUIRE(!f->use_grid_feature || table != nullptr, "grid table is NULL");
  SDFB_REQUIRE(workspace != nullptr, "workspace is NULL");
  uintptr_t wsp = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
  const size_t lost = wsp - (uintptr_t)workspace;
  SDFB_REQUIRE(workspace_bytes > lost, "workspace too small");
  const size_t ws_floats = (workspace_bytes - lost) / sizeof(float);
  // the tensor-core kernel never materialises the geo feature (colour layer 0 is pre-multiplied with the last geo layer);
  // the rare callers that want it (forward_geonetwork) take the exact-fp32 kernels, which read the same packed blob
  const bool fused = p.tc_bytes > 0;
  if (f->precision != SDFB200_PRECISION_FP32 && fused && out->geo_feature == nullptr)
    return field_tc_forward(*f, p, (const char*)packed, table, *in, *out, (float*)wsp, ws_floats, (cudaStream_t)stream);
  // shapes outside the fused family run the generic kernels with tensor-core GEMMs; the fused family's geo-feature requests keep
  // the exact-fp32 engine (unchanged behaviour)
  // numerical gradients divide sdf differences by 2 delta (~1e-3): they need the sdf to fp32 accuracy, which 2^-16-relative GEMMs do not
  // give (measured: gradient error 6e-3 vs 5e-4 of fp32 noise on the angelo-shaped case) -> those fields keep the exact engine
  const int gemm_planes = (f->precision != SDFB200_PRECISION_FP32 && !fused && !f->use_numerical_gradients)
                              ? (f->precision == SDFB200_PRECISION_BF16 ? 1 : 2) : 0;
  return field_forward_fp32(*f, p, (const char*)packed, table, *in, *out, (float*)wsp, ws_floats, (cudaStream_t)stream, gemm_planes);
}

// building-block test of the generic tcgen05 Linear (tc_linear.cu) against a reference GEMM: same arguments as the internal sgemm()
namespace sdfb200 {
int tc_gemm(int planes, int epi, const float* X, int ldx, const float* W, const float* bias, float* Y, int ldy, int64_t M, int Np, int Kp,
            const float* aux, int ldaux, int aux_cols, void* scratch, cudaStream_t st);
}
extern "C" int sdfb200_debug_tc_linear(int32_t planes, int32_t epi, const float* X, int32_t ldx, const float* W, const float* bias, float* Y,
                                       int32_t ldy, int64_t M, int32_t Np, int32_t Kp, const float* aux, int32_t ldaux, int32_t aux_cols,
                                       void* scratch, void* stream) {
  SDFB_REQUIRE(X && W && Y && scratch && M >= 0, "NULL pointer");
  return tc_gemm(planes, epi, X, ldx, W, bias, Y, ldy, M, Np, Kp, aux, ldaux, aux_cols, scratch, (cudaStream_t)stream);
}
