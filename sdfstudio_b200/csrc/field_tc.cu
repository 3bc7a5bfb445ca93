// tcgen05 (5th-gen tensor core) evaluation of the SDF field -- placeholder until the fused kernel lands.
#include "common.cuh"
#include "field_plan.h"

namespace sdfb200 {
bool field_tc_supported(const sdfb200_field_t&, const FieldPlan&) { return false; }
size_t field_tc_packed_bytes(const sdfb200_field_t&, const FieldPlan&) { return 0; }
size_t field_tc_workspace_floats(const sdfb200_field_t&, const FieldPlan&, int64_t) { return 0; }
int field_tc_pack(const sdfb200_field_t&, const FieldPlan&, char*, cudaStream_t) { return fail(SDFB200_EUNSUPPORTED, "tensor-core path not built%s", "", 0); }
int field_tc_forward(const sdfb200_field_t&, const FieldPlan&, const char*, const void*, const sdfb200_field_in_t&, const sdfb200_field_out_t&, float*,
                     size_t, cudaStream_t) {
  return fail(SDFB200_EUNSUPPORTED, "tensor-core path not built%s", "", 0);
}
}  // namespace sdfb200
