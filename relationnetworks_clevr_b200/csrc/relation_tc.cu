// placeholder until the tcgen05 kernels land
#include "relation.cuh"
namespace rn {
bool tc_supported(const RelShape&) { return false; }
size_t tc_saved_bytes(const RelShape&, bool) { return 0; }
size_t tc_scratch_bytes(const RelShape&, bool) { return 0; }
int tc_relation_fwd(const RelShape&, int, bool, const float*, const float*, const float* const*, const float* const*, float*, void*, void*, cudaStream_t) { return fail(RN_ERR_UNSUPPORTED, "tc path not built"); }
int tc_relation_bwd(const RelShape&, int, const float*, const float*, const float*, const float* const*, const void*, float*, float*, float* const*, float* const*, void*, cudaStream_t) { return fail(RN_ERR_UNSUPPORTED, "tc path not built"); }
}
