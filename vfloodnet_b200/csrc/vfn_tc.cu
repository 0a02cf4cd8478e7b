// placeholder until the tcgen05 kernels land
#include "vfn_tc.cuh"
namespace vfn {
bool tc_shapes_ok(int, int) { return false; }
void tc_pick_splits(int, int64_t, int64_t, int* a, int* b) { *a = 1; *b = 1; }
size_t tc_workspace_bytes(int, int64_t) { return 0; }
int tc_phase_a(const vfn_bank*, int, const float*, int64_t, int, float2*, char*, cudaStream_t) { return VFN_E_UNSUPPORTED; }
int tc_phase_b(const vfn_bank*, int, int64_t, int, const float*, float, int, float*, char*, cudaStream_t) { return VFN_E_UNSUPPORTED; }
}
extern "C" int vfn_debug_umma_ss(const uint16_t*, const uint16_t*, float*, int32_t, void*) { return VFN_E_UNSUPPORTED; }
