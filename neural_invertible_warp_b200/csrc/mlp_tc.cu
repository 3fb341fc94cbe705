// placeholder until the tcgen05 path lands
#include "common.cuh"
#include "mlp_shared.cuh"
namespace niw {
size_t tc_workspace_bytes(int64_t, int, int) { return 0; }
int tc_fwd(const float*, const float*, const float*, const float*, int64_t, int, const Bands3&, const BandsV&, int, void*, size_t, float*, float*, cudaStream_t) { return NIW_E_UNSUPP; }
int tc_bwd(const float*, const float*, const float*, const float*, int64_t, int, const Bands3&, const BandsV&, void*, size_t, const float*, const float*, float*, float*, float*, cudaStream_t) { return NIW_E_UNSUPP; }
}
extern "C" int niw_tc_selftest(const float*, const float*, int, int, int, float*, void*) { return NIW_E_UNSUPP; }
