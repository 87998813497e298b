// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace sdb {
bool tc_supported(const Geo&, const char** why) { *why = "tensor-core path not built yet"; return false; }
size_t tc_workspace_bytes(int, const Geo&, int) { return 0; }
size_t tc_packed_input_bytes(const Geo&) { return 0; }
int tc_forward(const void*, const float*, const float*, const void*, const void*, void*, const Geo&, int, void*, size_t, void*, cudaStream_t) { return SDB_ERR_UNSUPPORTED; }
int tc_backward_data(const void*, const float*, const float*, const void*, const void*, void*, float*, float*, const Geo&, int, void*, size_t, const void*, cudaStream_t) { return SDB_ERR_UNSUPPORTED; }
int tc_backward_weight(const void*, const float*, const float*, const void*, float*, float*, float, const Geo&, int, void*, size_t, const void*, cudaStream_t) { return SDB_ERR_UNSUPPORTED; }
}
