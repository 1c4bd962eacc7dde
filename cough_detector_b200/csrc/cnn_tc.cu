// CoughDetectorResidual forward on the tcgen05 tensor cores (CD_PRECISION_BF16) -- placeholder
// until the implicit-GEMM kernels land; reports CD_E_STATE so nothing silently falls back.
#include <vector>

#include "cd_common.cuh"

namespace cd {

int init_cnn_tc(cd_handle*) { return CD_OK; }
int upload_cnn_tc(cd_handle*, const std::vector<std::vector<float>>&, const std::vector<std::vector<float>>&) {
    return CD_OK;
}
size_t cnn_tc_ws_bytes(int64_t) { return 0; }
int launch_cnn_tc(cd_handle* h, const float*, int64_t, float*, cudaStream_t) {
    return set_err(h, CD_E_STATE, "CD_PRECISION_BF16 (tcgen05 path) is not built in this version");
}

}  // namespace cd
