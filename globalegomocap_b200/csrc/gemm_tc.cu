// tcgen05 / TMEM / TMA GEMM for the large contractions (placeholder until the kernel lands).
#include "kernels.cuh"

namespace gem {
bool tc_gemm_available() { return false; }
int launch_tap_gemm_tc(cudaStream_t, const TapGemmArgs&, void*, size_t) {
    set_error("tcgen05 GEMM path not built");
    return GEM_ERR_STATE;
}
}  // namespace gem
