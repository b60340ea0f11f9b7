// tcgen05 3xTF32 GEMM (placeholder until the kernel lands: reports "unsupported" so the
// dispatcher uses the SIMT path; FI_GEMM_TCGEN05 then fails loudly).
#include "fi_internal.cuh"

namespace fi {
bool gemm_tc_supported(int, int, int, int, const float*, int, const float*, int, const float*, int) { return false; }
size_t gemm_tc_workspace_bytes(int, int, int, int) { return 0; }
int launch_gemm_tc(int, int, int, int, const float*, int, const float*, int, float*, int, const float*, int,
                   const float*, int, void*, size_t, cudaStream_t) {
    return set_error(FI_ERR_STATE, "tcgen05 GEMM not available in this build");
}
}  // namespace fi
