// tcgen05 tensor-core implementation of the subspace-distance GEMM (impl 1) — placeholder until
// the UMMA kernel lands; reports "unsupported" rather than silently using another path.
#include "ume_common.cuh"

namespace ume {

size_t cdist_tc_workspace_bytes(int, int, int, int) { return 0; }

int cdist_tc_launch(const float*, const float*, int, int, int, int, float*, int64_t*, float*, void*, size_t, cudaStream_t) {
    set_error("ume_cdist_f32: impl 1 (tcgen05) is not built into this library yet");
    return UME_ERR_UNSUPPORTED;
}

}  // namespace ume
