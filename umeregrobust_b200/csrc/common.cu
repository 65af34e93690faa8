// Error plumbing, launch counter and ABI version for libumereg_b200.
#include "ume_common.cuh"

#include <atomic>
#include <stdarg.h>

namespace ume {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return UME_ERR_CUDA;
    }
    return UME_OK;
}

const char* last_error() { return g_err; }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace ume

extern "C" int ume_abi_version(void) { return UME_ABI_VERSION; }
extern "C" const char* ume_last_error(void) { return ume::last_error(); }
extern "C" uint64_t ume_launch_count(void) { return ume::launches(); }
extern "C" const char* ume_status_string(int status) {
    switch (status) {
        case UME_OK: return "ok";
        case UME_ERR_BAD_ARG: return "bad argument";
        case UME_ERR_WORKSPACE: return "workspace missing or too small";
        case UME_ERR_UNSUPPORTED: return "size not supported by this build";
        case UME_ERR_CUDA: return "CUDA error";
        default: return "unknown status";
    }
}
