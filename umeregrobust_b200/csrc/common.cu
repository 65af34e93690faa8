// Error plumbing, launch counter and ABI version for libumereg_b200.
#include "ume_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>
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

// ---------------------------------------------------------------- stage profiler
// Optional CUDA-event bracket around each stage's kernel launches, on the launching stream.
// Off by default; bench.py switches it on for the timed region to get the per-kernel durations
// the roofline figures are computed from.
namespace {
struct ProfSlot {
    std::vector<cudaEvent_t> ev;     // pairs (start, stop)
    size_t used = 0;                 // events handed out and not yet folded into total
    double total_ms = 0.0;
    uint64_t n = 0;
};
std::mutex g_prof_mu;
std::atomic<int> g_prof_on{0};
ProfSlot g_prof[UME_PROF_SLOTS];

void prof_fold(ProfSlot& s) {        // caller holds the mutex
    for (size_t i = 0; i + 1 < s.used; i += 2) {
        float ms = 0.f;
        if (cudaEventSynchronize(s.ev[i + 1]) == cudaSuccess && cudaEventElapsedTime(&ms, s.ev[i], s.ev[i + 1]) == cudaSuccess) {
            s.total_ms += ms;
            s.n += 1;
        }
    }
    s.used = 0;
}
}  // namespace

int prof_begin(int slot, cudaStream_t stream) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return -1;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfSlot& s = g_prof[slot];
    if (s.used + 2 > 8192) prof_fold(s);
    while (s.ev.size() < s.used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        s.ev.push_back(e);
    }
    const int at = (int)s.used;
    s.used += 2;
    cudaEventRecord(s.ev[at], stream);
    return at;
}

void prof_end(int slot, int token, cudaStream_t stream) {
    if (token < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEventRecord(g_prof[slot].ev[token + 1], stream);
}
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

}  // namespace ume

extern "C" void ume_profile_enable(int on) { ume::g_prof_on.store(on ? 1 : 0); }
extern "C" void ume_profile_reset(void) {
    std::lock_guard<std::mutex> lk(ume::g_prof_mu);
    for (auto& s : ume::g_prof) { ume::prof_fold(s); s.total_ms = 0.0; s.n = 0; }
}
extern "C" int ume_profile_read(int slot, double* total_ms, uint64_t* launches) {
    if (slot < 0 || slot >= UME_PROF_SLOTS) return UME_ERR_BAD_ARG;
    std::lock_guard<std::mutex> lk(ume::g_prof_mu);
    ume::prof_fold(ume::g_prof[slot]);
    if (total_ms) *total_ms = ume::g_prof[slot].total_ms;
    if (launches) *launches = ume::g_prof[slot].n;
    return UME_OK;
}

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
