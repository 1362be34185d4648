// Shared plumbing of the C-ABI translation units (error reporting, CUDA call checking).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/dspfe.h"

namespace dspfe {
inline thread_local std::string g_err;
inline int fail(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace dspfe

// After a kernel launch: always check the launch status; with DSPFE_DEBUG_SYNC=1 in the environment also wait for
// the kernel and name it if it faulted (debug aid, off by default: the entry points never synchronise).
namespace dspfe {
inline bool debug_sync() { static const bool v = [] { const char* e = getenv("DSPFE_DEBUG_SYNC"); return e && e[0] == '1'; }(); return v; }
}
// Optional per-kernel timing (dspfe_timing_begin / dspfe_timing_end, include/dspfe.h): while a collection is open on a
// stream, every kernel the library launches on that stream is followed by an event; a kernel's time is the distance
// between its event and the previous one (launches on one stream are serial).  Off by default: one predictable branch.
namespace dspfe {
struct StageTimer {
    bool on = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t start = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    long long launches = 0;          // every LAUNCH_CHECK since the library was loaded (for reports)
};
inline StageTimer g_timer;
inline void timer_mark(const char* name, cudaStream_t st) {
    ++g_timer.launches;
    if (!g_timer.on || st != g_timer.stream) return;
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    g_timer.marks.emplace_back(name, e);
}
}  // namespace dspfe
#define LAUNCH_CHECK(name, stream)                                                                      \
    do {                                                                                                \
        ::dspfe::timer_mark(name, stream);                                                              \
        cudaError_t e_ = cudaGetLastError();                                                            \
        if (e_ == cudaSuccess && ::dspfe::debug_sync()) e_ = cudaStreamSynchronize(stream);             \
        if (e_ != cudaSuccess)                                                                          \
            return ::dspfe::fail(DSPFE_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(e_));    \
    } while (0)

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return ::dspfe::fail(DSPFE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));   \
    } while (0)
