// Shared plumbing of the C-ABI translation units (error reporting, CUDA call checking).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <string>

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
#define LAUNCH_CHECK(name, stream)                                                                      \
    do {                                                                                                \
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
