// Shared helpers for the freeimpala-b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/fi_learner.h"

namespace fi {

// ---- thread-local error message (fi_last_error) -------------------------------------------
inline std::string& last_error_ref() {
    static thread_local std::string msg;
    return msg;
}
inline int set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    fprintf(stderr, "[freeimpala_b200][error] %s\n", buf);  // reference style: log and continue
    return code;
}

#define FI_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::fi::set_error(FI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                                   cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

#define FI_CUDA_OK_NULL(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::fi::set_error(FI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
            return nullptr;                                                                    \
        }                                                                                      \
    } while (0)

#define FI_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s < 0) return _s;        \
    } while (0)

// ---- kernel launch accounting (fi_kernel_launch_count) ------------------------------------
inline std::atomic<uint64_t>& launch_counter() {
    static std::atomic<uint64_t> n{0};
    return n;
}
inline void count_launch(uint64_t k = 1) { launch_counter().fetch_add(k, std::memory_order_relaxed); }

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(FI_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    count_launch();
    return FI_OK;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- record layout of one 1024-byte trajectory element (DESIGN.md "Record layout") --------
constexpr int kRecWords = 256;
constexpr int kZDim = 162;
constexpr int kXDim = 484;
constexpr int kNumActions = 16;
constexpr int kWMu = 162;
constexpr int kWAction = 178;
constexpr int kWReward = 179;
constexpr int kWDiscount = 180;
constexpr int kWAux = 181;
constexpr int kWX = 192;
constexpr int kXPerRec = 64;

}  // namespace fi
