// Shared host/device helpers of libampnet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/ampnet_b200.h"

namespace amp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// thread-local last-error string (the only mutable state besides the launch counter)
char* last_error_buf();
int fail(int code, const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// after a launch: turn a launch error into AMP_E_CUDA (never synchronises)
inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(AMP_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return AMP_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// debugging aid: AMP_DISABLE=name1,name2 switches optional fast paths off (they fall back to the generic kernels)
bool path_disabled(const char* name);

}  // namespace amp
