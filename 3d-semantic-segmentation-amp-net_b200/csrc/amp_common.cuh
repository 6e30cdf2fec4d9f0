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

// Programmatic dependent launch. Every kernel of the network path is launched with programmatic stream serialization:
// the grid may start while its predecessor in the stream is still draining, runs whatever touches no global memory
// (shared-memory carve-up, mbarrier init, TMEM allocation), and then blocks in pdl_wait() until the predecessor has
// completed and its writes are visible. Each kernel also releases ITS successor right away (pdl_trigger), so launch
// latency and CTA start-up of a chain of short dependent kernels overlap with the previous kernel instead of adding up.
// Rule: no global-memory access before pdl_wait(). Kernels launched without the attribute (torch's own, or AMP_DISABLE=pdl)
// see both instructions as no-ops and serialise as usual. Measured on B200 (batch 32 x 2048): the eval forward, replayed
// as a CUDA graph, gains 2 %; the eager training step loses 1.5 %, so only the eval entry points open a PdlScope.
bool pdl_enabled();
struct PdlScope {                       // RAII: launches of this thread use the attribute while an enabled scope is alive
    explicit PdlScope(bool on);
    ~PdlScope();
    bool prev;
};
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }
#endif

// debugging aid: AMP_DISABLE=name1,name2 switches optional fast paths off (they fall back to the generic kernels)
bool path_disabled(const char* name);
struct StrictScope {                    // RAII: calls of this thread avoid the split-bf16 tensor-core kernels while alive
    explicit StrictScope(bool on);
    ~StrictScope();
    bool prev;
};
// device word added to every dropout seed (amp_set_dropout_offset; null = none)
const unsigned long long* dropout_offset();
// per-path launch counter behind amp_path_count() (tests assert which kernel family served a call)
void count_path(const char* name, int n = 1);

}  // namespace amp
