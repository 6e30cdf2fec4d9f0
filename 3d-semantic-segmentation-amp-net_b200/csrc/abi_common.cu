// Error state, version and launch counter of the C ABI (include/ampnet_b200.h).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "amp_common.cuh"

namespace amp {
std::atomic<long long> g_launches{0};

char* last_error_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

static thread_local bool t_pdl_scope = false;
bool pdl_enabled() {
    static const bool allowed = !path_disabled("pdl");
    return allowed && t_pdl_scope;
}
PdlScope::PdlScope(bool on) : prev(t_pdl_scope) { t_pdl_scope = on; }
PdlScope::~PdlScope() { t_pdl_scope = prev; }

// AMP_DISABLE from the environment, or the list set at run time by amp_debug_set_disabled() (tests switch the tensor-core
// paths off and on inside one process to compare them with the CUDA-core kernels on the same inputs)
static std::mutex g_dbg_mu;
static char g_disabled[256];
static bool g_disabled_set = false;
static thread_local bool t_strict = false;
StrictScope::StrictScope(bool on) : prev(t_strict) { t_strict = t_strict || on; }
StrictScope::~StrictScope() { t_strict = prev; }
bool path_disabled(const char* name) {
    static const char* env = getenv("AMP_DISABLE");
    // AMP_PREC_FP32_STRICT: no split-bf16 tensor-core arithmetic anywhere in the call (plain fp32 FMA kernels serve it)
    if (t_strict && (!strcmp(name, "tc_layer") || !strcmp(name, "tc_wgrad") || !strcmp(name, "tc_chain32"))) return true;
    std::lock_guard<std::mutex> lk(g_dbg_mu);
    const char* list = g_disabled_set ? g_disabled : env;
    return list && strstr(list, name) != nullptr;
}

// process-wide, not thread-local: autograd runs the backward (amp_seg_bwd) on its own worker thread
static std::atomic<const unsigned long long*> g_drop_off{nullptr};
const unsigned long long* dropout_offset() { return g_drop_off.load(std::memory_order_acquire); }
void set_dropout_offset(const unsigned long long* p) { g_drop_off.store(p, std::memory_order_release); }

// per-path launch counters (amp_path_count): which kernel family actually served a call
struct PathCounter { char name[32]; long long n; };
static PathCounter g_paths[48];
static int g_n_paths = 0;
void count_path(const char* name, int n) {
    std::lock_guard<std::mutex> lk(g_dbg_mu);
    for (int i = 0; i < g_n_paths; ++i)
        if (strcmp(g_paths[i].name, name) == 0) { g_paths[i].n += n; return; }
    if (g_n_paths < 48) {
        strncpy(g_paths[g_n_paths].name, name, 31);
        g_paths[g_n_paths].name[31] = 0;
        g_paths[g_n_paths++].n = n;
    }
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace amp

extern "C" {
const char* amp_last_error(void) { return amp::last_error_buf(); }
int amp_abi_version(void) { return 1000; }
int64_t amp_launch_count(void) { return (int64_t)amp::g_launches.load(); }
int64_t amp_path_count(const char* name) {
    if (!name) return -1;
    std::lock_guard<std::mutex> lk(amp::g_dbg_mu);
    for (int i = 0; i < amp::g_n_paths; ++i)
        if (strcmp(amp::g_paths[i].name, name) == 0) return (int64_t)amp::g_paths[i].n;
    return 0;
}
int amp_set_dropout_offset(const void* device_u64) {
    amp::set_dropout_offset(reinterpret_cast<const unsigned long long*>(device_u64));
    return AMP_OK;
}
int amp_debug_set_disabled(const char* csv) {
    std::lock_guard<std::mutex> lk(amp::g_dbg_mu);
    if (!csv) { amp::g_disabled_set = false; return AMP_OK; }
    if (strlen(csv) >= sizeof amp::g_disabled) return AMP_E_BADARG;
    strcpy(amp::g_disabled, csv);
    amp::g_disabled_set = true;
    return AMP_OK;
}
}
