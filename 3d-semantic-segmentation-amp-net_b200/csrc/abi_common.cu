// Error state, version and launch counter of the C ABI (include/ampnet_b200.h).
#include <stdlib.h>
#include <string.h>

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

bool path_disabled(const char* name) {
    static const char* env = getenv("AMP_DISABLE");
    return env && strstr(env, name) != nullptr;
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
}
