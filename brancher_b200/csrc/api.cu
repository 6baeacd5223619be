// Error / variant reporting of the C ABI (include/brancher_cuda.h).
#include "common.cuh"

namespace brn {
static thread_local char g_err[512] = "";
static thread_local char g_variant[64] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void set_variant(const char* name) {
    strncpy(g_variant, name, sizeof(g_variant) - 1);
    g_variant[sizeof(g_variant) - 1] = 0;
}
}  // namespace brn

extern "C" int brn_abi_version(void) { return BRN_ABI_VERSION; }
extern "C" const char* brn_last_error(void) { return brn::g_err; }
extern "C" const char* brn_last_variant(void) { return brn::g_variant; }
