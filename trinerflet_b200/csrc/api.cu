// ABI bookkeeping: version and thread-local error text.
#include "common.cuh"
#include <stdarg.h>

namespace tnl {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace tnl

extern "C" {
int tnl_abi_version(void) { return TNL_ABI_VERSION; }
const char* tnl_last_error(void) { return tnl::g_err; }
}
