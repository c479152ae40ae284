// Error reporting and version for libnuhtc_b200.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void nuhtc_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

NUHTC_API int nuhtc_abi_version(void) { return 1; }
NUHTC_API const char *nuhtc_last_error(void) { return g_err; }
