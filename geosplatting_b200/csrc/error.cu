// Thread-local error string + version for libgeosplat_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "gsb_common.cuh"

static thread_local char g_err[512] = "";

void gsb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" __attribute__((visibility("default"))) const char *gsb_last_error(void) { return g_err; }
extern "C" __attribute__((visibility("default"))) int gsb_version(void) { return 100; }
