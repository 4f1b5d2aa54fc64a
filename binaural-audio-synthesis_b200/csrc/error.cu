// Per-thread error string and the version / device queries every library of this package exports.
#include <stdarg.h>

#include "bas_internal.cuh"

static thread_local char g_err[512] = "";

void bas_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int bas_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return BAS_E_ARG;
    strncpy(buf, g_err, len - 1);
    buf[len - 1] = 0;
    return 0;
}

extern "C" int bas_abi_version(void) { return BAS_ABI_VERSION; }

extern "C" int bas_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

bool& bas_pdl_flag() {
    static thread_local bool on = false;
    return on;
}
