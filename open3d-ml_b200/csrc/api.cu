// api.cu -- ABI version and the thread-local error message of libo3dml_b200.
#include "../../include/o3dml_b200.h"
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

extern "C" void o3dml_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* o3dml_last_error(void) { return g_err; }
extern "C" int o3dml_abi_version(void) { return O3DML_ABI_VERSION; }
