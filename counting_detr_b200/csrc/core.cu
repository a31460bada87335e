// Library-wide plumbing of the C ABI: version and the per-thread error string.
#include "common.cuh"
#include "../../include/cdetr.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void cdetr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cdetr_last_error(void) { return g_err; }
extern "C" int cdetr_version(void) { return 1; }
