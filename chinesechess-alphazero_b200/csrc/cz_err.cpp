// cz_err.cpp — error string storage and the two build-independent C-ABI entry points.
#include "../../include/cczero_b200.h"
#include "cz_err.h"
static thread_local char g_err[512] = "";
const char* cz_err_get() { return g_err; }
int cz_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char* cz_last_error(void) { return g_err; }
extern "C" int cz_build_is_cuda(void) {
#if defined(CZ_EMUL)
  return 0;
#else
  return 1;
#endif
}
