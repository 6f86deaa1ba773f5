// C-ABI housekeeping: version + thread-local last-error string.
#include <stdarg.h>

#include "common.cuh"

namespace pita {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pita

extern "C" int pita_abi_version(void) { return PITA_ABI_VERSION; }
extern "C" const char *pita_last_error(void) { return pita::g_err; }
