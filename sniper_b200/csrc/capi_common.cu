// Error reporting for the C-ABI (mirror of MXGetLastError, SNIPER-mxnet/include/mxnet/c_api.h:196-204).
#include "common.cuh"
#include <stdarg.h>

namespace sn {
char* last_error_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 1024, fmt, ap);
  va_end(ap);
}
}  // namespace sn

extern "C" {
const char* sniper_last_error(void) { return sn::last_error_buf(); }
int sniper_abi_version(void) { return 2; }
}
