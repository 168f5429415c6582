// Error plumbing and version of the C ABI (include/dropdec.h).
#include <cstdlib>
#include <string>

#include "dd_common.cuh"

namespace dd {
static thread_local std::string g_last_error;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
bool pdl_enabled() {
  static const bool on = [] {
    // Off by default: inside a CUDA graph the kernel-to-kernel gap is already ~1 us and programmatic
    // launch measured 1-5 % SLOWER on the 7B decode step (profiles/README.md); DD_PDL=1 turns it on.
    const char* e = getenv("DD_PDL");
    return e && e[0] == '1';
  }();
  return on;
}
}  // namespace dd

extern "C" int dd_version(void) { return 100; }  // 0.1.0
extern "C" const char* dd_last_error(void) { return dd::g_last_error.c_str(); }
