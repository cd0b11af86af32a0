#include "common.h"

namespace tfsc {

std::string& last_error_ref() {
  static thread_local std::string e;
  return e;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

}  // namespace tfsc

extern "C" {
int tfsc_abi_version(void) { return TFSC_ABI_VERSION; }
const char* tfsc_last_error(void) { return tfsc::last_error_ref().c_str(); }
const char* tfsc_strerror(int code) {
  switch (code) {
    case TFSC_OK: return "OK";
    case TFSC_E_INVALID: return "INVALID_ARGUMENT";
    case TFSC_E_TIMEOUT: return "DEADLINE_EXCEEDED";
    case TFSC_E_NOT_FOUND: return "NOT_FOUND";
    case TFSC_E_EXHAUSTED: return "RESOURCE_EXHAUSTED";
    case TFSC_E_UNIMPLEMENTED: return "UNIMPLEMENTED";
    case TFSC_E_INTERNAL: return "INTERNAL";
    case TFSC_E_NO_DEVICE: return "UNAVAILABLE";
    case TFSC_E_EMPTY_RING: return "EMPTY_RING";
    case TFSC_E_BUFFER: return "BUFFER_TOO_SMALL";
    default: return "UNKNOWN";
  }
}
void tfsc_free(void* p) { free(p); }
}
