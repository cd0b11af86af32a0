// Shared helpers for libtfsc_b200: thread-local error text, buffer copy-out, model identity.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <functional>

#include "../../include/tfsc_b200.h"

namespace tfsc {

std::string& last_error_ref();
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

// copy a std::string into a caller buffer; returns strlen or TFSC_E_BUFFER
inline int copy_out(const std::string& s, char* buf, size_t cap) {
  if (!buf || cap < s.size() + 1) return fail(TFSC_E_BUFFER, "buffer too small: need %zu bytes", s.size() + 1);
  memcpy(buf, s.data(), s.size());
  buf[s.size()] = 0;
  return (int)s.size();
}

// cachemanager.go:51-54 ModelIdentifier
struct ModelId {
  std::string name;
  int64_t version = 0;
  bool operator==(const ModelId& o) const { return version == o.version && name == o.name; }
};
struct ModelIdHash {
  size_t operator()(const ModelId& m) const {
    return std::hash<std::string>()(m.name) * 1000003u ^ std::hash<int64_t>()(m.version);
  }
};

}  // namespace tfsc
