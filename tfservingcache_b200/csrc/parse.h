// Request parsing of the protocol shim: pkg/tfservingproxy/tfservingproxy.go:24 (URL regex),
// :93-129 (404 / 400), :246-250 (gRPC ModelSpec -> name, version string) and
// pkg/cachemanager/cachemanager.go:297 (ParseInt).
#pragma once
#include <string>

#include "common.h"

namespace tfsc {

// returns 200 / 404 / 400; on 200 fills name + verbatim version string; on 400 fills name
int match_rest_url(const std::string& url, std::string* name, std::string* version);
const char* rest_error_body(int http_status);
bool parse_int64(const std::string& s, int64_t* out);  // strconv.ParseInt(s, 10, 64)

// minimal protobuf wire reader (shared with wire.cc)
struct PbReader {
  const uint8_t* p;
  const uint8_t* end;
  PbReader(const void* d, size_t n) : p((const uint8_t*)d), end((const uint8_t*)d + n) {}
  bool done() const { return p >= end; }
  bool varint(uint64_t* v);
  // reads one field header + payload. For LEN: data/len set; for VARINT: val; FIXED32/64: val.
  bool next(uint32_t* field, uint32_t* wt, uint64_t* val, const uint8_t** data, size_t* len);
};

// model_spec (field 1 of Predict/Classify/Regress/GetModelMetadata/SessionRun requests)
bool scan_model_spec(const void* req, size_t len, std::string* name, bool* has_version, int64_t* version,
                     std::string* signature);

}  // namespace tfsc
