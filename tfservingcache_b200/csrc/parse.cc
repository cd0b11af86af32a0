#include "parse.h"

#include <cctype>
#include <climits>

namespace tfsc {

static bool iprefix(const std::string& s, size_t pos, const char* lit) {
  size_t n = strlen(lit);
  if (s.size() < pos + n) return false;
  for (size_t i = 0; i < n; ++i)
    if (tolower((unsigned char)s[pos + i]) != lit[i]) return false;
  return true;
}

// (?i)^/v1/models/(?P<modelName>[^/]+)(/versions/(?P<version>[0-9]+))?
int match_rest_url(const std::string& url, std::string* name, std::string* version) {
  name->clear();
  version->clear();
  if (!iprefix(url, 0, "/v1/models/")) return 404;
  size_t p = 11, q = p;
  while (q < url.size() && url[q] != '/') ++q;
  if (q == p) return 404;  // [^/]+ needs one char
  *name = url.substr(p, q - p);
  if (iprefix(url, q, "/versions/")) {
    size_t d = q + 10, e = d;
    while (e < url.size() && url[e] >= '0' && url[e] <= '9') ++e;
    if (e > d) *version = url.substr(d, e - d);
  }
  return version->empty() ? 400 : 200;
}

const char* rest_error_body(int http_status) {
  switch (http_status) {
    case 404: return "{\"Status\":\"Error\",\"Message\":\"Not found\"}\n";
    case 400: return "{\"Status\":\"Error\",\"Message\":\"Model version must be provided\"}\n";
    default: return "";
  }
}

bool parse_int64(const std::string& s, int64_t* out) {
  size_t i = 0;
  bool neg = false;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
  if (i >= s.size()) return false;
  uint64_t v = 0;
  const uint64_t lim = neg ? (uint64_t)1 << 63 : ((uint64_t)1 << 63) - 1;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    uint64_t d = s[i] - '0';
    if (v > (lim - d) / 10) return false;  // out of range
    v = v * 10 + d;
  }
  *out = neg ? (int64_t)(0 - v) : (int64_t)v;
  return true;
}

bool PbReader::varint(uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}

bool PbReader::next(uint32_t* field, uint32_t* wt, uint64_t* val, const uint8_t** data, size_t* len) {
  uint64_t key;
  if (!varint(&key)) return false;
  *field = (uint32_t)(key >> 3);
  *wt = (uint32_t)(key & 7);
  *val = 0;
  *data = nullptr;
  *len = 0;
  switch (*wt) {
    case 0: return varint(val);
    case 1:
      if (end - p < 8) return false;
      memcpy(val, p, 8);
      *data = p;
      *len = 8;
      p += 8;
      return true;
    case 2: {
      uint64_t n;
      if (!varint(&n) || (uint64_t)(end - p) < n) return false;
      *data = p;
      *len = (size_t)n;
      p += n;
      return true;
    }
    case 5: {
      if (end - p < 4) return false;
      uint32_t v32;
      memcpy(&v32, p, 4);
      *val = v32;
      *data = p;
      *len = 4;
      p += 4;
      return true;
    }
    default: return false;
  }
}

static bool parse_spec(const uint8_t* d, size_t n, std::string* name, bool* has_version, int64_t* version,
                       std::string* signature) {
  PbReader r(d, n);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t l;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &l)) return false;
    if (f == 1 && wt == 2) name->assign((const char*)p, l);
    else if (f == 2 && wt == 2) {  // google.protobuf.Int64Value{value=1}
      *has_version = true;
      *version = 0;
      PbReader r2(p, l);
      while (!r2.done()) {
        uint32_t f2, wt2;
        uint64_t v2;
        const uint8_t* p2;
        size_t l2;
        if (!r2.next(&f2, &wt2, &v2, &p2, &l2)) return false;
        if (f2 == 1 && wt2 == 0) *version = (int64_t)v2;
      }
    } else if (f == 3 && wt == 2 && signature) signature->assign((const char*)p, l);
  }
  return true;
}

bool scan_model_spec(const void* req, size_t len, std::string* name, bool* has_version, int64_t* version,
                     std::string* signature) {
  name->clear();
  *has_version = false;
  *version = 0;
  PbReader r(req, len);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t l;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &l)) return false;
    if (f == 1 && wt == 2 && !parse_spec(p, l, name, has_version, version, signature)) return false;
  }
  return true;
}

}  // namespace tfsc

extern "C" {
int tfsc_rest_match_url(const char* url, char* model_name, size_t name_cap, char* version, size_t version_cap) {
  if (!url) return tfsc::fail(TFSC_E_INVALID, "rest_match_url: null url");
  std::string n, v;
  int st = tfsc::match_rest_url(url, &n, &v);
  if (model_name && tfsc::copy_out(n, model_name, name_cap) < 0) return TFSC_E_BUFFER;
  if (version && tfsc::copy_out(v, version, version_cap) < 0) return TFSC_E_BUFFER;
  return st;
}
const char* tfsc_rest_error_body(int http_status) { return tfsc::rest_error_body(http_status); }
int tfsc_parse_version(const char* version, int64_t* out) {
  int64_t v;
  if (!version || !tfsc::parse_int64(version, &v))
    return tfsc::fail(TFSC_E_INVALID, "strconv.ParseInt: parsing \"%s\": invalid syntax", version ? version : "");
  if (out) *out = v;
  return 0;
}
int tfsc_grpc_model_spec(const void* req, size_t len, char* model_name, size_t name_cap, char* version, size_t version_cap) {
  std::string n;
  bool has;
  int64_t v;
  if (!tfsc::scan_model_spec(req, len, &n, &has, &v, nullptr)) return tfsc::fail(TFSC_E_INVALID, "malformed request");
  if (model_name && tfsc::copy_out(n, model_name, name_cap) < 0) return TFSC_E_BUFFER;
  if (version && tfsc::copy_out(std::to_string(v), version, version_cap) < 0) return TFSC_E_BUFFER;
  return 0;
}
}
