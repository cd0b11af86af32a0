// Small JSON value + parser + writer (config strings, model manifests, TF-Serving REST bodies).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tfsc {

struct Json {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  double num = 0;
  bool is_int = false;
  int64_t i64 = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

  const Json* get(const std::string& k) const {
    if (type != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  double number(double dflt = 0) const { return type == Num ? num : dflt; }
  int64_t integer(int64_t dflt = 0) const { return type == Num ? (is_int ? i64 : (int64_t)num) : dflt; }
  std::string string(const std::string& dflt = "") const { return type == Str ? str : dflt; }
  int64_t get_int(const std::string& k, int64_t dflt) const {
    const Json* j = get(k);
    if (!j) return dflt;
    if (j->type == Str) return strtoll(j->str.c_str(), nullptr, 10);
    if (j->type == Bool) return j->b;
    return j->integer(dflt);
  }
  double get_num(const std::string& k, double dflt) const {
    const Json* j = get(k);
    if (!j) return dflt;
    if (j->type == Str) return strtod(j->str.c_str(), nullptr);
    return j->number(dflt);
  }
  std::string get_str(const std::string& k, const std::string& dflt) const {
    const Json* j = get(k);
    return j && j->type == Str ? j->str : dflt;
  }
};

class JsonParser {
 public:
  JsonParser(const char* s, size_t n) : p_(s), e_(s + n) {}
  bool parse(Json* out, std::string* err) {
    if (!value(out, 0)) {
      if (err) *err = err_.empty() ? "JSON parse error" : err_;
      return false;
    }
    ws();
    if (p_ != e_) {
      if (err) *err = "trailing characters after JSON value";
      return false;
    }
    return true;
  }

 private:
  const char* p_;
  const char* e_;
  std::string err_;
  void ws() {
    while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
  }
  bool lit(const char* s) {
    size_t n = strlen(s);
    if ((size_t)(e_ - p_) < n || memcmp(p_, s, n)) return false;
    p_ += n;
    return true;
  }
  bool str(std::string* out) {
    if (p_ >= e_ || *p_ != '"') return false;
    ++p_;
    out->clear();
    while (p_ < e_ && *p_ != '"') {
      if (*p_ == '\\') {
        if (++p_ >= e_) return false;
        switch (*p_) {
          case 'n': out->push_back('\n'); break;
          case 't': out->push_back('\t'); break;
          case 'r': out->push_back('\r'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'u': {
            if (e_ - p_ < 5) return false;
            unsigned cp = (unsigned)strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out->push_back((char)cp);
            else if (cp < 0x800) {
              out->push_back((char)(0xC0 | (cp >> 6)));
              out->push_back((char)(0x80 | (cp & 0x3F)));
            } else {
              out->push_back((char)(0xE0 | (cp >> 12)));
              out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
              out->push_back((char)(0x80 | (cp & 0x3F)));
            }
            break;
          }
          default: out->push_back(*p_);
        }
        ++p_;
      } else {
        out->push_back(*p_++);
      }
    }
    if (p_ >= e_) return false;
    ++p_;
    return true;
  }
  bool value(Json* out, int depth) {
    if (depth > 64) {
      err_ = "JSON nesting too deep";
      return false;
    }
    ws();
    if (p_ >= e_) return false;
    char c = *p_;
    if (c == '{') {
      ++p_;
      out->type = Json::Obj;
      ws();
      if (p_ < e_ && *p_ == '}') {
        ++p_;
        return true;
      }
      while (true) {
        ws();
        std::string k;
        if (!str(&k)) return false;
        ws();
        if (p_ >= e_ || *p_++ != ':') return false;
        out->obj.emplace_back(k, Json());
        if (!value(&out->obj.back().second, depth + 1)) return false;
        ws();
        if (p_ < e_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < e_ && *p_ == '}') {
          ++p_;
          return true;
        }
        return false;
      }
    }
    if (c == '[') {
      ++p_;
      out->type = Json::Arr;
      ws();
      if (p_ < e_ && *p_ == ']') {
        ++p_;
        return true;
      }
      while (true) {
        out->arr.emplace_back();
        if (!value(&out->arr.back(), depth + 1)) return false;
        ws();
        if (p_ < e_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < e_ && *p_ == ']') {
          ++p_;
          return true;
        }
        return false;
      }
    }
    if (c == '"') {
      out->type = Json::Str;
      return str(&out->str);
    }
    if (lit("true")) {
      out->type = Json::Bool;
      out->b = true;
      return true;
    }
    if (lit("false")) {
      out->type = Json::Bool;
      out->b = false;
      return true;
    }
    if (lit("null")) {
      out->type = Json::Null;
      return true;
    }
    // number
    const char* s = p_;
    bool isint = true;
    if (p_ < e_ && (*p_ == '-' || *p_ == '+')) ++p_;
    while (p_ < e_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '-' || *p_ == '+')) {
      if (*p_ == '.' || *p_ == 'e' || *p_ == 'E') isint = false;
      ++p_;
    }
    if (p_ == s) return false;
    std::string t(s, p_ - s);
    out->type = Json::Num;
    out->num = strtod(t.c_str(), nullptr);
    out->is_int = isint;
    if (isint) out->i64 = strtoll(t.c_str(), nullptr, 10);
    return true;
  }
};

inline bool json_parse(const std::string& s, Json* out, std::string* err) {
  JsonParser p(s.data(), s.size());
  return p.parse(out, err);
}

inline void json_escape(const std::string& s, std::string* out) {
  out->push_back('"');
  for (char c : s) {
    switch (c) {
      case '"': *out += "\\\""; break;
      case '\\': *out += "\\\\"; break;
      case '\n': *out += "\\n"; break;
      case '\t': *out += "\\t"; break;
      case '\r': *out += "\\r"; break;
      default:
        if ((unsigned char)c < 0x20) {
          char b[8];
          snprintf(b, sizeof b, "\\u%04x", c);
          *out += b;
        } else {
          out->push_back(c);
        }
    }
  }
  out->push_back('"');
}

// shortest decimal that round-trips an fp32, always with a fractional part for finite values
// (TF-Serving prints 3.0, 2.5, 4.5 -- deploy/docker-compose/readme.md:42)
inline void json_float(float v, std::string* out) {
  if (std::isnan(v)) { *out += "NaN"; return; }
  if (std::isinf(v)) { *out += v > 0 ? "Infinity" : "-Infinity"; return; }
  char b[32];
  for (int prec = 1; prec <= 9; ++prec) {
    snprintf(b, sizeof b, "%.*g", prec, (double)v);
    if (strtof(b, nullptr) == v) break;
  }
  *out += b;
  if (!strpbrk(b, ".eE")) *out += ".0";
}

}  // namespace tfsc
