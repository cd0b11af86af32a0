#include "wire.h"

#include "parse.h"

namespace tfsc {

static void put_varint(std::string* s, uint64_t v) {
  while (v >= 0x80) {
    s->push_back((char)(v | 0x80));
    v >>= 7;
  }
  s->push_back((char)v);
}
static void put_tag(std::string* s, uint32_t field, uint32_t wt) { put_varint(s, (field << 3) | wt); }
static void put_ld(std::string* s, uint32_t field, const std::string& payload) {
  put_tag(s, field, 2);
  put_varint(s, payload.size());
  s->append(payload);
}

static bool decode_shape(const uint8_t* d, size_t n, std::vector<int64_t>* shape) {
  PbReader r(d, n);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t l;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &l)) return false;
    if (f == 2 && wt == 2) {  // dim
      int64_t size = 0;
      PbReader r2(p, l);
      while (!r2.done()) {
        uint32_t f2, wt2;
        uint64_t v2;
        const uint8_t* p2;
        size_t l2;
        if (!r2.next(&f2, &wt2, &v2, &p2, &l2)) return false;
        if (f2 == 1 && wt2 == 0) size = (int64_t)v2;
      }
      shape->push_back(size);
    }
  }
  return true;
}

static bool decode_tensor(const uint8_t* d, size_t n, TensorView* t) {
  PbReader r(d, n);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t l;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &l)) return false;
    switch (f) {
      case 1: t->dtype = (int)v; break;
      case 2:
        if (wt == 2 && !decode_shape(p, l, &t->shape)) return false;
        break;
      case 4:
        if (wt == 2) {
          t->content = p;
          t->content_len = l;
        }
        break;
      case 5:  // float_val: packed (wt 2) or one fixed32 per entry (wt 5)
        if (wt == 2) {
          if (!t->packed_f32 && t->loose_f32.empty()) {
            t->packed_f32 = p;
            t->packed_f32_len = l;
          } else {  // several packed chunks: concatenate
            if (t->packed_f32) {
              t->loose_f32.resize(t->packed_f32_len / 4);
              memcpy(t->loose_f32.data(), t->packed_f32, t->packed_f32_len / 4 * 4);
              t->packed_f32 = nullptr;
            }
            size_t old = t->loose_f32.size();
            t->loose_f32.resize(old + l / 4);
            memcpy(t->loose_f32.data() + old, p, l / 4 * 4);
          }
        } else if (wt == 5) {
          float fv;
          uint32_t u = (uint32_t)v;
          memcpy(&fv, &u, 4);
          if (t->packed_f32) {
            t->loose_f32.resize(t->packed_f32_len / 4);
            memcpy(t->loose_f32.data(), t->packed_f32, t->packed_f32_len / 4 * 4);
            t->packed_f32 = nullptr;
          }
          t->loose_f32.push_back(fv);
        }
        break;
      case 7:  // int_val: packed varints (wt 2) or one varint per entry (wt 0)
        if (wt == 0) t->ints.push_back((int32_t)(int64_t)v);
        else if (wt == 2) {
          PbReader r2(p, l);
          uint64_t vv;
          while (!r2.done()) {
            if (!r2.varint(&vv)) return false;
            t->ints.push_back((int32_t)(int64_t)vv);
          }
        }
        break;
      default: break;  // other typed value fields are rejected later by dtype
    }
  }
  return true;
}

bool decode_predict_request(const void* data, size_t len, PredictRequestView* out, std::string* err) {
  PbReader r(data, len);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t l;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &l)) {
      *err = "malformed PredictRequest";
      return false;
    }
    if (f == 1 && wt == 2) {
      // model_spec: re-wrap as a one-field message for the shared scanner
      std::string wrapped;
      put_ld(&wrapped, 1, std::string((const char*)p, l));
      if (!scan_model_spec(wrapped.data(), wrapped.size(), &out->model_name, &out->has_version, &out->version,
                           &out->signature_name)) {
        *err = "malformed ModelSpec";
        return false;
      }
    } else if (f == 2 && wt == 2) {  // map<string, TensorProto> entry
      TensorView t;
      PbReader r2(p, l);
      while (!r2.done()) {
        uint32_t f2, wt2;
        uint64_t v2;
        const uint8_t* p2;
        size_t l2;
        if (!r2.next(&f2, &wt2, &v2, &p2, &l2)) {
          *err = "malformed inputs entry";
          return false;
        }
        if (f2 == 1 && wt2 == 2) t.name.assign((const char*)p2, l2);
        else if (f2 == 2 && wt2 == 2 && !decode_tensor(p2, l2, &t)) {
          *err = "malformed TensorProto";
          return false;
        }
      }
      out->inputs.push_back(std::move(t));
    } else if (f == 3 && wt == 2) {
      out->output_filter.emplace_back((const char*)p, l);
    }
  }
  return true;
}

// 16 Mi elements = 64 MB = 4x the reference's 16 MiB gRPC message limit (cachemanager.go:230-233)
static constexpr int64_t kMaxBroadcastElements = (int64_t)16 << 20;

bool tensor_f32(const TensorView& t, const float** data, int64_t* n, std::vector<float>* scratch, std::string* err) {
  if (t.dtype != TFSC_DT_FLOAT) {
    *err = "input '" + t.name + "' has dtype " + std::to_string(t.dtype) + "; only DT_FLOAT (1) is supported";
    return false;
  }
  for (auto d : t.shape)
    if (d < 0) {
      *err = "input '" + t.name + "' has an unknown dimension";
      return false;
    }
  const int64_t want = t.num_elements();
  if (want < 0) {
    *err = "input '" + t.name + "': tensor_shape is too large";
    return false;
  }
  if (t.content_len) {
    if ((int64_t)(t.content_len / 4) != want || t.content_len % 4) {
      *err = "tensor_content size does not match tensor_shape";
      return false;
    }
    if (reinterpret_cast<uintptr_t>(t.content) % 4 == 0) {
      *data = reinterpret_cast<const float*>(t.content);
    } else {  // protobuf payloads are byte-aligned
      scratch->resize(want);
      memcpy(scratch->data(), t.content, t.content_len);
      *data = scratch->data();
    }
    *n = want;
    return true;
  }
  const uint8_t* src = t.packed_f32;
  int64_t have = (int64_t)(t.packed_f32_len / 4);
  if (!src) {
    src = reinterpret_cast<const uint8_t*>(t.loose_f32.data());
    have = (int64_t)t.loose_f32.size();
  }
  if (have == want) {
    if (reinterpret_cast<uintptr_t>(src) % 4 == 0) {
      *data = reinterpret_cast<const float*>(src);
    } else {
      scratch->resize(want);
      memcpy(scratch->data(), src, (size_t)want * 4);
      *data = scratch->data();
    }
    *n = want;
    return true;
  }
  if (have == 1 && want > 1) {  // TF semantics: a single value fills the tensor
    if (want > kMaxBroadcastElements) {  // the only path where a few request bytes size a large allocation
      *err = "input '" + t.name + "': scalar broadcast to " + std::to_string(want) + " elements exceeds the limit of " +
             std::to_string(kMaxBroadcastElements);
      return false;
    }
    float fv;
    memcpy(&fv, src, 4);
    scratch->assign(want, fv);
    *data = scratch->data();
    *n = want;
    return true;
  }
  *err = "float_val count " + std::to_string(have) + " does not match tensor_shape (" + std::to_string(want) + ")";
  return false;
}

bool tensor_i32(const TensorView& t, const int32_t** data, int64_t* n, std::vector<int32_t>* scratch, std::string* err) {
  if (t.dtype != TFSC_DT_INT32) {
    *err = "input '" + t.name + "' has dtype " + std::to_string(t.dtype) + "; expected DT_INT32 (3)";
    return false;
  }
  for (auto d : t.shape)
    if (d < 0) {
      *err = "input '" + t.name + "' has an unknown dimension";
      return false;
    }
  const int64_t want = t.num_elements();
  if (want < 0) {
    *err = "input '" + t.name + "': tensor_shape is too large";
    return false;
  }
  if (t.content_len) {
    if ((int64_t)(t.content_len / 4) != want || t.content_len % 4) {
      *err = "tensor_content size does not match tensor_shape";
      return false;
    }
    scratch->resize(want);
    memcpy(scratch->data(), t.content, t.content_len);
  } else if ((int64_t)t.ints.size() == want) {
    *scratch = t.ints;
  } else if (t.ints.size() == 1 && want > 1) {
    if (want > kMaxBroadcastElements) {
      *err = "input '" + t.name + "': scalar broadcast to " + std::to_string(want) + " elements exceeds the limit of " +
             std::to_string(kMaxBroadcastElements);
      return false;
    }
    scratch->assign(want, t.ints[0]);
  } else {
    *err = "int_val count " + std::to_string(t.ints.size()) + " does not match tensor_shape (" + std::to_string(want) + ")";
    return false;
  }
  *data = scratch->data();
  *n = want;
  return true;
}

void predict_response_frame(const std::string& model_name, int64_t version, const std::string& signature_name,
                            const std::string& output_name, const std::vector<int64_t>& shape, std::string* prefix,
                            std::string* suffix) {
  int64_t n = 1;
  for (auto d : shape) n *= d;
  const size_t payload = (size_t)n * 4;
  // TensorProto head: dtype, tensor_shape, then the float_val length header
  std::string thead;
  put_tag(&thead, 1, 0);
  put_varint(&thead, TFSC_DT_FLOAT);
  std::string sh;
  for (auto d : shape) {
    std::string dim;
    if (d != 0) {
      put_tag(&dim, 1, 0);
      put_varint(&dim, (uint64_t)d);
    }
    put_ld(&sh, 2, dim);
  }
  put_ld(&thead, 2, sh);
  if (payload) {
    put_tag(&thead, 5, 2);
    put_varint(&thead, payload);
  }
  const size_t tensor_len = thead.size() + payload;
  std::string ehead;  // map entry: key, then value header
  put_ld(&ehead, 1, output_name);
  put_tag(&ehead, 2, 2);
  put_varint(&ehead, tensor_len);
  const size_t entry_len = ehead.size() + tensor_len;
  prefix->clear();
  put_tag(prefix, 1, 2);
  put_varint(prefix, entry_len);
  prefix->append(ehead);
  prefix->append(thead);
  // model_spec
  std::string spec;
  if (!model_name.empty()) put_ld(&spec, 1, model_name);
  std::string ver;
  if (version != 0) {
    put_tag(&ver, 1, 0);
    put_varint(&ver, (uint64_t)version);
  }
  put_ld(&spec, 2, ver);
  if (!signature_name.empty()) put_ld(&spec, 3, signature_name);
  suffix->clear();
  put_ld(suffix, 2, spec);
}

}  // namespace tfsc
