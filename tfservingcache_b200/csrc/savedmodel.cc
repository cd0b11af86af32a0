#include "savedmodel.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>

#include "parse.h"

namespace tfsc {

// ------------------------------------------------------------------------------- crc32c ----
static uint32_t g_c_tab[8][256];
static bool g_c_init = [] {
  for (uint32_t n = 0; n < 256; ++n) {
    uint32_t c = n;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
    g_c_tab[0][n] = c;
  }
  for (uint32_t n = 0; n < 256; ++n)
    for (int t = 1; t < 8; ++t) g_c_tab[t][n] = (g_c_tab[t - 1][n] >> 8) ^ g_c_tab[0][g_c_tab[t - 1][n] & 0xFF];
  return true;
}();

uint32_t crc32c(const void* data, size_t len) {
  const uint8_t* p = (const uint8_t*)data;
  uint32_t c = 0xFFFFFFFFu;
  while (len >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = g_c_tab[7][lo & 0xFF] ^ g_c_tab[6][(lo >> 8) & 0xFF] ^ g_c_tab[5][(lo >> 16) & 0xFF] ^ g_c_tab[4][lo >> 24] ^
        g_c_tab[3][hi & 0xFF] ^ g_c_tab[2][(hi >> 8) & 0xFF] ^ g_c_tab[1][(hi >> 16) & 0xFF] ^ g_c_tab[0][hi >> 24];
    p += 8;
    len -= 8;
  }
  while (len--) c = g_c_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
static uint32_t mask_crc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xA282EAD8u; }

// -------------------------------------------------------------------------------- files ----
static bool slurp(const std::string& path, std::string* out) {
  int fd = open(path.c_str(), O_RDONLY);
  if (fd < 0) return false;
  struct stat st;
  if (fstat(fd, &st) != 0) {
    close(fd);
    return false;
  }
  out->resize((size_t)st.st_size);
  size_t got = 0;
  while (got < out->size()) {
    ssize_t r = read(fd, &(*out)[got], out->size() - got);
    if (r <= 0) break;
    got += (size_t)r;
  }
  close(fd);
  return got == out->size();
}

bool savedmodel_present(const std::string& dir) { return access((dir + "/saved_model.pb").c_str(), R_OK) == 0; }

// ------------------------------------------------------------------ LevelDB-format table ----
static bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
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

static bool read_block(const std::string& f, uint64_t off, uint64_t size, const uint8_t** b, const uint8_t** e, std::string* err) {
  if (off > f.size() || size > f.size() - off || f.size() - off - size < 5) {
    *err = "table block out of range";
    return false;
  }
  const uint8_t* p = (const uint8_t*)f.data() + off;
  uint32_t want;
  memcpy(&want, p + size + 1, 4);
  if (mask_crc(crc32c(p, size + 1)) != want) {
    *err = "table block checksum mismatch";
    return false;
  }
  if (p[size] != 0) {
    *err = "compressed table blocks (snappy) are not supported";
    return false;
  }
  *b = p;
  *e = p + size;
  return true;
}

// entries of one block, keys rebuilt from the shared-prefix encoding
static bool block_entries(const uint8_t* b, const uint8_t* e, std::vector<std::pair<std::string, std::string>>* out, std::string* err) {
  if (e - b < 4) {
    *err = "table block too short";
    return false;
  }
  uint32_t n_restarts;
  memcpy(&n_restarts, e - 4, 4);
  if ((uint64_t)n_restarts * 4 + 4 > (uint64_t)(e - b)) {
    *err = "table block restart array out of range";
    return false;
  }
  const uint8_t* lim = e - 4 - (size_t)n_restarts * 4;
  std::string key;
  const uint8_t* p = b;
  while (p < lim) {
    uint64_t shared, non_shared, vlen;
    if (!get_varint(p, lim, &shared) || !get_varint(p, lim, &non_shared) || !get_varint(p, lim, &vlen) || shared > key.size() ||
        non_shared > (uint64_t)(lim - p) || vlen > (uint64_t)(lim - p) - non_shared) {
      *err = "malformed table entry";
      return false;
    }
    key.resize(shared);
    key.append((const char*)p, non_shared);
    p += non_shared;
    out->emplace_back(key, std::string((const char*)p, vlen));
    p += vlen;
  }
  return true;
}

static bool read_table(const std::string& path, std::map<std::string, std::string>* out, std::string* err) {
  std::string f;
  if (!slurp(path, &f)) {
    *err = path + " not readable";
    return false;
  }
  uint64_t magic = 0;
  if (f.size() >= 48) memcpy(&magic, f.data() + f.size() - 8, 8);
  if (f.size() < 48 || magic != 0xDB4775248B80FB57ull) {
    *err = path + ": not a table file (bad magic)";
    return false;
  }
  const uint8_t* p = (const uint8_t*)f.data() + f.size() - 48;
  const uint8_t* pe = p + 40;
  uint64_t mi_off, mi_size, idx_off, idx_size;
  if (!get_varint(p, pe, &mi_off) || !get_varint(p, pe, &mi_size) || !get_varint(p, pe, &idx_off) || !get_varint(p, pe, &idx_size)) {
    *err = path + ": malformed footer";
    return false;
  }
  const uint8_t *b, *e;
  std::vector<std::pair<std::string, std::string>> index;
  if (!read_block(f, idx_off, idx_size, &b, &e, err) || !block_entries(b, e, &index, err)) return false;
  for (auto& ie : index) {
    const uint8_t* h = (const uint8_t*)ie.second.data();
    const uint8_t* he = h + ie.second.size();
    uint64_t off, size;
    if (!get_varint(h, he, &off) || !get_varint(h, he, &size)) {
      *err = path + ": malformed block handle";
      return false;
    }
    std::vector<std::pair<std::string, std::string>> entries;
    if (!read_block(f, off, size, &b, &e, err) || !block_entries(b, e, &entries, err)) return false;
    for (auto& kv : entries) (*out)[kv.first] = kv.second;
  }
  return true;
}

// -------------------------------------------------------------------------- tensor bundle ----
struct BundleTensor {
  int dtype = 0;
  std::vector<int64_t> shape;
  std::vector<float> f32;  // DT_FLOAT payload
  size_t elems() const {
    size_t n = 1;
    for (auto d : shape) n *= (size_t)d;
    return n;
  }
};

static bool read_bundle(const std::string& prefix, std::map<std::string, BundleTensor>* out, std::string* err) {
  std::map<std::string, std::string> table;
  if (!read_table(prefix + ".index", &table, err)) return false;
  uint64_t num_shards = 1;
  auto hit = table.find("");
  if (hit != table.end()) {
    PbReader r(hit->second.data(), hit->second.size());
    uint32_t f, wt;
    uint64_t v;
    const uint8_t* d;
    size_t n;
    while (!r.done()) {
      if (!r.next(&f, &wt, &v, &d, &n)) {
        *err = "malformed bundle header";
        return false;
      }
      if (f == 1 && wt == 0) num_shards = v;
      if (f == 2 && wt == 0 && v != 0) {
        *err = "big-endian tensor bundles are not supported";
        return false;
      }
    }
  }
  std::map<uint64_t, std::string> shards;
  for (auto& kv : table) {
    if (kv.first.empty()) continue;
    BundleTensor t;
    uint64_t shard = 0, offset = 0, size = 0;
    uint32_t crc = 0;
    bool has_crc = false;
    PbReader r(kv.second.data(), kv.second.size());
    uint32_t f, wt;
    uint64_t v;
    const uint8_t* d;
    size_t n;
    while (!r.done()) {
      if (!r.next(&f, &wt, &v, &d, &n)) {
        *err = "malformed bundle entry for " + kv.first;
        return false;
      }
      if (f == 1 && wt == 0) t.dtype = (int)v;
      else if (f == 2 && wt == 2) {
        PbReader s(d, n);
        uint32_t f2, wt2;
        uint64_t v2;
        const uint8_t* d2;
        size_t n2;
        while (!s.done()) {
          if (!s.next(&f2, &wt2, &v2, &d2, &n2)) {
            *err = "malformed tensor shape for " + kv.first;
            return false;
          }
          if (f2 == 2 && wt2 == 2) {
            int64_t dim = 0;
            PbReader q(d2, n2);
            uint32_t f3, wt3;
            uint64_t v3;
            const uint8_t* d3;
            size_t n3;
            while (!q.done()) {
              if (!q.next(&f3, &wt3, &v3, &d3, &n3)) {
                *err = "malformed tensor dim for " + kv.first;
                return false;
              }
              if (f3 == 1 && wt3 == 0) dim = (int64_t)v3;
            }
            t.shape.push_back(dim);
          }
        }
      } else if (f == 3 && wt == 0) shard = v;
      else if (f == 4 && wt == 0) offset = v;
      else if (f == 5 && wt == 0) size = v;
      else if (f == 6 && wt == 5) {
        crc = (uint32_t)v;
        has_crc = true;
      }
    }
    if (t.dtype != 1) continue;  // only DT_FLOAT variables matter for the supported templates
    for (auto dd : t.shape)
      if (dd < 0 || dd > (int64_t)1 << 40) {
        *err = "tensor " + kv.first + ": bad shape";
        return false;
      }
    if (!shards.count(shard)) {
      char name[64];
      snprintf(name, sizeof name, ".data-%05llu-of-%05llu", (unsigned long long)shard, (unsigned long long)num_shards);
      if (!slurp(prefix + name, &shards[shard])) {
        *err = prefix + name + " not readable";
        return false;
      }
    }
    const std::string& blob = shards[shard];
    if (offset > blob.size() || size > blob.size() - offset || size != t.elems() * 4) {
      *err = "tensor " + kv.first + ": data out of range";
      return false;
    }
    if (has_crc && crc != 0 && mask_crc(crc32c(blob.data() + offset, size)) != crc) {
      *err = "tensor " + kv.first + ": data checksum mismatch";
      return false;
    }
    t.f32.resize(t.elems());
    if (size) memcpy(t.f32.data(), blob.data() + offset, size);
    (*out)[kv.first] = std::move(t);
  }
  return true;
}

// ------------------------------------------------------------------------------ graph side ----
struct GNode {
  std::string op;
  std::vector<std::string> inputs;
};
struct Signature {
  std::map<std::string, std::string> inputs, outputs;  // key -> tensor name
  std::string method;
};

static bool pb_each(const uint8_t* d, size_t n, const std::function<bool(uint32_t, uint32_t, uint64_t, const uint8_t*, size_t)>& fn) {
  PbReader r(d, n);
  uint32_t f, wt;
  uint64_t v;
  const uint8_t* p;
  size_t len;
  while (!r.done()) {
    if (!r.next(&f, &wt, &v, &p, &len)) return false;
    if (!fn(f, wt, v, p, len)) return false;
  }
  return true;
}

static bool parse_saved_model(const std::string& buf, std::map<std::string, GNode>* nodes, std::map<std::string, Signature>* sigs,
                              std::string* err) {
  bool found = false, ok = true;
  auto S = [](const uint8_t* p, size_t n) { return std::string((const char*)p, n); };
  ok = pb_each((const uint8_t*)buf.data(), buf.size(), [&](uint32_t f, uint32_t wt, uint64_t, const uint8_t* mg, size_t mgn) {
    if (f != 2 || wt != 2 || found) return true;  // first MetaGraphDef
    found = true;
    return pb_each(mg, mgn, [&](uint32_t f2, uint32_t wt2, uint64_t, const uint8_t* p2, size_t n2) {
      if (f2 == 2 && wt2 == 2) {  // GraphDef
        return pb_each(p2, n2, [&](uint32_t f3, uint32_t wt3, uint64_t, const uint8_t* p3, size_t n3) {
          if (f3 != 1 || wt3 != 2) return true;
          std::string name;
          GNode g;
          if (!pb_each(p3, n3, [&](uint32_t f4, uint32_t wt4, uint64_t, const uint8_t* p4, size_t n4) {
                if (f4 == 1 && wt4 == 2) name = S(p4, n4);
                else if (f4 == 2 && wt4 == 2) g.op = S(p4, n4);
                else if (f4 == 3 && wt4 == 2) g.inputs.push_back(S(p4, n4));
                return true;
              }))
            return false;
          (*nodes)[name] = std::move(g);
          return true;
        });
      }
      if (f2 == 5 && wt2 == 2) {  // signature_def map entry
        std::string key;
        Signature sg;
        if (!pb_each(p2, n2, [&](uint32_t f3, uint32_t wt3, uint64_t, const uint8_t* p3, size_t n3) {
              if (f3 == 1 && wt3 == 2) key = S(p3, n3);
              else if (f3 == 2 && wt3 == 2)
                return pb_each(p3, n3, [&](uint32_t f4, uint32_t wt4, uint64_t, const uint8_t* p4, size_t n4) {
                  if ((f4 == 1 || f4 == 2) && wt4 == 2) {
                    std::string k, tname;
                    if (!pb_each(p4, n4, [&](uint32_t f5, uint32_t wt5, uint64_t, const uint8_t* p5, size_t n5) {
                          if (f5 == 1 && wt5 == 2) k = S(p5, n5);
                          else if (f5 == 2 && wt5 == 2)
                            return pb_each(p5, n5, [&](uint32_t f6, uint32_t wt6, uint64_t, const uint8_t* p6, size_t n6) {
                              if (f6 == 1 && wt6 == 2) tname = S(p6, n6);
                              return true;
                            });
                          return true;
                        }))
                      return false;
                    (f4 == 1 ? sg.inputs : sg.outputs)[k] = tname;
                  } else if (f4 == 3 && wt4 == 2) {
                    sg.method = S(p4, n4);
                  }
                  return true;
                });
              return true;
            }))
          return false;
        (*sigs)[key] = std::move(sg);
      }
      return true;
    });
  });
  if (!ok) {
    *err = "malformed saved_model.pb";
    return false;
  }
  if (!found) {
    *err = "saved_model.pb holds no MetaGraphDef";
    return false;
  }
  return true;
}

static std::string node_of(const std::string& tensor) {
  size_t b = 0;
  while (b < tensor.size() && tensor[b] == '^') ++b;
  size_t c = tensor.find(':', b);
  return tensor.substr(b, c == std::string::npos ? std::string::npos : c - b);
}

struct Matcher {
  const std::map<std::string, GNode>& nodes;
  const std::map<std::string, BundleTensor>& vars;
  std::string x;

  const GNode* get(const std::string& n) const {
    auto it = nodes.find(n);
    return it == nodes.end() ? nullptr : &it->second;
  }
  std::string skip_identity(std::string n) const {
    for (int i = 0; i < 64; ++i) {
      const GNode* g = get(n);
      if (!g || g->op != "Identity" || g->inputs.empty()) break;
      n = node_of(g->inputs[0]);
    }
    return n;
  }
  bool is_x(const std::string& t) const { return skip_identity(node_of(t)) == x; }
  // follow Identity / ReadVariableOp chains down to a variable node present in the bundle
  std::string variable(std::string n) const {
    for (int i = 0; i < 16; ++i) {
      const GNode* g = get(n);
      if (!g) return "";
      if ((g->op == "VariableV2" || g->op == "Variable" || g->op == "VarHandleOp") && vars.count(n)) return n;
      if ((g->op == "Identity" || g->op == "ReadVariableOp") && !g->inputs.empty()) {
        n = node_of(g->inputs[0]);
        continue;
      }
      return "";
    }
    return "";
  }
};

static void json_str(const std::string& s, std::string* o) {
  o->push_back('"');
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') {
      o->push_back('\\');
      o->push_back((char)c);
    } else if (c < 0x20) {
      char b[8];
      snprintf(b, sizeof b, "\\u%04x", c);
      o->append(b);
    } else {
      o->push_back((char)c);
    }
  }
  o->push_back('"');
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

bool savedmodel_import(const std::string& dir, SavedModelBundle* out, std::string* err) {
  std::string pb;
  if (!slurp(dir + "/saved_model.pb", &pb)) {
    *err = dir + "/saved_model.pb not readable";
    return false;
  }
  std::map<std::string, GNode> nodes;
  std::map<std::string, Signature> sigs;
  if (!parse_saved_model(pb, &nodes, &sigs, err)) return false;
  std::map<std::string, BundleTensor> vars;
  if (!read_bundle(dir + "/variables/variables", &vars, err)) return false;

  const Signature* sg = nullptr;
  auto sit = sigs.find("serving_default");
  if (sit != sigs.end()) sg = &sit->second;
  else
    for (auto& kv : sigs)  // std::map order = the sorted-key choice of the Python importer
      if (kv.second.method.size() >= 7 && kv.second.method.compare(kv.second.method.size() - 7, 7, "predict") == 0) {
        sg = &kv.second;
        break;
      }
  if (!sg) {
    *err = "signature 'serving_default' not found and no predict signature present";
    return false;
  }
  if (sg->inputs.size() != 1 || sg->outputs.size() != 1) {
    *err = "only single-input single-output predict signatures are supported";
    return false;
  }
  const std::string in_key = sg->inputs.begin()->first, out_key = sg->outputs.begin()->first;
  Matcher m{nodes, vars, node_of(sg->inputs.begin()->second)};
  std::string cur = m.skip_identity(node_of(sg->outputs.begin()->second));
  const GNode* g = m.get(cur);
  if (!g) {
    *err = "output node " + cur + " not in graph";
    return false;
  }
  std::string sig_json = "\"signature\": {\"input\": ";
  json_str(in_key, &sig_json);
  sig_json += ", \"output\": ";
  json_str(out_key, &sig_json);
  sig_json += "}";
  // classify / regress signatures of the SavedModel are kept (tensorflow/serving/{classify,regress}): their tf.Example
  // feature is the predict signature's input key (half_plus_two: feature "x" for regress_x_to_y / classify_x_to_y)
  std::string extra;
  for (auto& kv : sigs) {
    const std::string& mth = kv.second.method;
    const char* kind = mth.size() >= 8 && mth.compare(mth.size() - 8, 8, "classify") == 0 ? "classify"
                       : mth.size() >= 7 && mth.compare(mth.size() - 7, 7, "regress") == 0 ? "regress" : nullptr;
    if (!kind) continue;
    extra += extra.empty() ? "" : ", ";
    extra += "{\"name\": ";
    json_str(kv.first, &extra);
    extra += std::string(", \"method\": \"") + kind + "\", \"feature\": ";
    json_str(in_key, &extra);
    extra += "}";
  }
  if (!extra.empty()) sig_json += ", \"extra_signatures\": [" + extra + "]";

  // ---- affine: y = Add(Mul(a, x), b), scalar variables
  if ((g->op == "Add" || g->op == "AddV2") && g->inputs.size() == 2) {
    for (int o = 0; o < 2; ++o) {
      const std::string mul = node_of(g->inputs[o]);
      const GNode* mg = m.get(mul);
      if (!mg || mg->op != "Mul" || mg->inputs.size() != 2) continue;
      for (int p = 0; p < 2; ++p) {
        const std::string av = m.variable(node_of(mg->inputs[p])), bv = m.variable(node_of(g->inputs[1 - o]));
        if (av.empty() || bv.empty() || !m.is_x(mg->inputs[1 - p])) continue;
        const BundleTensor &A = vars.at(av), &Bv = vars.at(bv);
        if (A.f32.size() != 1 || Bv.f32.size() != 1) continue;
        out->weights.assign(512, 0);
        memcpy(out->weights.data(), &A.f32[0], 4);
        memcpy(out->weights.data() + 256, &Bv.f32[0], 4);
        out->manifest_json = "{\"format\": \"tfsc-b200-v1\", \"template\": \"affine\", \"dtype\": \"float32\", " + sig_json +
                             ", \"a_offset\": 0, \"b_offset\": 256, \"weights_bytes\": 512}";
        return true;
      }
    }
  }
  // ---- dense MLP: walk back from the output through [Relu] <- BiasAdd/Add <- MatMul
  struct Layer {
    const BundleTensor *w, *b;
    bool relu;
  };
  std::vector<Layer> layers;
  for (int guard = 0; guard < 4096 && !m.is_x(cur); ++guard) {
    g = m.get(cur);
    if (!g) {
      *err = "node " + cur + " not in graph";
      return false;
    }
    bool relu = false;
    if (g->op == "Relu" && !g->inputs.empty()) {
      relu = true;
      cur = node_of(g->inputs[0]);
      g = m.get(cur);
      if (!g) {
        *err = "node " + cur + " not in graph";
        return false;
      }
    }
    if ((g->op != "BiasAdd" && g->op != "Add" && g->op != "AddV2") || g->inputs.size() != 2) {
      *err = "unsupported op '" + g->op + "' at node '" + cur + "': not an affine or dense-MLP graph";
      return false;
    }
    std::string mm, bias;
    for (int o = 0; o < 2; ++o) {
      const GNode* c = m.get(node_of(g->inputs[o]));
      if (c && c->op == "MatMul") {
        mm = node_of(g->inputs[o]);
        bias = m.variable(node_of(g->inputs[1 - o]));
      }
    }
    const GNode* mg = mm.empty() ? nullptr : m.get(mm);
    if (!mg || bias.empty() || mg->inputs.size() != 2) {
      *err = "node '" + cur + "': expected MatMul + bias variable";
      return false;
    }
    const std::string wv = m.variable(node_of(mg->inputs[1]));
    if (wv.empty() || vars.at(wv).shape.size() != 2) {
      *err = "node '" + mm + "': MatMul weight is not a 2-D variable";
      return false;
    }
    const BundleTensor &W = vars.at(wv), &Bv = vars.at(bias);
    if ((int64_t)Bv.f32.size() != W.shape[1]) {
      *err = "node '" + cur + "': bias length does not match the kernel";
      return false;
    }
    layers.push_back({&W, &Bv, relu});
    cur = m.skip_identity(node_of(mg->inputs[0]));
  }
  if (layers.empty() || !m.is_x(cur)) {
    *err = "graph is neither y = a*x + b nor a dense MLP";
    return false;
  }
  std::string lj;
  size_t off = 0;
  std::vector<std::pair<size_t, size_t>> offs;
  for (size_t i = layers.size(); i-- > 0;) {
    const Layer& L = layers[i];
    const size_t w_off = off;
    off = align256(off + L.w->f32.size() * 4);
    const size_t b_off = off;
    off = align256(off + L.b->f32.size() * 4);
    offs.push_back({w_off, b_off});
    if (!lj.empty()) lj += ", ";
    lj += "{\"in\": " + std::to_string(L.w->shape[0]) + ", \"out\": " + std::to_string(L.w->shape[1]) + ", \"activation\": \"" +
          (L.relu ? "relu" : "linear") + "\", \"w_offset\": " + std::to_string(w_off) + ", \"b_offset\": " + std::to_string(b_off) + "}";
  }
  out->weights.assign(off, 0);
  size_t k = 0;
  for (size_t i = layers.size(); i-- > 0; ++k) {
    memcpy(out->weights.data() + offs[k].first, layers[i].w->f32.data(), layers[i].w->f32.size() * 4);
    memcpy(out->weights.data() + offs[k].second, layers[i].b->f32.data(), layers[i].b->f32.size() * 4);
  }
  out->manifest_json = "{\"format\": \"tfsc-b200-v1\", \"template\": \"mlp\", \"dtype\": \"float32\", " + sig_json + ", \"layers\": [" + lj +
                       "], \"weights_bytes\": " + std::to_string(off) + "}";
  return true;
}

}  // namespace tfsc
