#include "model.h"

namespace tfsc {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void finish(ModelDesc* d) {
  if (d->tmpl == Template::Mlp) {
    d->in_dim = d->layers.front().in;
    d->out_dim = d->layers.back().out;
    d->max_width = 0;
    for (auto& l : d->layers) {
      if (l.in > d->max_width) d->max_width = l.in;
      if (l.out > d->max_width) d->max_width = l.out;
    }
  } else {
    d->in_dim = d->out_dim = 0;
  }
}

ModelDesc make_mlp_desc(const std::vector<int>& dims, const std::vector<std::string>& activations) {
  ModelDesc d;
  d.tmpl = Template::Mlp;
  size_t off = 0;
  for (size_t l = 0; l + 1 < dims.size(); ++l) {
    DenseLayer L;
    L.in = dims[l];
    L.out = dims[l + 1];
    if (!activations.empty()) L.relu = activations[l] == "relu";
    else L.relu = (l + 2 < dims.size());  // relu on all but the last layer
    L.w_off = off;
    off = align256(off + (size_t)L.in * L.out * 4);
    L.b_off = off;
    off = align256(off + (size_t)L.out * 4);
    d.layers.push_back(L);
  }
  d.weights_bytes = off;
  finish(&d);
  return d;
}

ModelDesc make_affine_desc() {
  ModelDesc d;
  d.tmpl = Template::Affine;
  d.a_off = 0;
  d.b_off = 256;
  d.weights_bytes = 512;
  finish(&d);
  return d;
}

bool parse_manifest(const Json& j, ModelDesc* d, std::string* err) {
  if (j.type != Json::Obj) {
    *err = "manifest is not a JSON object";
    return false;
  }
  if (j.get_str("format", "") != "tfsc-b200-v1") {
    *err = "unsupported manifest format '" + j.get_str("format", "") + "'";
    return false;
  }
  if (j.get_str("dtype", "float32") != "float32") {
    *err = "only float32 bundles are supported";
    return false;
  }
  if (const Json* sig = j.get("signature")) {
    d->input_name = sig->get_str("input", "x");
    d->output_name = sig->get_str("output", "y");
  }
  d->weights_bytes = (size_t)j.get_int("weights_bytes", 0);
  std::string t = j.get_str("template", "");
  if (t == "affine") {
    d->tmpl = Template::Affine;
    d->a_off = (size_t)j.get_int("a_offset", 0);
    d->b_off = (size_t)j.get_int("b_offset", 256);
    if (d->a_off + 4 > d->weights_bytes || d->b_off + 4 > d->weights_bytes || (d->a_off & 3) || (d->b_off & 3)) {
      *err = "affine offsets out of range";
      return false;
    }
  } else if (t == "mlp") {
    d->tmpl = Template::Mlp;
    const Json* layers = j.get("layers");
    if (!layers || layers->type != Json::Arr || layers->arr.empty()) {
      *err = "mlp manifest needs a non-empty 'layers' array";
      return false;
    }
    int prev_out = -1;
    for (auto& lj : layers->arr) {
      DenseLayer L;
      L.in = (int)lj.get_int("in", 0);
      L.out = (int)lj.get_int("out", 0);
      L.relu = lj.get_str("activation", "linear") == "relu";
      L.w_off = (size_t)lj.get_int("w_offset", -1);
      L.b_off = (size_t)lj.get_int("b_offset", -1);
      if (L.in <= 0 || L.out <= 0 || (L.w_off & 255) || (L.b_off & 255) ||
          L.w_off + (size_t)L.in * L.out * 4 > d->weights_bytes || L.b_off + (size_t)L.out * 4 > d->weights_bytes) {
        *err = "mlp layer out of range or misaligned";
        return false;
      }
      if (prev_out >= 0 && prev_out != L.in) {
        *err = "mlp layer dims do not chain";
        return false;
      }
      prev_out = L.out;
      d->layers.push_back(L);
    }
  } else {
    *err = "unknown template '" + t + "'";
    return false;
  }
  finish(d);
  return true;
}

std::string manifest_json(const ModelDesc& d) {
  std::string s = "{\"format\":\"tfsc-b200-v1\",\"dtype\":\"float32\",\"signature\":{\"input\":";
  json_escape(d.input_name, &s);
  s += ",\"output\":";
  json_escape(d.output_name, &s);
  s += "},\"weights_bytes\":" + std::to_string(d.weights_bytes);
  if (d.tmpl == Template::Affine) {
    s += ",\"template\":\"affine\",\"a_offset\":" + std::to_string(d.a_off) + ",\"b_offset\":" + std::to_string(d.b_off);
  } else {
    s += ",\"template\":\"mlp\",\"layers\":[";
    for (size_t i = 0; i < d.layers.size(); ++i) {
      auto& L = d.layers[i];
      if (i) s += ",";
      s += "{\"in\":" + std::to_string(L.in) + ",\"out\":" + std::to_string(L.out) + ",\"activation\":\"" +
           (L.relu ? "relu" : "linear") + "\",\"w_offset\":" + std::to_string(L.w_off) +
           ",\"b_offset\":" + std::to_string(L.b_off) + "}";
    }
    s += "]";
  }
  s += "}";
  return s;
}

}  // namespace tfsc
