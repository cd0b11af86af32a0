#include "model.h"

#include <algorithm>

namespace tfsc {

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t ModelDesc::scratch_bytes(int64_t rows) const {
  if (tmpl == Template::Mlp) return 2 * (((size_t)rows * (size_t)(max_width > 0 ? max_width : 1) * 4 + 255) & ~(size_t)255);
  if (tmpl == Template::Graph) return ((size_t)rows * (size_t)(n_buffers * buf_elems + col_elems) * 4 + 255) & ~(size_t)255;
  return 256;
}

static void finish(ModelDesc* d) {
  if (d->tmpl == Template::Graph) {
    d->in_dim = 1;
    for (auto v : d->input_shape) d->in_dim *= v;
    d->out_dim = 1;
    for (auto v : d->output_shape) d->out_dim *= v;
    return;
  }
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
  if (const Json* ex = j.get("extra_signatures")) {
    for (auto& e : ex->arr) {
      ExtraSignature g;
      g.name = e.get_str("name", "");
      const std::string m = e.get_str("method", "");
      g.method = m == "classify" ? 1 : m == "regress" ? 2 : 0;
      g.feature = e.get_str("feature", d->input_name);
      if (g.name.empty() || g.method == 0) {
        *err = "extra_signatures entries need a name and method classify|regress";
        return false;
      }
      d->extra_sigs.push_back(g);
    }
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
  } else if (t == "graph") {
    d->tmpl = Template::Graph;
    const Json* ops = j.get("ops");
    const Json* ish = j.get("input_shape");
    if (!ops || ops->type != Json::Arr || ops->arr.empty() || !ish || ish->type != Json::Arr) {
      *err = "graph manifest needs 'ops' and 'input_shape'";
      return false;
    }
    for (auto& v : ish->arr) d->input_shape.push_back(v.integer());
    d->input_dtype = j.get_str("input_dtype", "float32") == "int32" ? TFSC_DT_INT32 : TFSC_DT_FLOAT;
    d->n_buffers = (int)j.get_int("n_buffers", 0);
    if (d->n_buffers < 1 || d->n_buffers > 16) {
      *err = "graph manifest: n_buffers out of range";
      return false;
    }
    int64_t in_elems = 1;
    for (auto v : d->input_shape) in_elems *= v;
    std::vector<int64_t> written(d->n_buffers, -1);  // elements per image held by each scratch buffer
    int64_t out_elems = -1;
    for (auto& oj : ops->arr) {
      GraphOp o;
      const std::string kind = oj.get_str("op", "");
      if (kind == "conv") o.kind = OpKind::Conv;
      else if (kind == "maxpool") o.kind = OpKind::MaxPool;
      else if (kind == "avgpool") o.kind = OpKind::AvgPool;
      else if (kind == "dense") o.kind = OpKind::Dense;
      else if (kind == "embed") o.kind = OpKind::Embed;
      else if (kind == "layernorm") o.kind = OpKind::LayerNorm;
      else if (kind == "attention") o.kind = OpKind::Attention;
      else {
        *err = "graph manifest: unknown op '" + kind + "'";
        return false;
      }
      o.src = (int)oj.get_int("src", -1);
      o.dst = (int)oj.get_int("dst", 0);
      o.res = (int)oj.get_int("res", -100);
      o.h = (int)oj.get_int("h", 1);
      o.w = (int)oj.get_int("w", 1);
      o.c = (int)oj.get_int("c", 1);
      o.kh = (int)oj.get_int("kh", 1);
      o.kw = (int)oj.get_int("kw", 1);
      o.stride = (int)oj.get_int("stride", 1);
      o.pad = (int)oj.get_int("pad", 0);
      o.cout = (int)oj.get_int("cout", o.c);
      const std::string act = oj.get_str("act", "none");
      o.act = act == "relu" ? 1 : act == "gelu" ? 2 : act == "tanh" ? 3 : 0;
      o.heads = (int)oj.get_int("heads", 1);
      o.vocab = (int)oj.get_int("vocab", 0);
      o.max_pos = (int)oj.get_int("max_pos", 0);
      o.word_off = (size_t)oj.get_int("word_offset", 0);
      o.pos_off = (size_t)oj.get_int("pos_offset", 0);
      o.type_off = (size_t)oj.get_int("type_offset", 0);
      o.eps = (float)oj.get_num("eps", 1e-12);
      o.w_off = (size_t)oj.get_int("w_offset", 0);
      o.b_off = (size_t)oj.get_int("b_offset", 0);
      if (o.h < 1 || o.w < 1 || o.c < 1 || o.kh < 1 || o.kw < 1 || o.stride < 1 || o.pad < 0 || o.cout < 1) {
        *err = "graph manifest: bad op geometry";
        return false;
      }
      if (o.kind == OpKind::Embed || o.kind == OpKind::LayerNorm) {
        o.oh = o.h;
        o.ow = o.w;
        o.cout = o.c;
      } else if (o.kind == OpKind::Attention) {
        o.oh = o.h;
        o.ow = o.w;
        o.cout = o.c / 3;
        if (o.c % 3 || o.heads < 1 || o.cout % o.heads || o.w != 1) {
          *err = "graph manifest: attention expects a packed [S,1,3H] qkv source";
          return false;
        }
      } else if (o.kind == OpKind::AvgPool) {
        o.oh = o.ow = 1;
        o.cout = o.c;
      } else if (o.kind == OpKind::Dense) {
        o.oh = o.ow = 1;
        o.kh = o.kw = 1;
      } else {
        o.oh = (o.h + 2 * o.pad - o.kh) / o.stride + 1;
        o.ow = (o.w + 2 * o.pad - o.kw) / o.stride + 1;
        if (o.kind == OpKind::MaxPool) o.cout = o.c;
      }
      const int64_t in_e = o.kind == OpKind::Embed ? (int64_t)o.h * o.w : (int64_t)o.h * o.w * o.c;
      const int64_t out_e = (int64_t)o.oh * o.ow * o.cout;
      auto buf_ok = [&](int b) { return b == -1 || (b >= 0 && b < d->n_buffers); };
      if (!buf_ok(o.src) || !(o.dst == -2 || (o.dst >= 0 && o.dst < d->n_buffers)) || (o.res != -100 && !buf_ok(o.res)) ||
          o.dst == o.src || o.dst == o.res) {
        *err = "graph manifest: bad buffer index";
        return false;
      }
      const int64_t have = o.src == -1 ? in_elems : written[o.src];
      o.lda = have;
      const bool size_ok = o.kind == OpKind::Dense ? have >= in_e : have == in_e;  // Dense may read the first token only
      if (o.kind == OpKind::Embed && (o.src != -1 || o.vocab < 1 || o.max_pos < o.h || (o.word_off & 255) || (o.pos_off & 255) ||
                                      (o.type_off & 255) || o.word_off + (size_t)o.vocab * o.c * 4 > d->weights_bytes ||
                                      o.pos_off + (size_t)o.max_pos * o.c * 4 > d->weights_bytes ||
                                      o.type_off + (size_t)2 * o.c * 4 > d->weights_bytes)) {
        *err = "graph manifest: bad embed op";
        return false;
      }
      if ((o.kind == OpKind::Embed || o.kind == OpKind::LayerNorm) &&
          ((o.w_off & 255) || (o.b_off & 255) || o.w_off + (size_t)o.c * 4 > d->weights_bytes || o.b_off + (size_t)o.c * 4 > d->weights_bytes)) {
        *err = "graph manifest: LayerNorm gamma/beta out of range";
        return false;
      }
      if (!size_ok || (o.res != -100 && (o.res == -1 ? in_elems : written[o.res]) != out_e)) {
        *err = "graph manifest: op input size does not match its producer";
        return false;
      }
      if (o.kind == OpKind::Conv || o.kind == OpKind::Dense) {
        const size_t wbytes = (size_t)o.kh * o.kw * o.c * o.cout * 4;
        if ((o.w_off & 255) || (o.b_off & 255) || o.w_off + wbytes > d->weights_bytes || o.b_off + (size_t)o.cout * 4 > d->weights_bytes) {
          *err = "graph manifest: weights out of range or misaligned";
          return false;
        }
        const bool direct = o.kh == 1 && o.kw == 1 && o.stride == 1 && o.pad == 0;
        if (!direct) {
          const int64_t ldc = ((int64_t)o.kh * o.kw * o.c + 3) / 4 * 4;
          d->col_elems = std::max<int64_t>(d->col_elems, (int64_t)o.oh * o.ow * ldc);
        }
      }
      if (o.dst == -2) out_elems = out_e;
      else {
        written[o.dst] = out_e;
        d->buf_elems = std::max<int64_t>(d->buf_elems, out_e);
      }
      d->ops.push_back(o);
    }
    if (out_elems < 0 || d->ops.back().dst != -2) {
      *err = "graph manifest: the last op must write the response (dst = -2)";
      return false;
    }
    d->output_shape.clear();
    const GraphOp& last = d->ops.back();
    if (last.oh * last.ow > 1) {
      d->output_shape.push_back(last.oh);
      d->output_shape.push_back(last.ow);
    }
    d->output_shape.push_back(last.cout);
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
