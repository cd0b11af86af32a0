// Model bundle description ("tfsc-b200-v1"): what the provider hands to the cache manager and
// what the executor runs. A bundle is <baseDir>/<name>/<version>/{tfsc_model.json, weights.bin};
// weights.bin is copied verbatim into pinned host memory and from there into the HBM arena, so
// offsets in the manifest are valid in all three places (256-byte aligned tensors).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "common.h"
#include "json.h"

namespace tfsc {

enum class Template { Affine, Mlp, Graph };

// One node of a "graph" bundle (conv nets): activations are NHWC fp32, `src`/`dst`/`res` index scratch
// activation buffers (-1 = the request tensor, -2 = the response tensor). Weights: conv/dense kernel
// flattened to [kh*kw*c, cout] row-major (= TF HWIO / dense layout) at w_off, bias (folded BN) at b_off.
enum class OpKind { Conv, MaxPool, AvgPool, Dense, Embed, LayerNorm, Attention };
struct GraphOp {
  OpKind kind = OpKind::Conv;
  int src = -1, dst = 0, res = -100;  // res = -100: no residual input
  int h = 1, w = 1, c = 1;            // input H, W, C per image
  int kh = 1, kw = 1, stride = 1, pad = 0, cout = 1, oh = 1, ow = 1;
  int act = 0;                        // 0 none, 1 relu, 2 gelu(erf), 3 tanh
  size_t w_off = 0, b_off = 0;        // kernel / bias; LayerNorm + Embed: gamma / beta
  // transformer ops (a "image" is a sequence: h = S tokens, w = 1, c = hidden width)
  int heads = 1, vocab = 0, max_pos = 0;
  size_t word_off = 0, pos_off = 0, type_off = 0;  // Embed tables [vocab,c], [max_pos,c], [2,c]
  float eps = 1e-12f;
  int64_t lda = 0;                    // Dense: elements per image of the source (> c selects the first token)
};

struct DenseLayer {
  int in = 0, out = 0;
  bool relu = false;
  size_t w_off = 0, b_off = 0;  // bytes into the blob; W row-major [in,out] fp32, b[out]
};

// classify / regress signatures of a model (tensorflow/serving/{classify,regress}): their input is a list of tf.Example,
// `feature` names the float feature that holds one input row per example (half_plus_two: "x")
struct ExtraSignature {
  std::string name;     // e.g. "regress_x_to_y"
  int method = 0;       // 1 = classify, 2 = regress
  std::string feature;
};

struct ModelDesc {
  Template tmpl = Template::Mlp;
  std::vector<ExtraSignature> extra_sigs;
  std::vector<DenseLayer> layers;  // Mlp
  size_t a_off = 0, b_off = 0;     // Affine scalars
  size_t weights_bytes = 0;
  std::string input_name = "x", output_name = "y";
  // elements per batch row; 0 = elementwise / any shape (Affine)
  int64_t in_dim = 0, out_dim = 0;
  int max_width = 0;  // widest activation (for scratch sizing)
  // Graph
  std::vector<GraphOp> ops;
  int n_buffers = 0;
  int64_t buf_elems = 0;  // largest activation, elements per image
  int64_t col_elems = 0;  // largest im2col matrix, elements per image
  std::vector<int64_t> input_shape;   // per image, e.g. [224,224,3]
  std::vector<int64_t> output_shape;  // per image, e.g. [1000]
  int input_dtype = TFSC_DT_FLOAT;    // TFSC_DT_INT32 for token-id inputs (BERT)
  // bytes of executor scratch (activation buffers + im2col) for `rows` images / batch rows
  size_t scratch_bytes(int64_t rows) const;
};

bool parse_manifest(const Json& j, ModelDesc* d, std::string* err);
ModelDesc make_mlp_desc(const std::vector<int>& dims, const std::vector<std::string>& activations);
ModelDesc make_affine_desc();
std::string manifest_json(const ModelDesc& d);

// A model held in the host tier (pinned memory when a CUDA device is present).
struct HostModel {
  ModelId id;
  ModelDesc desc;
  void* data = nullptr;
  size_t bytes = 0;
  std::function<void(void*, size_t)> release;  // returns the block to its pool
  ~HostModel() {
    if (data && release) release(data, bytes);
  }
};

}  // namespace tfsc
