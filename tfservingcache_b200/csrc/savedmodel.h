// TensorFlow SavedModel ingestion without TensorFlow: `saved_model.pb` + tensor-bundle
// `variables/variables.{index,data-*}` -> the native "tfsc-b200-v1" manifest + weights blob the disk provider pages
// into HBM. Formats restated from their public definitions, field numbers as in the reference's generated protos:
// saved_model / meta_graph / graph / node_def (proto/tensorflow/core/protobuf/{saved_model,meta_graph}.pb.go,
// core/framework/{graph,node_def}.pb.go), BundleHeaderProto / BundleEntryProto
// (proto/tensorflow/core/protobuf/tensor_bundle.pb.go:63-66,121-137), LevelDB table layout of variables.index.
// Recognised graphs: y = a*x + b with scalar variables (half_plus_two) and MatMul + BiasAdd/Add (+ Relu) chains.
// Format parity is unpinned (no TensorFlow-written file is available to test against): tests use files written by
// tests/test_savedmodel.py and require byte-identical output with the Python importer (tfservingcache_b200/savedmodel.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace tfsc {

struct SavedModelBundle {
  std::string manifest_json;   // tfsc_model.json text
  std::vector<char> weights;   // weights.bin bytes
};

bool savedmodel_present(const std::string& version_dir);
bool savedmodel_import(const std::string& version_dir, SavedModelBundle* out, std::string* err);
uint32_t crc32c(const void* data, size_t len);

}  // namespace tfsc
