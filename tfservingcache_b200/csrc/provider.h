// Model providers: the ModelProvider plug point of pkg/cachemanager/modelprovider.go:3-7
// (LoadModel / ModelSize / Check). LoadModel's destination is the pinned-host tier instead of
// the hostModelPath directory: "copy the tree to local disk" becomes "read weights.bin into
// pinned memory", from where cudaMemcpyAsync pages it into the HBM arena.
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>

#include "json.h"
#include "model.h"

namespace tfsc {

// allocates `bytes` of host memory for a model blob; the returned release functor frees it
using HostAllocFn = std::function<void*(size_t bytes, std::function<void(void*, size_t)>* release)>;

class ModelProvider {
 public:
  virtual ~ModelProvider() = default;
  virtual int64_t model_size(const std::string& name, int64_t version, std::string* err) = 0;     // ModelSize
  virtual std::shared_ptr<HostModel> load_model(const std::string& name, int64_t version, const HostAllocFn& alloc,
                                                std::string* err) = 0;                             // LoadModel
  virtual bool check() { return true; }                                                            // Check
};

// pkg/cachemanager/modelproviders/diskmodelprovider/diskmodelprovider.go
class DiskModelProvider : public ModelProvider {
 public:
  explicit DiskModelProvider(std::string base_dir) : base_dir_(std::move(base_dir)) {}
  // findSrcPathForModel :46-69 -- the directory whose NAME parses (ParseInt) to `version`
  static bool find_src_path(const std::string& model_dir, int64_t version, std::string* out, std::string* err);
  int64_t model_size(const std::string& name, int64_t version, std::string* err) override;
  std::shared_ptr<HostModel> load_model(const std::string& name, int64_t version, const HostAllocFn& alloc,
                                        std::string* err) override;
  bool check() override;

 private:
  std::string base_dir_;
};

// Deterministic synthetic tenants (bench / tests): model "<prefix><j>" gets weights from an
// integer hash of (seedBase + j, tensor, index) -- bit-identical to oracle/models.py.
class SyntheticModelProvider : public ModelProvider {
 public:
  explicit SyntheticModelProvider(const Json& cfg);
  int64_t model_size(const std::string& name, int64_t version, std::string* err) override;
  std::shared_ptr<HostModel> load_model(const std::string& name, int64_t version, const HostAllocFn& alloc,
                                        std::string* err) override;
  static void fill(float* dst, uint32_t seed, uint32_t tensor_id, uint64_t n, float scale, int threads);

 private:
  bool index_of(const std::string& name, int64_t* j) const;
  ModelDesc desc_;
  std::string prefix_;
  int64_t count_ = 0, seed_base_ = 1000;
  double affine_a_ = 0.5, affine_b_ = 2.0;
  int threads_ = 8;
  std::string bad_;  // configuration error reported on first use
};

std::unique_ptr<ModelProvider> create_provider(const Json& cfg, std::string* err);  // main.go:152-185

}  // namespace tfsc
