// Byte-capacity LRU over (model, version): the ModelCache of pkg/cachemanager/lrucache.go:11-105.
// In this build an entry is a model held in the pinned-host tier; eviction hands the entry back
// to the owner (cache manager) through `on_evict` instead of deleting files (lrucache.go:73-78).
#pragma once
#include <functional>
#include <list>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace tfsc {

struct CachedModel {  // cachemanager.go:45-49 Model
  ModelId id;
  std::string path;
  int64_t size_on_disk = 0;
};

class LRUCache {
 public:
  LRUCache(std::string base_dir, int64_t capacity) : base_dir_(std::move(base_dir)), capacity_(capacity) {}
  bool get(const ModelId& id, CachedModel* out);          // :43-51, touches
  bool peek(const ModelId& id, CachedModel* out) const;   // no touch (new: status queries)
  int put(const ModelId& id, const CachedModel& m);       // :54-65, returns #evicted
  int ensure_free_bytes(int64_t bytes);                   // :68-87, returns #evicted
  std::vector<CachedModel> list_models() const;           // :89-97, MRU -> LRU
  int64_t current_size() const { return current_; }
  int64_t capacity() const { return capacity_; }
  size_t size() const { return map_.size(); }
  const std::string& base_dir() const { return base_dir_; }
  std::function<void(const CachedModel&)> on_evict;

 private:
  std::string base_dir_;
  int64_t capacity_;
  int64_t current_ = 0;
  std::list<CachedModel> list_;  // front = MRU
  std::unordered_map<ModelId, std::list<CachedModel>::iterator, ModelIdHash> map_;
};

}  // namespace tfsc
