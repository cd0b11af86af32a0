#include "lru.h"

namespace tfsc {

bool LRUCache::get(const ModelId& id, CachedModel* out) {
  auto it = map_.find(id);
  if (it == map_.end()) return false;
  list_.splice(list_.begin(), list_, it->second);  // MoveToFront
  if (out) *out = *it->second;
  return true;
}

bool LRUCache::peek(const ModelId& id, CachedModel* out) const {
  auto it = map_.find(id);
  if (it == map_.end()) return false;
  if (out) *out = *it->second;
  return true;
}

int LRUCache::put(const ModelId& id, const CachedModel& m) {
  auto it = map_.find(id);
  if (it != map_.end()) {  // existing: touch only, size NOT updated (:62-64)
    list_.splice(list_.begin(), list_, it->second);
    return 0;
  }
  int ev = ensure_free_bytes(m.size_on_disk);
  list_.push_front(m);
  map_[id] = list_.begin();
  current_ += m.size_on_disk;
  return ev;
}

int LRUCache::ensure_free_bytes(int64_t bytes) {
  int ev = 0;
  while (!list_.empty() && capacity_ - current_ < bytes) {
    CachedModel victim = list_.back();
    current_ -= victim.size_on_disk;
    map_.erase(victim.id);
    list_.pop_back();
    ++ev;
    if (on_evict) on_evict(victim);
  }
  return ev;
}

std::vector<CachedModel> LRUCache::list_models() const {
  return std::vector<CachedModel>(list_.begin(), list_.end());
}

}  // namespace tfsc

struct tfsc_lru {
  tfsc::LRUCache c;
  tfsc_lru(const char* d, int64_t cap) : c(d ? d : "", cap) {}
};

static std::string lru_lines(const std::vector<tfsc::CachedModel>& v) {
  std::string s;
  for (auto& m : v)
    s += m.id.name + "\t" + std::to_string(m.id.version) + "\t" + std::to_string(m.size_on_disk) + "\t" + m.path + "\n";
  return s;
}

extern "C" {
tfsc_lru* tfsc_lru_new(const char* base_dir, int64_t capacity_bytes) { return new tfsc_lru(base_dir, capacity_bytes); }
void tfsc_lru_free(tfsc_lru* c) { delete c; }
int tfsc_lru_put(tfsc_lru* c, const char* name, int64_t version, const char* path, int64_t size) {
  if (!c || !name) return tfsc::fail(TFSC_E_INVALID, "lru_put: bad arguments");
  tfsc::CachedModel m{{name, version}, path ? path : "", size};
  return c->c.put(m.id, m);
}
int tfsc_lru_get(tfsc_lru* c, const char* name, int64_t version, int64_t* size, char* path, size_t cap) {
  if (!c || !name) return tfsc::fail(TFSC_E_INVALID, "lru_get: bad arguments");
  tfsc::CachedModel m;
  if (!c->c.get({name, version}, &m)) return 0;
  if (size) *size = m.size_on_disk;
  if (path) {
    int rc = tfsc::copy_out(m.path, path, cap);
    if (rc < 0) return rc;
  }
  return 1;
}
int tfsc_lru_ensure_free_bytes(tfsc_lru* c, int64_t bytes) { return c ? c->c.ensure_free_bytes(bytes) : TFSC_E_INVALID; }
int64_t tfsc_lru_current_size(const tfsc_lru* c) { return c ? c->c.current_size() : 0; }
int64_t tfsc_lru_capacity(const tfsc_lru* c) { return c ? c->c.capacity() : 0; }
int tfsc_lru_len(const tfsc_lru* c) { return c ? (int)c->c.size() : 0; }
int tfsc_lru_list(const tfsc_lru* c, char* buf, size_t cap) {
  if (!c) return tfsc::fail(TFSC_E_INVALID, "lru_list: bad arguments");
  auto v = c->c.list_models();
  int rc = tfsc::copy_out(lru_lines(v), buf, cap);
  return rc < 0 ? rc : (int)v.size();
}
}
