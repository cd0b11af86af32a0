// Offset allocator for the per-GPU HBM arena: one cudaMalloc'd slab sized by gpu.arenaBytes,
// carved into model-sized blocks. First fit over an address-ordered free list with coalescing
// on free; models are MB..GB sized, so the list stays tiny. The slab replaces the reference's
// hostModelPath directory + TF-Serving's own allocations as "where resident models live".
#pragma once
#include <cstddef>
#include <map>

namespace tfsc {

class Arena {
 public:
  void init(size_t capacity, size_t align = 1024) {
    cap_ = capacity / align * align;
    align_ = align;
    used_ = 0;
    free_.clear();
    live_.clear();
    if (cap_) free_[0] = cap_;
  }
  bool alloc(size_t bytes, size_t* off) {
    size_t need = (bytes + align_ - 1) / align_ * align_;
    if (need == 0) need = align_;
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= need) {
        *off = it->first;
        size_t rem = it->second - need, base = it->first;
        free_.erase(it);
        if (rem) free_[base + need] = rem;
        live_[base] = need;
        used_ += need;
        return true;
      }
    }
    return false;
  }
  void release(size_t off) {
    auto it = live_.find(off);
    if (it == live_.end()) return;
    size_t len = it->second;
    live_.erase(it);
    used_ -= len;
    auto nx = free_.lower_bound(off);
    if (nx != free_.end() && off + len == nx->first) {  // merge with next
      len += nx->second;
      nx = free_.erase(nx);
    }
    if (nx != free_.begin()) {  // merge with previous
      auto pv = std::prev(nx);
      if (pv->first + pv->second == off) {
        pv->second += len;
        return;
      }
    }
    free_[off] = len;
  }
  size_t used() const { return used_; }
  size_t capacity() const { return cap_; }
  size_t largest_free() const {
    size_t m = 0;
    for (auto& kv : free_)
      if (kv.second > m) m = kv.second;
    return m;
  }
  size_t blocks() const { return live_.size(); }
  size_t aligned(size_t bytes) const {
    size_t need = (bytes + align_ - 1) / align_ * align_;
    return need ? need : align_;
  }
  // after a compaction: the complete list of live blocks (offset, aligned length), address-ordered and non-overlapping
  void relayout(const std::map<size_t, size_t>& live) {
    live_ = live;
    free_.clear();
    used_ = 0;
    size_t cur = 0;
    for (auto& kv : live_) {
      if (kv.first > cur) free_[cur] = kv.first - cur;
      cur = kv.first + kv.second;
      used_ += kv.second;
    }
    if (cur < cap_) free_[cur] = cap_ - cur;
  }

 private:
  size_t cap_ = 0, align_ = 1024, used_ = 0;
  std::map<size_t, size_t> free_;  // offset -> length
  std::map<size_t, size_t> live_;
};

}  // namespace tfsc
