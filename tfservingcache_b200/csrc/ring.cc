#include "ring.h"

#include <algorithm>

#include "common.h"

namespace tfsc {

namespace {
struct CrcTable {
  uint32_t t[8][256];
  CrcTable() {
    for (uint32_t n = 0; n < 256; ++n) {
      uint32_t c = n;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : (c >> 1);
      t[0][n] = c;
    }
    for (uint32_t n = 0; n < 256; ++n)
      for (int s = 1; s < 8; ++s) t[s][n] = (t[s - 1][n] >> 8) ^ t[0][t[s - 1][n] & 0xFF];
  }
};
const CrcTable& table() {
  static CrcTable tb;
  return tb;
}
}  // namespace

// slicing-by-8 CRC-32/IEEE (reflected 0xEDB88320, init/xorout 0xFFFFFFFF) == Go ChecksumIEEE
uint32_t crc32_ieee(const void* data, size_t len) {
  const auto& T = table().t;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xFFFFFFFFu;
  while (len >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T[7][lo & 0xFF] ^ T[6][(lo >> 8) & 0xFF] ^ T[5][(lo >> 16) & 0xFF] ^ T[4][lo >> 24] ^
        T[3][hi & 0xFF] ^ T[2][(hi >> 8) & 0xFF] ^ T[1][(hi >> 16) & 0xFF] ^ T[0][hi >> 24];
    p += 8;
    len -= 8;
  }
  while (len--) c = T[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

static uint32_t point(const std::string& member, int idx) {
  std::string k = std::to_string(idx) + member;  // eltKey: strconv.Itoa(idx) + elt
  return crc32_ieee(k.data(), k.size());
}

void Ring::add(const std::string& m) {
  for (int i = 0; i < kVnodes; ++i) circle_[point(m, i)] = m;
  members_.insert(m);
}

void Ring::remove(const std::string& m) {
  for (int i = 0; i < kVnodes; ++i) circle_.erase(point(m, i));
  members_.erase(m);
}

void Ring::rebuild_sorted() {
  sorted_.clear();
  sorted_member_.clear();
  sorted_.reserve(circle_.size());
  for (auto& kv : circle_) {  // std::map iterates ascending == updateSortedHashes
    sorted_.push_back(kv.first);
    sorted_member_.push_back(&kv.second);
  }
}

void Ring::set(const std::vector<std::string>& members) {
  std::vector<std::string> gone;
  for (auto& m : members_)
    if (std::find(members.begin(), members.end(), m) == members.end()) gone.push_back(m);
  for (auto& m : gone) remove(m);
  for (auto& m : members)
    if (!members_.count(m)) add(m);
  rebuild_sorted();
}

bool Ring::get_n(const std::string& key, int n, std::vector<std::string>* out) const {
  out->clear();
  if (sorted_.empty()) return false;
  if ((int)members_.size() < n) n = (int)members_.size();
  if (n <= 0) return true;
  uint32_t h = crc32_ieee(key.data(), key.size());
  size_t i = std::upper_bound(sorted_.begin(), sorted_.end(), h) - sorted_.begin();
  if (i >= sorted_.size()) i = 0;
  size_t start = i;
  out->push_back(*sorted_member_[i]);
  if ((int)out->size() == n) return true;
  for (i = start + 1;; ++i) {
    if (i >= sorted_.size()) i = 0;
    if (i == start) break;
    const std::string& e = *sorted_member_[i];
    if (std::find(out->begin(), out->end(), e) == out->end()) out->push_back(e);
    if ((int)out->size() == n) break;
  }
  return true;
}

ReplicaPicker::ReplicaPicker(const std::string& policy, uint64_t seed, double hot_fraction)
    : policy_(policy), rng_(0x9E3779B97F4A7C15ull ^ (seed * 0xD1342543DE82EF95ull + 1)), hot_fraction_(hot_fraction) {}

uint64_t ReplicaPicker::next() {  // xorshift64*
  rng_ ^= rng_ >> 12;
  rng_ ^= rng_ << 25;
  rng_ ^= rng_ >> 27;
  return rng_ * 0x2545F4914F6CDD1Dull;
}

bool ReplicaPicker::note_and_is_hot(const std::string& key, int members) {
  // sliding-window request share, halved every 64 Ki requests
  uint32_t& c = counts_[key];
  ++c;
  const uint32_t cur = c;
  if (++window_ >= 65536) {
    for (auto it = counts_.begin(); it != counts_.end();) {
      it->second >>= 1;
      if (it->second == 0) it = counts_.erase(it);
      else ++it;
    }
    window_ >>= 1;
  }
  return window_ >= 256 && (double)cur * (double)(members > 0 ? members : 1) > hot_fraction_ * (double)window_;
}

int ReplicaPicker::pick(const std::string& key, int n_replicas, int members) {
  if (n_replicas <= 1 || policy_ == "first") return 0;
  if (policy_ == "random") return (int)(next() % (uint64_t)n_replicas);
  if (policy_ == "hash") {  // stateless: every process picks the same replica for a key without sharing any state
    const std::string k2 = key + "\x01replica";
    return (int)(crc32_ieee(k2.data(), k2.size()) % (uint32_t)n_replicas);
  }
  return note_and_is_hot(key, members) ? (int)(next() % (uint64_t)n_replicas) : 0;
}

int ReplicaPicker::pick_ids(const std::string& key, const int* member_ids, int n_replicas, int members) {
  if (policy_ != "balanced" || n_replicas <= 1 || !member_ids) return pick(key, n_replicas, members);
  if (note_and_is_hot(key, members)) return (int)(next() % (uint64_t)n_replicas);
  auto it = bound_.find(key);
  if (it != bound_.end()) {
    for (int i = 0; i < n_replicas; ++i)
      if (member_ids[i] == it->second) return i;
    load_[it->second]--;  // the bound member left the candidate set (membership change): rebind
    bound_.erase(it);
  }
  int best = 0;
  for (int i = 1; i < n_replicas; ++i)
    if (load_[member_ids[i]] < load_[member_ids[best]]) best = i;
  if (bound_.size() >= ((size_t)1 << 20)) {  // bounded memory: forget all bindings (deterministic for a given call sequence)
    bound_.clear();
    load_.clear();
  }
  bound_[key] = member_ids[best];
  load_[member_ids[best]]++;
  return best;
}

}  // namespace tfsc

using tfsc::Ring;
struct tfsc_picker {
  tfsc::ReplicaPicker p;
  tfsc_picker(const char* policy, uint64_t seed, double hf) : p(policy, seed, hf) {}
};
struct tfsc_ring {
  Ring r;
};

extern "C" {
uint32_t tfsc_crc32_ieee(const void* data, size_t len) { return tfsc::crc32_ieee(data, len); }
tfsc_ring* tfsc_ring_new(void) { return new tfsc_ring(); }
void tfsc_ring_free(tfsc_ring* r) { delete r; }
int tfsc_ring_set(tfsc_ring* r, const char* const* members, int n) {
  if (!r || n < 0 || (n > 0 && !members)) return tfsc::fail(TFSC_E_INVALID, "ring_set: bad arguments");
  std::vector<std::string> v;
  for (int i = 0; i < n; ++i) v.emplace_back(members[i]);
  r->r.set(v);
  return r->r.members();
}
int tfsc_ring_members(const tfsc_ring* r) { return r ? r->r.members() : 0; }
int tfsc_ring_points(const tfsc_ring* r) { return r ? r->r.points() : 0; }
int tfsc_ring_getn(const tfsc_ring* r, const char* key, int n, char* buf, size_t cap) {
  if (!r || !key) return tfsc::fail(TFSC_E_INVALID, "ring_getn: bad arguments");
  std::vector<std::string> out;
  if (!r->r.get_n(key, n, &out)) return tfsc::fail(TFSC_E_EMPTY_RING, "empty circle");
  std::string joined;
  for (size_t i = 0; i < out.size(); ++i) {
    if (i) joined += '\n';
    joined += out[i];
  }
  int rc = tfsc::copy_out(joined, buf, cap);
  return rc < 0 ? rc : (int)out.size();
}
tfsc_picker* tfsc_picker_new(const char* policy, uint64_t seed, double hot_fraction) {
  std::string p = policy ? policy : "random";
  if (p != "random" && p != "first" && p != "hot-spread" && p != "balanced" && p != "hash") {
    tfsc::fail(TFSC_E_INVALID, "unknown proxy.replicaPick '%s'", p.c_str());
    return nullptr;
  }
  return new tfsc_picker(p.c_str(), seed, hot_fraction > 0 ? hot_fraction : 0.5);
}
void tfsc_picker_free(tfsc_picker* p) { delete p; }
int tfsc_picker_pick(tfsc_picker* p, const char* key, int n_replicas, int members) {
  if (!p || !key || n_replicas < 1) return tfsc::fail(TFSC_E_INVALID, "picker_pick: bad arguments");
  return p->p.pick(key, n_replicas, members);
}
int tfsc_picker_pick_ids(tfsc_picker* p, const char* key, const int* member_ids, int n_replicas, int members) {
  if (!p || !key || n_replicas < 1) return tfsc::fail(TFSC_E_INVALID, "picker_pick_ids: bad arguments");
  return p->p.pick_ids(key, member_ids, n_replicas, members);
}
int tfsc_model_key(const char* model_name, const char* version, char* buf, size_t cap) {
  if (!model_name || !version) return tfsc::fail(TFSC_E_INVALID, "model_key: bad arguments");
  return tfsc::copy_out(std::string(model_name) + "##" + version, buf, cap);
}
}
