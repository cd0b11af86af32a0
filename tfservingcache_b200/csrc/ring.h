// Consistent-hash ring placing (model, version) keys on ring members (GPUs / nodes).
// Replaces stathat.com/c/consistent v1.0.0 as used by pkg/taskhandler/cluster.go:55,111,117;
// the integer arithmetic (CRC-32/IEEE of "<vnode><member>", 20 vnodes, first point strictly
// greater than crc32(key), clockwise distinct walk) is bit-exact with that module.
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace tfsc {

uint32_t crc32_ieee(const void* data, size_t len);

class Ring {
 public:
  static constexpr int kVnodes = 20;  // consistent.New(): NumberOfReplicas
  void set(const std::vector<std::string>& members);  // Consistent.Set
  // Consistent.GetN; empty ring -> returns false
  bool get_n(const std::string& key, int n, std::vector<std::string>* out) const;
  int members() const { return (int)members_.size(); }
  int points() const { return (int)sorted_.size(); }

 private:
  void add(const std::string& m);
  void remove(const std::string& m);
  void rebuild_sorted();
  std::map<uint32_t, std::string> circle_;  // hash -> member (later insert overwrites)
  std::set<std::string> members_;
  std::vector<uint32_t> sorted_;
  std::vector<const std::string*> sorted_member_;  // parallel to sorted_
};

// Replica choice among the GetN candidates (taskhandler.go:91 picks uniformly at random).
//   "random"     the reference's policy: every model ends up cached on all k replicas.
//   "first"      always the primary (ring order).
//   "hot-spread" primary unless the key is hot: a key whose share of recent requests exceeds
//                hot_fraction / members is spread uniformly over its replicas. Keeps one HBM copy of cold
//                models (paging a 1 GB model over PCIe costs ~100x serving it) and k copies of the few hot
//                ones. Deterministic given the seed and the request sequence, so independent processes
//                that see the same stream agree without communicating.
//   "hash"       replica index = crc32(key + "\x01replica") mod k: stateless, so every process (one rank per GPU, requests
//                entering anywhere) sends a key to the same replica and a model is cached once, yet keys spread over
//                all k candidates of the ring.
//   "balanced"   hot-spread + sticky power-of-two-choices: the first request of a cold key binds it to the
//                candidate replica that currently holds the fewest keys (the 20-vnode ring alone is +-16 %
//                uneven in keys per member; every resident model costs one pass over its weights per tick).
class ReplicaPicker {
 public:
  ReplicaPicker(const std::string& policy, uint64_t seed, double hot_fraction = 0.5);
  int pick(const std::string& key, int n_replicas, int members);
  // same, with the identity (any stable integer id) of each candidate; needed by "balanced"
  int pick_ids(const std::string& key, const int* member_ids, int n_replicas, int members);
  const std::string& policy() const { return policy_; }

 private:
  uint64_t next();
  std::string policy_;
  uint64_t rng_;
  double hot_fraction_;
  bool note_and_is_hot(const std::string& key, int members);
  std::map<std::string, uint32_t> counts_;
  uint64_t window_ = 0;
  std::map<std::string, int> bound_;  // balanced: key -> member id
  std::map<int, int> load_;           // balanced: member id -> bound keys
};

}  // namespace tfsc
