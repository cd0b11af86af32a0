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

}  // namespace tfsc
