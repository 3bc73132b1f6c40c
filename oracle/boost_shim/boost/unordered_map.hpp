// Stand-in for <boost/unordered_map.hpp>: TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Surface used by /root/reference/data/pillars.cpp:264-265,292-303,311-322,332-362:
// find/end/insert({k,v})/at/begin/++/->first/->second.
//
// Iteration order: boost::unordered_map iterates in hash-bucket order, which is
// implementation-defined (Boost version + boost::hash of a 2-double array) and cannot be
// reproduced without Boost.  This stand-in iterates in INSERTION order, which is the
// canonical pillar order this repo declares (DESIGN.md "Pillar order").  So a build of the
// reference against this header pins everything in create_pillars EXCEPT the hash order.
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>
#include <cstdint>
namespace boost {
namespace unordered {
template <class K, class V>
class unordered_map {
  struct KeyHash {
    std::size_t operator()(const K& k) const {
      // FNV-1a over the object bytes (keys here are arrays of doubles holding integers).
      const unsigned char* p = reinterpret_cast<const unsigned char*>(&k);
      std::uint64_t h = 1469598103934665603ull;
      for (std::size_t i = 0; i < sizeof(K); ++i) { h ^= p[i]; h *= 1099511628211ull; }
      return static_cast<std::size_t>(h);
    }
  };
  struct KeyEq {
    bool operator()(const K& a, const K& b) const { return a == b; }
  };
  std::vector<std::pair<K, V>> items_;
  std::unordered_map<K, std::size_t, KeyHash, KeyEq> index_;

 public:
  typedef typename std::vector<std::pair<K, V>>::iterator iterator;
  iterator begin() { return items_.begin(); }
  iterator end() { return items_.end(); }
  iterator find(const K& k) {
    auto it = index_.find(k);
    if (it == index_.end()) return items_.end();
    return items_.begin() + static_cast<std::ptrdiff_t>(it->second);
  }
  void insert(const std::pair<K, V>& kv) {
    if (index_.find(kv.first) != index_.end()) return;
    index_.emplace(kv.first, items_.size());
    items_.push_back(kv);
  }
  V& at(const K& k) {
    auto it = index_.find(k);
    if (it == index_.end()) throw std::out_of_range("unordered_map::at");
    return items_[it->second].second;
  }
  std::size_t size() const { return items_.size(); }
};
}  // namespace unordered
using unordered::unordered_map;
}  // namespace boost
