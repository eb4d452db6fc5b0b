// Minimal stand-in for tbb::concurrent_unordered_set (see concurrent_vector.h
// in this directory for the why). TEST INFRASTRUCTURE ONLY.
// Used by the reference as the adjacency set of one conflict-graph vertex:
// concurrent insert(), iteration, size() (csr_matrix.tpp:1447-1475, 2047-2049).
#pragma once
#include <mutex>
#include <unordered_set>

namespace tbb {

template <typename T> class concurrent_unordered_set {
public:
  typedef typename std::unordered_set<T>::iterator iterator;
  typedef typename std::unordered_set<T>::const_iterator const_iterator;

  concurrent_unordered_set() {}
  concurrent_unordered_set(const concurrent_unordered_set &other)
      : keys_(other.keys_) {}
  concurrent_unordered_set &operator=(const concurrent_unordered_set &other) {
    if (this != &other)
      keys_ = other.keys_;
    return *this;
  }

  void insert(const T &v) {
    std::lock_guard<std::mutex> hold(guard_);
    keys_.insert(v);
  }
  size_t size() const { return keys_.size(); }
  bool empty() const { return keys_.empty(); }
  size_t count(const T &v) const { return keys_.count(v); }
  iterator begin() { return keys_.begin(); }
  iterator end() { return keys_.end(); }
  const_iterator begin() const { return keys_.begin(); }
  const_iterator end() const { return keys_.end(); }
  void clear() { keys_.clear(); }

private:
  std::unordered_set<T> keys_;
  std::mutex guard_;
};

} // namespace tbb
