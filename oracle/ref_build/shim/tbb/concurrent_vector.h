// Minimal stand-in for tbb::concurrent_vector, enough to compile the UNMODIFIED
// reference (/root/reference) in a container without Intel TBB.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Not part of the product.
// The reference only needs: sized construction, operator[], push_back from
// several OpenMP threads, iteration, size/clear/shrink_to_fit and copying
// (include/matrix/csr_matrix.hpp:83-84, csr_matrix.tpp:1365-1425).
#pragma once
#include <mutex>
#include <vector>

namespace tbb {

template <typename T> class concurrent_vector {
public:
  typedef typename std::vector<T>::iterator iterator;
  typedef typename std::vector<T>::const_iterator const_iterator;
  typedef T value_type;

  concurrent_vector() {}
  explicit concurrent_vector(size_t n) : items_(n) {}
  concurrent_vector(size_t n, const T &v) : items_(n, v) {}
  concurrent_vector(const concurrent_vector &other) : items_(other.items_) {}
  concurrent_vector &operator=(const concurrent_vector &other) {
    if (this != &other)
      items_ = other.items_;
    return *this;
  }

  T &operator[](size_t i) { return items_[i]; }
  const T &operator[](size_t i) const { return items_[i]; }
  size_t size() const { return items_.size(); }
  bool empty() const { return items_.empty(); }
  iterator begin() { return items_.begin(); }
  iterator end() { return items_.end(); }
  const_iterator begin() const { return items_.begin(); }
  const_iterator end() const { return items_.end(); }
  void clear() { items_.clear(); }
  void shrink_to_fit() { items_.shrink_to_fit(); }

  iterator push_back(const T &v) {
    std::lock_guard<std::mutex> hold(guard_);
    items_.push_back(v);
    return items_.end() - 1;
  }

private:
  std::vector<T> items_;
  std::mutex guard_;
};

} // namespace tbb
