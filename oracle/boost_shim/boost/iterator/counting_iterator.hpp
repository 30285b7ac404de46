// Minimal stand-in for boost::counting_iterator, written for this repo so that the
// reference's rmq/ headers (which need Boost only for iterator plumbing, no arithmetic)
// compile in an image without Boost.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#pragma once
#include <cstddef>
#include <iterator>
namespace boost {
template <class T>
class counting_iterator {
  T v_;
 public:
  using iterator_category = std::random_access_iterator_tag;
  using value_type = T;
  using difference_type = std::ptrdiff_t;
  using pointer = const T*;
  using reference = T;
  counting_iterator() : v_() {}
  explicit counting_iterator(T v) : v_(v) {}
  T operator*() const { return v_; }
  T operator[](difference_type d) const { return v_ + d; }
  counting_iterator& operator++() { ++v_; return *this; }
  counting_iterator operator++(int) { counting_iterator t(*this); ++v_; return t; }
  counting_iterator& operator--() { --v_; return *this; }
  counting_iterator operator--(int) { counting_iterator t(*this); --v_; return t; }
  counting_iterator& operator+=(difference_type d) { v_ += d; return *this; }
  counting_iterator& operator-=(difference_type d) { v_ -= d; return *this; }
  friend counting_iterator operator+(counting_iterator a, difference_type d) { return counting_iterator(a.v_ + d); }
  friend counting_iterator operator-(counting_iterator a, difference_type d) { return counting_iterator(a.v_ - d); }
  friend difference_type operator-(counting_iterator a, counting_iterator b) { return (difference_type)(a.v_ - b.v_); }
  friend bool operator==(counting_iterator a, counting_iterator b) { return a.v_ == b.v_; }
  friend bool operator!=(counting_iterator a, counting_iterator b) { return a.v_ != b.v_; }
  friend bool operator<(counting_iterator a, counting_iterator b) { return a.v_ < b.v_; }
};
}  // namespace boost
