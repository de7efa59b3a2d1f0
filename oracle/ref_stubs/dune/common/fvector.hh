// stand-in (see ../../README.md): fixed-size vector with the few members the channel headers touch
#pragma once
#include <cstddef>
namespace Dune {
template <class T, int N>
class FieldVector {
 public:
  FieldVector() { for (int i = 0; i < N; ++i) d_[i] = T(); }
  FieldVector(const T& v) { for (int i = 0; i < N; ++i) d_[i] = v; }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  FieldVector& operator=(const T& v) { for (int i = 0; i < N; ++i) d_[i] = v; return *this; }
  static constexpr int size() { return N; }
 private:
  T d_[N];
};
}  // namespace Dune
