// stand-in (see ../../README.md): fixed-size vector / matrix with the members the reference headers touch
#pragma once
#include <ostream>
#include <cmath>
#include <cstddef>
namespace Dune {
// iterator over the entries / rows of a small dense container that knows its position (dune's index())
template <class C, class T>
class IndexedIterator {
 public:
  IndexedIterator(C* c, std::size_t i) : c_(c), i_(i) {}
  T& operator*() const { return (*c_)[i_]; }
  T* operator->() const { return &(*c_)[i_]; }
  std::size_t index() const { return i_; }
  IndexedIterator& operator++() { ++i_; return *this; }
  bool operator!=(const IndexedIterator& o) const { return i_ != o.i_; }
  bool operator==(const IndexedIterator& o) const { return i_ == o.i_; }
 private:
  C* c_;
  std::size_t i_;
};
template <class T, int N>
class FieldVector {
 public:
  typedef T value_type;
  enum { dimension = N };
  FieldVector() { for (int i = 0; i < N; ++i) d_[i] = T(); }
  FieldVector(const T& v) { for (int i = 0; i < N; ++i) d_[i] = v; }
  T& operator[](std::size_t i) { return d_[i]; }
  const T& operator[](std::size_t i) const { return d_[i]; }
  FieldVector& operator=(const T& v) { for (int i = 0; i < N; ++i) d_[i] = v; return *this; }
  FieldVector& operator+=(const FieldVector& o) { for (int i = 0; i < N; ++i) d_[i] += o.d_[i]; return *this; }
  FieldVector& operator-=(const FieldVector& o) { for (int i = 0; i < N; ++i) d_[i] -= o.d_[i]; return *this; }
  FieldVector& operator*=(const T& a) { for (int i = 0; i < N; ++i) d_[i] *= a; return *this; }
  FieldVector& operator/=(const T& a) { for (int i = 0; i < N; ++i) d_[i] /= a; return *this; }
  FieldVector& axpy(const T& a, const FieldVector& x) { for (int i = 0; i < N; ++i) d_[i] += a * x.d_[i]; return *this; }
  T operator*(const FieldVector& o) const { T s = T(); for (int i = 0; i < N; ++i) s += d_[i] * o.d_[i]; return s; }
  T two_norm() const { T s = T(); for (int i = 0; i < N; ++i) s += d_[i] * d_[i]; return std::sqrt(s); }
  T infinity_norm() const { T m = T(); for (int i = 0; i < N; ++i) m = std::fabs(d_[i]) > m ? std::fabs(d_[i]) : m; return m; }
  static constexpr int size() { return N; }
  typedef IndexedIterator<FieldVector, T> Iterator;
  Iterator begin() { return Iterator(this, 0); }
  Iterator end() { return Iterator(this, N); }
 private:
  T d_[N];
};
// a vector of length one converts to its entry (dune-common does the same)
template <class T>
class FieldVector<T, 1> {
 public:
  typedef T value_type;
  enum { dimension = 1 };
  FieldVector() : d_(T()) {}
  FieldVector(const T& v) : d_(v) {}
  T& operator[](std::size_t) { return d_; }
  const T& operator[](std::size_t) const { return d_; }
  FieldVector& operator=(const T& v) { d_ = v; return *this; }
  FieldVector& operator+=(const T& v) { d_ += v; return *this; }
  FieldVector& operator-=(const T& v) { d_ -= v; return *this; }
  FieldVector& operator*=(const T& v) { d_ *= v; return *this; }
  operator T&() { return d_; }
  operator const T&() const { return d_; }
  static constexpr int size() { return 1; }
 private:
  T d_;
};
template <class T, int N, int M>
class FieldMatrix {
 public:
  FieldMatrix() {}
  FieldMatrix(const T& v) { for (int i = 0; i < N; ++i) r_[i] = v; }
  FieldVector<T, M>& operator[](std::size_t i) { return r_[i]; }
  const FieldVector<T, M>& operator[](std::size_t i) const { return r_[i]; }
  typedef IndexedIterator<FieldMatrix, FieldVector<T, M> > RowIterator;
  typedef RowIterator Iterator;
  RowIterator begin() { return RowIterator(this, 0); }
  RowIterator end() { return RowIterator(this, N); }
  template <class X, class Y>
  void mv(const X& x, Y& y) const {                       // y = A x
    for (int i = 0; i < N; ++i) { y[i] = T(); for (int j = 0; j < M; ++j) y[i] += r_[i][j] * x[j]; }
  }
  template <class X, class Y>
  void umv(const X& x, Y& y) const {                      // y += A x
    for (int i = 0; i < N; ++i) for (int j = 0; j < M; ++j) y[i] += r_[i][j] * x[j];
  }
 private:
  FieldVector<T, M> r_[N];
};
template <class K, int n>
std::ostream& operator<<(std::ostream& os, const FieldVector<K, n>& v) {
  for (int i = 0; i < n; ++i) os << (i ? " " : "") << v[i];
  return os;
}
}  // namespace Dune
