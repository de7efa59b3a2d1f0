// stand-in (see ../../../README.md) for the reference's rank-aware debug stream: swallows all output
#pragma once
#include <iostream>
#include <map>
#include <sstream>
template <class Stream>
class Ax1ParallelDebugStream {
 public:
  explicit Ax1ParallelDebugStream(std::ostream&) {}
  template <class T>
  Ax1ParallelDebugStream& operator<<(const T&) { return *this; }
  Ax1ParallelDebugStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
  void push(bool) {}
  void pop() {}
};
