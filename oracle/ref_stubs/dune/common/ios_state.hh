// oracle/ref_stubs -- TEST INFRASTRUCTURE ONLY (see ../../README.md). Stand-in for dune-common's stream-state saver.
#pragma once
#include <ios>
namespace Dune {
class ios_base_all_saver {
 public:
  explicit ios_base_all_saver(std::ios_base& s) : s_(s), f_(s.flags()), p_(s.precision()), w_(s.width()) {}
  ~ios_base_all_saver() { s_.flags(f_); s_.precision(p_); s_.width(w_); }
 private:
  std::ios_base& s_;
  std::ios_base::fmtflags f_;
  std::streamsize p_, w_;
};
}  // namespace Dune
