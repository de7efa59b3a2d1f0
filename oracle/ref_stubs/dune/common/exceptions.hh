// stand-in (see ../../README.md): exception types and the DUNE_THROW macro
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace Dune {
class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m = "") : std::runtime_error(m) {}
};
class NotImplemented : public Exception { public: using Exception::Exception; };
class RangeError : public Exception { public: using Exception::Exception; };
}  // namespace Dune
#define DUNE_THROW(E, m) do { std::ostringstream dune_throw_os_; dune_throw_os_ << m; throw E(dune_throw_os_.str()); } while (0)
