// stand-in (see ../../README.md): Dune::DebugStream as a type tag only
#pragma once
#include <iostream>
namespace Dune {
typedef unsigned int DebugLevel;
template <DebugLevel thislevel = 1, DebugLevel dlevel = 1, DebugLevel alevel = 1>
class DebugStream {};
}  // namespace Dune
