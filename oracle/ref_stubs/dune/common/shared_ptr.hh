// stand-in (see ../../README.md)
#pragma once
#include <memory>
namespace Dune { using std::shared_ptr; using std::make_shared; }
