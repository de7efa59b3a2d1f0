// oracle/ld_prelude.hpp -- TEST INFRASTRUCTURE ONLY.
// Force-included (g++ -include) when the oracle is compiled a second time in EXTENDED precision
// (libax1oracle_ld.so: x87 80-bit long double, 64-bit mantissa, eps = 1.1e-19): every standard header the oracle
// uses is included first, then `double` is re-defined, so the whole restatement -- element kernels, membrane flux,
// channels, assembly, ILU, BiCGStab, Newton -- runs the same operations on the same (double-rounded) literals and
// inputs with 2048 times less rounding error. Used to settle where the truth lies when two double-precision
// implementations (device, oracle) disagree in the worst-conditioned unknown (tests/test_precision_floor.py).
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#define double long double
