// oracle/ref_stubs -- TEST INFRASTRUCTURE ONLY (see ../../../README.md).
#pragma once
#include <newton_standins.hh>
