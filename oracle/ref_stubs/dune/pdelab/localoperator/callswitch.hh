// stand-in (see oracle/ref_stubs/README.md)
#pragma once
#include <dune_standins.hh>
