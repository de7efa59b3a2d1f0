// stand-in (see ../../../README.md)
#pragma once
#include <dune_standins.hh>
