// stand-in (see ../../../README.md): the geometry wrapper header only needs this include to exist; the driver
// supplies its own axis-aligned host geometry
#pragma once
