// stand-in (see ../../README.md): nothing of Dune::FloatCmp is used by the headers compiled into oracle/_ref
#pragma once
