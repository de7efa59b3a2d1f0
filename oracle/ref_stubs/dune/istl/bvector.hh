// stand-in (see ../../README.md): the channel set is instantiated with the driver's own vector type
#pragma once
