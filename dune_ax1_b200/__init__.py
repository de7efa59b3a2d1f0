"""dune_ax1_b200 -- B200-native PNP Newton hot path of dune-ax1 (CUDA kernels behind a C ABI)."""
from . import api  # noqa: F401
from .api import (Ax1Gpu, Ax1GpuError, Ax1Newton, GridOperator, ISTLBackend_BCGS_ILU0, MembraneFlux,  # noqa: F401
                  NewtonError, default_newton_opts, default_params)
