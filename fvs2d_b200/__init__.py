"""fvs2d_b200 -- B200-native residual + Runge-Kutta hot path of the fvs2d Euler solver.

The product is the C-ABI shared library ``fvs2d_b200/csrc/libfvs2d_gpu.so`` (declared in
``include/fvs2d_gpu.h``); this package is the host-side mirror of the reference's interface
(``input_read``, ``grid_read``, ``time_integration``, ``compute_residual``) over that library.
"""
from .config import Fvs2dConfig, RunInput, read_input, write_input  # noqa: F401
from .meshio import Mesh, read_mesh, write_mesh  # noqa: F401
