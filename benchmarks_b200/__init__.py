"""benchmarks_b200 -- B200-native matrix-free high-order FE operator path.

Python host-side mirror of the reference's operator interface over the C ABI in
include/b200fe.h (libb200fe.so, hand-written sm_100a CUDA).  PyTorch is used only for device
memory, streams and torch.distributed plumbing.  There is no CPU fallback: importing the
package without the built library raises, and every compute call needs a CUDA device.
"""
from ._lib import lib, LIB_PATH, B200feError, check  # noqa: F401
from .mesh import BoxMesh, HangingBoxMesh, basis_1d, QUAD_GAUSS, QUAD_GLL, PARTITION_P4EST, PARTITION_BLOCKS, GHOSTS_MINIMAL, GHOSTS_RELEVANT  # noqa: F401
from .operator import (LaplaceOperator, ReductionControl, SolverCG, NoConvergence, PreconditionChebyshev, PreconditionPMG, PTransfer,  # noqa: F401
                       OP_LAPLACE, OP_MASS, OP_HELMHOLTZ, overlap_permutation)
from .bk import bk1_apply, bk3_apply, bk5_apply, sum_squares, bk_launch_info  # noqa: F401

__all__ = ["lib", "LIB_PATH", "B200feError", "check", "bk1_apply", "bk3_apply", "bk5_apply",
           "sum_squares", "bk_launch_info"]
