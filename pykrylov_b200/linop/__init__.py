"""Linear Operator Type"""
from .linop import *      # noqa: F401,F403
from .linop import (BaseLinearOperator, LinearOperator, IdentityOperator, DiagonalOperator,  # noqa: F401
                    ZeroOperator, ReducedLinearOperator, SymmetricallyReducedLinearOperator,
                    ShapeError, CoordLinearOperator, PysparseLinearOperator, linop_from_ndarray,
                    CsrLinearOperator, DeviceChainOperator, csr_operator, linop_from_scipy, sqrt, null_log)
