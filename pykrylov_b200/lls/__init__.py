"""Iterative Methods for Linear Least-Squares Problems"""
from .lsqr import LSQRFramework, LSQR      # noqa: F401
from .lsmr import LSMRFramework            # noqa: F401
